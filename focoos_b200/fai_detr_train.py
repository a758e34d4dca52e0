"""Training-mode forward of FAIDetr: the graph `FAIDetr.forward` runs under `model.train()` in the reference
(models/fai_detr/modelling.py:1344-1358 with `self.training`), built from focoos_b200.autograd_ops so that both the forward and
the backward pass execute the hand-written kernels.  Differences to the eval engine (fai_detr.DetrEngine), all dictated by the
reference's training semantics:

  * BatchNorm uses BATCH statistics and updates its running buffers (the registry configs ship freeze_norm=false), so nothing is
    folded: conv -> BN(+residual)(+act) per ConvNormLayer; RepVGG blocks run un-re-parameterised (two branches, modelling.py:39-45);
  * all six decoder layers and the encoder proposals emit predictions (aux losses, :1005-1011, :1259-1261);
  * `target` and the initial reference points are detached (:1229-1231); reference points are detached between layers (:1018).

Layout NHWC fp32; parameters are the nn.Module tree's own tensors (autograd leaves, the optimiser's flat-buffer views).
torch is used for plumbing only (concat/slice views, the [B,300,4] box arithmetic, index gathers).
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

from . import autograd_ops as A
from . import ops
from .fai_detr import generate_anchors

_ACT = {None: ops.ACT_NONE, "relu": ops.ACT_RELU, "silu": ops.ACT_SILU}


def _inverse_sigmoid(x, eps=1e-5):  # nn/layers/functional.py:4-6
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


class DetrTrainGraph:
    def __init__(self, model, precision: str = "fp32_tc"):
        assert precision in ("fp32", "fp32_tc", "amp")
        self.m, self.prec = model, precision
        self._const = {}
        self.taps = None  # optional dict: named intermediate tensors (detached) for parity debugging
        self.forced_topk = None  # optional [B, num_queries] int tensor: use this query selection instead of the top-k (teacher forcing in parity tests)

    @property
    def last_topk(self):
        """query selection of the last forward, on the host (parity tests)"""
        t = getattr(self, "_last_topk_dev", None)
        return None if t is None else t.cpu()

    def _tap(self, name, t):
        if self.taps is not None:
            self.taps[name] = t.detach()

    # ---- building blocks ---------------------------------------------------------------------------------------------
    def cnl(self, x, layer, res=None, act="default"):
        """ConvNormLayer.forward (nn/layers/conv.py:93-97) with train-mode BN; `res` is added before the activation."""
        k = layer.conv.kernel_size[0]
        y = A.conv2d(x, layer.conv.weight, None, layer.conv.stride[0], (k - 1) // 2, self.prec)
        a = layer.act_name if act == "default" else act
        return self.bn(y, layer.norm, res, _ACT[a])

    def bn(self, y, norm, res=None, act=ops.ACT_NONE):
        """train-mode BatchNorm of one ConvNormLayer: frozen running statistics (FrozenBatchNorm2d: backbone_config.freeze_norm -> norm.frozen_stats,
        or TrainerArgs.freeze_bn -> model.freeze_bn), statistics over all data-parallel ranks (model.sync_bn, trainer.py:334), or batch statistics"""
        import torch.distributed as dist
        frozen = bool(getattr(self.m, "freeze_bn", False)) or bool(getattr(norm, "frozen_stats", False))
        sync = bool(getattr(self.m, "sync_bn", False)) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        return A.batch_norm_train(y, norm, res, act, sync_group=True if sync else None, frozen=frozen)

    def bottleneck(self, x, blk):
        out = self.cnl(x, blk.branch2a)
        out = self.cnl(out, blk.branch2b)
        if blk.shortcut:
            short = x
        elif blk.stride == 2:
            short = self.cnl(A.AvgPoolFn.apply(x), blk.short.conv)
        else:
            short = self.cnl(x, blk.short)
        return self.cnl(out, blk.branch2c, res=short, act="relu")  # relu(BN(conv) + short)  (resnet.py:118-119)

    def backbone(self, x):
        bb = self.m.pixel_decoder.backbone
        for layer in bb.conv1:
            x = self.cnl(x, layer)
        x = A.MaxPoolFn.apply(x)
        outs = []
        for stage in bb.res_layers:
            for blk in stage.blocks:
                x = self.bottleneck(x, blk)
            outs.append(x)
        return outs[1:]  # res3, res4, res5

    def csp(self, x, blk):
        x1 = self.cnl(x, blk.conv1)
        for rep in blk.bottlenecks:
            x1 = A.AddActFn.apply(self.cnl(x1, rep.conv1, act=None), self.cnl(x1, rep.conv2, act=None), ops.ACT_SILU)
        x2 = self.cnl(x, blk.conv2)
        return A.AddActFn.apply(x1, x2, ops.ACT_NONE)

    def mha(self, q_in, k_in, v_in, attn):
        d, h = attn.embed_dim, attn.num_heads
        w, b = attn.in_proj_weight, attn.in_proj_bias
        q = A.linear(q_in, w[:d], b[:d], precision=self.prec)
        k = A.linear(k_in, w[d:2 * d], b[d:2 * d], precision=self.prec)
        v = A.linear(v_in, w[2 * d:], b[2 * d:], precision=self.prec)
        o = A.AttentionFn.apply(q, k, v, h, 1.0 / math.sqrt(d // h), self.prec in ("fp32_tc", "amp"))
        return A.linear(o, attn.out_proj.weight, attn.out_proj.bias, precision=self.prec)

    def mlp(self, x, mlp):
        n = len(mlp.layers)
        for i, l in enumerate(mlp.layers):
            x = A.linear(x, l.weight, l.bias, ops.ACT_RELU if i < n - 1 else ops.ACT_NONE, self.prec)
        return x

    def _aifi_pos(self, h, w, dev):
        key = ("pos", h, w, str(dev))
        if key not in self._const:
            from .fai_detr import aifi_position_embedding
            self._const[key] = aifi_position_embedding(h, w, self.m.config.pixel_decoder_feat_dim // 2).to(dev)
        return self._const[key]

    # ---- encoder (modelling.py:297-347) ----------------------------------------------------------------------------------
    def encoder(self, feats):
        pd = self.m.pixel_decoder
        proj = []
        for f, ip in zip(feats, pd.input_proj):
            proj.append(self.bn(A.conv2d(f, ip[0].weight, None, 1, 0, self.prec), ip[1]))
        B, h, w, C = proj[2].shape
        src = proj[2].reshape(B, h * w, C)
        pos = self._aifi_pos(h, w, src.device)[None].expand(B, -1, -1).contiguous()
        lay = pd.encoder[0].layers[0]
        qk = A.AddActFn.apply(src, pos, ops.ACT_NONE)
        src = A.layer_norm(src, lay.norm1, res=self.mha(qk, qk, src, lay.self_attn))
        f = A.linear(A.AddActFn.apply(A.linear(src, lay.linear1.weight, lay.linear1.bias, precision=self.prec), None, ops.ACT_GELU), lay.linear2.weight,
                     lay.linear2.bias, precision=self.prec)
        src = A.layer_norm(src, lay.norm2, res=f)
        proj[2] = src.reshape(B, h, w, C)
        inner = [proj[2]]
        for idx in (2, 1):
            hi = self.cnl(inner[0], pd.lateral_convs[2 - idx])
            inner[0] = hi
            lo = proj[idx - 1]
            up = A.ResizeFn.apply(hi, (lo.shape[1], lo.shape[2]))
            inner.insert(0, self.csp(torch.cat([up, lo], -1), pd.fpn_blocks[2 - idx]))
        outs = [inner[0]]
        for idx in range(2):
            hi = inner[idx + 1]
            down = self.cnl(A.ResizeFn.apply(outs[-1], (hi.shape[1], hi.shape[2])), pd.downsample_convs[idx])
            outs.append(self.csp(torch.cat([down, hi], -1), pd.pan_blocks[idx]))
        return outs[::-1]  # [1/32, 1/16, 1/8]; the dead mask_features conv (:347) is not executed - its output feeds nothing

    # ---- predictor (modelling.py:1145-1263) ------------------------------------------------------------------------------
    def predictor(self, feats) -> Dict:
        tp = self.m.head.predictor
        toks, shapes = [], []
        for f, ip in zip(feats, tp.input_proj):
            y = self.bn(A.conv2d(f, ip.conv.weight, None, 1, 0, self.prec), ip.norm)
            B, h, w, C = y.shape
            toks.append(y.reshape(B, h * w, C))
            shapes.append((h, w))
        memory = torch.cat(toks, 1)
        self._tap("memory", memory)
        dev = memory.device
        key = ("anchors", tuple(shapes), str(dev))
        if key not in self._const:
            a, valid = generate_anchors(shapes)
            self._const[key] = (a.to(dev), valid.to(dev))
        anchors, valid = self._const[key]
        mem_v = memory * valid.view(1, -1, 1).to(memory.dtype)
        om = A.layer_norm(A.linear(mem_v, tp.enc_output[0].weight, tp.enc_output[0].bias, precision=self.prec), tp.enc_output[1])
        enc_cls = A.linear(om, tp.enc_score_classifier.weight, tp.enc_score_classifier.bias, precision=self.prec)
        enc_box_unact = self.mlp(om, tp.enc_bbox_classifier) + anchors.view(1, -1, 4)
        with torch.no_grad():
            scores = ops.rowmax(enc_cls.detach())
            _, topk_ind = ops.topk(scores.reshape(B, -1).contiguous(), tp.num_queries)
            if self.forced_topk is not None:
                topk_ind = self.forced_topk.to(device=dev, dtype=torch.int32)
            idx = topk_ind.to(torch.int64)
        ref_unact = enc_box_unact.gather(1, idx.unsqueeze(-1).expand(-1, -1, 4))
        enc_topk_bboxes = torch.sigmoid(ref_unact)
        enc_topk_logits = enc_cls.gather(1, idx.unsqueeze(-1).expand(-1, -1, enc_cls.shape[-1]))
        target = om.gather(1, idx.unsqueeze(-1).expand(-1, -1, om.shape[-1])).detach()
        ref_unact = ref_unact.detach()
        self._tap("om", om); self._tap("target", target); self._tap("ref_unact", ref_unact); self._tap("enc_cls", enc_cls)

        # decoder (:969-1020)
        out = target
        ref_detach = torch.sigmoid(ref_unact)
        ref_points = None
        dec_boxes, dec_logits = [], []
        for i, layer in enumerate(tp.decoder.layers):
            pos = self.mlp(ref_detach, tp.query_pos_head)
            self._tap(f"dec{i}.pos", pos)
            qk = A.AddActFn.apply(out, pos, ops.ACT_NONE)
            out = A.layer_norm(out, layer.norm1, res=self.mha(qk, qk, out, layer.self_attn))
            self._tap(f"dec{i}.after_sa", out)
            ca = layer.cross_attn
            value = A.linear(memory, ca.value_proj.weight, ca.value_proj.bias, precision=self.prec)
            q = A.AddActFn.apply(out, pos, ops.ACT_NONE)
            oa = torch.cat([A.linear(q, ca.sampling_offsets.weight, ca.sampling_offsets.bias, precision=self.prec),
                            A.linear(q, ca.attention_weights.weight, ca.attention_weights.bias, precision=self.prec)], -1)
            sampled = A.MSDAFn.apply(value, oa, ref_detach, shapes, tp.num_points, tp.nhead)
            self._tap(f"dec{i}.value", value); self._tap(f"dec{i}.oa", oa); self._tap(f"dec{i}.sampled", sampled)
            out = A.layer_norm(out, layer.norm2, res=A.linear(sampled, ca.output_proj.weight, ca.output_proj.bias, precision=self.prec))
            ff = A.linear(A.linear(out, layer.linear1.weight, layer.linear1.bias, ops.ACT_RELU, self.prec), layer.linear2.weight, layer.linear2.bias, precision=self.prec)
            out = A.layer_norm(out, layer.norm3, res=ff)
            self._tap(f"dec{i}.out", out)
            delta = self.mlp(out, tp.dec_bbox_classifier[i])
            inter = torch.sigmoid(delta + _inverse_sigmoid(ref_detach))
            dec_logits.append(A.linear(out, tp.dec_score_classifier[i].weight, tp.dec_score_classifier[i].bias, precision=self.prec))
            dec_boxes.append(inter if i == 0 else torch.sigmoid(delta + _inverse_sigmoid(ref_points)))
            ref_points = inter
            ref_detach = inter.detach()
        res = {"pred_logits": dec_logits[-1], "pred_boxes": dec_boxes[-1],
               "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(dec_logits[:-1], dec_boxes[:-1])]}
        res["aux_outputs"].append({"pred_logits": enc_topk_logits, "pred_boxes": enc_topk_bboxes})
        res["_topk_ind"] = topk_ind
        self._last_topk_dev = topk_ind.detach()  # read back lazily (last_topk): a .cpu() here would stall the launch queue in every training step
        return res

    def forward(self, images) -> Dict:
        """images: [B,3,H,W] float 0..255 (the reference's input) or [B,H,W,3] uint8."""
        m = self.m
        if images.dtype == torch.uint8:
            x = images.to(torch.float32)
        else:
            x = images.to(torch.float32).permute(0, 2, 3, 1)
        mean, std = m.pixel_mean.view(1, 1, 1, 3), m.pixel_std.view(1, 1, 1, 3)
        x = ((x - mean) / std).contiguous()
        return self.predictor(self.encoder(self.backbone(x)))
