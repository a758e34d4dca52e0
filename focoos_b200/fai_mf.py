"""FAIMaskFormer (Mask2Former-style segmenter) — host-side mirror of `focoos/models/fai_mf/modelling.py`.

Same pattern as `fai_detr.py`: the module tree only holds parameters under the reference's state_dict keys
(SURVEY Appendix B, 937 entries for fai-mf-l-coco-ins); `FAIMaskFormer.forward` runs `MFEngine`, a fused NHWC graph:

  * ResNet-101-vd backbone (shared kernels / packing with FAIDetr),
  * TransformerFPN pixel decoder: 1x1 input_proj -> 6 PRE-norm encoder layers with the normalised sine embedding
    (nn/layers/position_encoding.py:45-74) -> final LayerNorm -> 3x3+BN+ReLU; lateral 1x1+BN, nearest x2 upsample + add
    fused in one kernel, 3x3+BN+ReLU; mask_features 3x3 (fai_mf/modelling.py:348-369),
  * MultiScaleMaskedTransformerDecoder: 9 x (masked cross-attention -> self-attention -> FFN), pre-norm; the boolean
    attention mask is built on the device from the previous mask prediction (bilinear resize of the mask logits, `< 0`,
    all-masked rows released) and consumed by a streaming tensor-core attention kernel — the [B*8, Q, HW] bool tensor of the
    reference (:510-513) is never replicated per head,
  * PredictionHeads: LN -> class Linear / mask MLP -> per-image mask GEMM `bqc,bchw->bqhw` on tensor cores, executed
    dec_layers+1 times; intermediate class logits (dead in eval) are skipped,
  * head: softmax[..., :-1]; sigmoid at 1/4 resolution THEN bilinear upsample to the input size, written as the reference's
    [B,Q,H,W] fp32 probability tensor by one kernel (:618-619,722-723).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, fields
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .fai_detr import _split3_weights, MLP, DetrEngine, ResNet, _bn_fold, _Conv, _CriterionStub, _enable_split3, _Linear
from .ports import ModelOutput, ResnetConfig


@dataclass
class MaskFormerConfig:
    """models/fai_mf/config.py (same field names / defaults as the registry JSONs use)."""

    backbone_config: ResnetConfig = field(default_factory=lambda: ResnetConfig(depth=101))
    num_classes: int = 80
    num_queries: int = 100
    resolution: Optional[int] = 1024
    pixel_mean: List[float] = field(default_factory=lambda: [123.675, 116.28, 103.53])
    pixel_std: List[float] = field(default_factory=lambda: [58.395, 57.12, 57.375])
    size_divisibility: int = 0
    pixel_decoder_out_dim: int = 256
    pixel_decoder_feat_dim: int = 256
    pixel_decoder_transformer_layers: int = 6
    pixel_decoder_transformer_dropout: float = 0.0
    pixel_decoder_transformer_nheads: int = 8
    pixel_decoder_transformer_dim_feedforward: int = 1024
    transformer_predictor_out_dim: int = 256
    transformer_predictor_hidden_dim: int = 256
    transformer_predictor_dec_layers: int = 9
    transformer_predictor_dim_feedforward: int = 2048
    head_out_dim: int = 256
    cls_sigmoid: bool = False
    postprocessing_type: str = "instance"
    mask_threshold: float = 0.5
    predict_all_pixels: bool = False
    use_mask_score: bool = True
    threshold: float = 0.5
    top_k: int = 100
    criterion_deep_supervision: bool = True
    criterion_eos_coef: float = 0.1
    criterion_num_points: int = 12544
    weight_dict_loss_dice: int = 5
    weight_dict_loss_mask: int = 5
    weight_dict_loss_ce: int = 2
    matcher_cost_class: int = 2
    matcher_cost_mask: int = 5
    matcher_cost_dice: int = 5

    @classmethod
    def from_dict(cls, d: dict) -> "MaskFormerConfig":
        d = dict(d)
        bc = d.pop("backbone_config", {}) or {}
        if isinstance(bc, dict):
            bc = ResnetConfig(**{k: v for k, v in bc.items() if k in {f.name for f in fields(ResnetConfig)}})
        unknown = set(d) - {f.name for f in fields(cls)}
        if unknown:
            raise ValueError(f"Invalid parameters for MaskFormerConfig: {sorted(unknown)}")
        return cls(backbone_config=bc, **d)


@dataclass
class MaskFormerModelOutput(ModelOutput):
    """models/fai_mf/ports.py: field order (masks, logits, loss)."""

    masks: torch.Tensor  # [B, Q, H, W] fp32 probabilities at the input resolution
    logits: torch.Tensor  # [B, Q, num_classes] softmax probabilities without the no-object column
    loss: Optional[dict] = None


class LazyMasks:
    """The model output `masks` before the final sigmoid + bilinear upsampling (fai_mf/modelling.py:619,722-723): low-resolution logits
    NHWC [B,h,w,Qp] plus the target size.  `materialize()` gives the reference's [B,Q,H,W] fp32 probabilities; a processor that understands
    this object fuses the upsampling into its own reduction instead (semantic argmax: 13.4 GB of HBM traffic avoided at config 4).
    Enabled by `model.lazy_masks = True` (FocoosModel sets it: it owns model + processor); plain `model(images)` returns tensors."""

    def __init__(self, logits_nhwc: torch.Tensor, num_queries: int, size):
        self.logits, self.num_queries, self.size = logits_nhwc, num_queries, (int(size[0]), int(size[1]))

    @property
    def shape(self):
        return (self.logits.shape[0], self.num_queries, *self.size)

    @property
    def device(self):
        return self.logits.device

    def materialize(self) -> torch.Tensor:
        return ops.mask_sigmoid_upsample(self.logits, self.num_queries, self.size)


# ---- parameter containers -------------------------------------------------------------------------
class _ConvBN(nn.Conv2d):
    """nn/layers/conv.py:22 `Conv2d` wrapper whose norm is a CHILD module (`<name>.weight`, `<name>.norm.*`)."""

    def __init__(self, cin, cout, k, bias=False, norm=True):
        super().__init__(cin, cout, k, padding=(k - 1) // 2, bias=bias)
        self.norm = nn.BatchNorm2d(cout) if norm else None


class _EncLayer(nn.Module):
    def __init__(self, d, nhead, dff):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, nhead, 0.0)
        self.linear1, self.linear2 = nn.Linear(d, dff), nn.Linear(dff, d)
        self.norm1, self.norm2 = nn.LayerNorm(d), nn.LayerNorm(d)


class _Encoder(nn.Module):
    def __init__(self, d, nhead, dff, n):
        super().__init__()
        self.layers = nn.ModuleList([_EncLayer(d, nhead, dff) for _ in range(n)])
        self.norm = nn.LayerNorm(d)


class _EncoderOnly(nn.Module):  # fai_mf/modelling.py:130
    def __init__(self, d, nhead, dff, n):
        super().__init__()
        self.encoder = _Encoder(d, nhead, dff, n)


class TransformerFPN(nn.Module):  # fai_mf/modelling.py:201
    def __init__(self, backbone: ResNet, feat_dim, out_dim, layers, nhead, dff):
        super().__init__()
        self.backbone = backbone
        ch = backbone.out_channels  # res2..res5
        self.input_proj = _ConvBN(ch[3], feat_dim, 1, bias=True, norm=False)
        self.transformer = _EncoderOnly(feat_dim, nhead, dff, layers)
        for idx in (1, 2, 3):
            self.add_module(f"adapter_{idx}", _ConvBN(ch[idx - 1], feat_dim, 1))
            self.add_module(f"layer_{idx}", _ConvBN(feat_dim, feat_dim, 3))
        self.layer_4 = _ConvBN(feat_dim, feat_dim, 3)
        self.mask_features = _ConvBN(feat_dim, out_dim, 3, bias=True, norm=False)


class _AttnLayer(nn.Module):
    def __init__(self, d, nhead, name):
        super().__init__()
        setattr(self, name, nn.MultiheadAttention(d, nhead, dropout=0.0))
        self.norm = nn.LayerNorm(d)


class _FFNLayer(nn.Module):
    def __init__(self, d, dff):
        super().__init__()
        self.linear1, self.linear2, self.norm = nn.Linear(d, dff), nn.Linear(dff, d), nn.LayerNorm(d)


class PredictionHeads(nn.Module):  # fai_mf/modelling.py:28
    def __init__(self, d, num_classes, mask_dim):
        super().__init__()
        self.decoder_norm = nn.LayerNorm(d)
        self.classifier = nn.Linear(d, num_classes + 1)
        self.mask_classifier = MLP(d, d, mask_dim, 3)


class MultiScaleMaskedTransformerDecoder(nn.Module):  # fai_mf/modelling.py:372
    def __init__(self, in_ch, out_dim, num_classes, d, num_queries, nhead, dff, layers):
        super().__init__()
        self.transformer_self_attention_layers = nn.ModuleList([_AttnLayer(d, nhead, "self_attn") for _ in range(layers)])
        self.transformer_cross_attention_layers = nn.ModuleList([_AttnLayer(d, nhead, "multihead_attn") for _ in range(layers)])
        self.transformer_ffn_layers = nn.ModuleList([_FFNLayer(d, dff) for _ in range(layers)])
        self.query_feat, self.query_embed = nn.Embedding(num_queries, d), nn.Embedding(num_queries, d)
        self.input_proj = nn.ModuleList([_ConvBN(in_ch, d, 1, bias=True, norm=False) for _ in range(3)])
        self.forward_prediction_heads = PredictionHeads(d, num_classes, out_dim)


class MaskFormerHead(nn.Module):  # fai_mf/modelling.py:563
    def __init__(self, predictor, num_classes):
        super().__init__()
        self.criterion = _CriterionStub(num_classes)
        self.predictor = predictor


def position_embedding_sine_normalized(h, w, num_pos_feats=128, temperature=10000.0):
    """nn/layers/position_encoding.py:45-74, normalize=True -> [h*w, 2*num_pos_feats] (token-major for the NHWC graph)."""
    y = torch.arange(1, h + 1, dtype=torch.float32).view(h, 1).expand(h, w)
    x = torch.arange(1, w + 1, dtype=torch.float32).view(1, w).expand(h, w)
    eps, scale = 1e-6, 2 * math.pi
    y = y / (float(h) + eps) * scale
    x = x / (float(w) + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    px, py = x[:, :, None] / dim_t, y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).view(h, w, -1)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).view(h, w, -1)
    return torch.cat((py, px), dim=2).reshape(h * w, -1)


class MFEngine(DetrEngine):
    """Packs a FAIMaskFormer state_dict and runs the fused forward (reuses DetrEngine's packing helpers and backbone)."""

    lazy_masks = False  # True: return LazyMasks instead of the materialised [B,Q,H,W] probabilities

    def __init__(self, sd: Dict[str, torch.Tensor], cfg: MaskFormerConfig, device, precision: str = "fp16", algo: int = ops.ALGO_AUTO):
        self.cfg, self.device, self.precision, self.algo = cfg, torch.device(device), precision, algo
        assert precision in ("fp32", "fp16", "fp32_tc")
        self.dt = torch.float16 if precision == "fp16" else torch.float32
        self._host_w3 = {} if precision == "fp32_tc" else None  # fp32 storage, three fp16 tensor-core products per conv / linear (fai_detr._split3_weights)
        self.depth = cfg.backbone_config.depth
        self.nhead, self.d = 8, cfg.transformer_predictor_hidden_dim
        self._consts = {}
        # pack on the HOST (BN folding, re-parameterisation, concatenations are a few hundred tiny tensor ops: as device launches they were ~700 `at::`
        # kernels in front of the first forward); only the packed tensors travel to the device
        sd = {k: v.detach().to("cpu") for k, v in sd.items()}
        self._pack_backbone(sd)
        pd = "pixel_decoder"
        self.pd_in = self._conv_bias(sd, pd + ".input_proj", 0)
        self.enc = [self._pack_attn_block(sd, f"{pd}.transformer.encoder.layers.{i}", ffn_norms=("norm1", "norm2")) for i in range(cfg.pixel_decoder_transformer_layers)]
        self.enc_norm = (self._f32(sd[pd + ".transformer.encoder.norm.weight"]), self._f32(sd[pd + ".transformer.encoder.norm.bias"]))
        self.layer = {i: self._conv_bn(sd, f"{pd}.layer_{i}", 1, ops.ACT_RELU) for i in (1, 2, 3, 4)}
        self.adapter = {i: self._conv_bn(sd, f"{pd}.adapter_{i}", 0, ops.ACT_NONE) for i in (1, 2, 3)}
        self.mask_features = self._conv_bias(sd, pd + ".mask_features", 1)
        self._pack_decoder(sd, 3)
        self._finish_pack()

    def _finish_pack(self):
        if self.precision == "fp32_tc":
            _enable_split3(vars(self), host_w3=self._host_w3)
        self._host_w3 = None

    def _pack_decoder(self, sd, num_levels):
        """head.predictor.* of the masked transformer decoder (same key names in fai_mf and bisenetformer)."""
        cfg = self.cfg
        hp = "head.predictor"
        self.dec_in = [self._conv_bias(sd, f"{hp}.input_proj.{i}", 0) for i in range(num_levels)]
        self.query_feat, self.query_embed = self._to(sd[hp + ".query_feat.weight"].float()), self._to(sd[hp + ".query_embed.weight"].float())
        d = self.d
        self.dec = []
        for i in range(cfg.transformer_predictor_dec_layers):
            c, s, f = (f"{hp}.transformer_cross_attention_layers.{i}", f"{hp}.transformer_self_attention_layers.{i}", f"{hp}.transformer_ffn_layers.{i}")
            wc, bc = sd[c + ".multihead_attn.in_proj_weight"].float(), sd[c + ".multihead_attn.in_proj_bias"].float()
            ws, bs = sd[s + ".self_attn.in_proj_weight"].float(), sd[s + ".self_attn.in_proj_bias"].float()
            self.dec.append({
                "cq": _Linear(self._to(wc[:d]), self._f32(bc[:d])), "ck": _Linear(self._to(wc[d:2 * d]), self._f32(bc[d:2 * d])),
                "cv": _Linear(self._to(wc[2 * d:]), self._f32(bc[2 * d:])), "cout": self._lin(sd, c + ".multihead_attn.out_proj"),
                "cn": (self._f32(sd[c + ".norm.weight"]), self._f32(sd[c + ".norm.bias"])),
                "sqk": _Linear(self._to(ws[:2 * d]), self._f32(bs[:2 * d])), "sv": _Linear(self._to(ws[2 * d:]), self._f32(bs[2 * d:])),
                "sout": self._lin(sd, s + ".self_attn.out_proj"), "sn": (self._f32(sd[s + ".norm.weight"]), self._f32(sd[s + ".norm.bias"])),
                "l1": self._lin(sd, f + ".linear1"), "l2": self._lin(sd, f + ".linear2"), "fn": (self._f32(sd[f + ".norm.weight"]), self._f32(sd[f + ".norm.bias"])),
            })
        h = hp + ".forward_prediction_heads"
        self.head_norm = (self._f32(sd[h + ".decoder_norm.weight"]), self._f32(sd[h + ".decoder_norm.bias"]))
        self.classifier = self._lin(sd, h + ".classifier")
        self.mask_mlp = [self._lin(sd, f"{h}.mask_classifier.layers.{j}") for j in range(3)]

    def _conv_bias(self, sd, p, pad):
        return _Conv(self._to(sd[p + ".weight"].float().permute(0, 2, 3, 1)), None, self._f32(sd[p + ".bias"]), 1, pad, ops.ACT_NONE)

    def _conv_bn(self, sd, p, pad, act):
        s, b = _bn_fold(sd, p + ".norm")
        return _Conv(self._to(sd[p + ".weight"].float().permute(0, 2, 3, 1)), self._f32(s), self._f32(b), 1, pad, act)

    def _pos(self, h, w):
        key = ("pos", h, w)
        if key not in self._consts:
            self._consts[key] = self._to(position_embedding_sine_normalized(h, w, self.d // 2))
        return self._consts[key]

    def _heads(self, out, mask_features, size, want_class):
        """PredictionHeads.forward (:69-112) -> (class logits fp32 or None, mask logits NHWC [B,h4,w4,Qp], (mask, allowed) or None)."""
        A, dt = self.algo, self.dt
        B, Q, d = out.shape
        if getattr(self, "_fused_glue", False):  # fp32_tc: LayerNorm writes the pair operand of the mask MLP, whose hidden layers stay in the pair format
            dn, dnp, _ = ops.layernorm_ex(out, *self.head_norm, want_f32=want_class)
            cls = self.classifier(dn, out_dtype=torch.float32, algo=ops.ALGO_SIMT) if want_class else None
            me = self._plin(self.mask_mlp[2], self._plin(self.mask_mlp[1], self._plin(self.mask_mlp[0], dnp, act=ops.ACT_RELU, out_pair=True), act=ops.ACT_RELU, out_pair=True))
        else:
            dn = ops.layernorm(out, *self.head_norm)
            cls = self.classifier(dn, out_dtype=torch.float32, algo=ops.ALGO_SIMT) if want_class else None
            me = self.mask_mlp[2](self.mask_mlp[1](self.mask_mlp[0](dn, act=ops.ACT_RELU, algo=A), act=ops.ACT_RELU, algo=A), algo=A)  # [B,Q,256]
        _, h4, w4, C = mask_features.shape
        Qp = (Q + 7) // 8 * 8
        masks = torch.zeros((B, h4, w4, Qp), dtype=dt, device=out.device)
        # einsum("bqc,bchw->bqhw"): a [h4*w4, C] x [C, Q] GEMM per image whose "weights" (the mask embeddings) differ per image - ONE launch
        mfp = getattr(self, "_mf_pair", None)
        if mfp is not None and mfp[0] is mask_features:
            # fp32_tc: the same GEMM as three fp16 tensor-core products - mask_features split ONCE per forward (_run_decoder), the per-image embeddings as
            # [W_hi | W_lo | W_hi] triples.  (On the CUDA-core fp32 kernel this product was a third of the parity-mode step: 16.9 of 50.4 ms at bs=16 800x800.)
            ops.conv2d_per_image(mfp[1], _split3_weights(me).reshape(B, Q, 1, 1, 3 * C), out=masks[..., :Q], algo=ops.ALGO_TCGEN05_SPLIT3)
        else:
            ops.conv2d_per_image(mask_features, me.reshape(B, Q, 1, 1, C), out=masks[..., :Q], algo=A)
        attn = None
        if size is not None:
            low = masks if (h4, w4) == tuple(size) else ops.resize_bilinear(masks, size)
            attn = ops.attn_mask_build(low, Q)
        return cls, masks, attn

    @torch.no_grad()
    def forward(self, images: torch.Tensor, taps: Optional[dict] = None):
        cfg, dt, A = self.cfg, self.dt, self.algo
        if images.dtype == torch.uint8:
            B, H, W, _ = images.shape
        else:
            assert images.dim() == 4 and images.shape[1] == 3 and images.dtype == torch.float32
            B, _, H, W = images.shape
        if H % 32 or W % 32:
            # the reference's torch graph takes any size (odd feature maps from ceil-mode pools / stride-2 convs); the B200 kernels tile the stride-2 layers on even
            # maps, so the engine takes multiples of 32 - resize or pad in the processor (image_size) for other inputs
            raise ValueError(f"focoos_b200: input size {H}x{W} is not a multiple of 32; resize/pad the image (e.g. ModelInfo.im_size) before the model")
        pair = self.pair_capable() and all(getattr(c, "w3", None) is not None for c in [self.pd_in] + [self.adapter[i] for i in (1, 2, 3)])
        if pair:
            # fp32_tc: the backbone keeps its activations as fp16 [hi | lo] planes between convs (no split pass in front of every conv, DetrEngine._run_backbone_pair);
            # the four pixel-decoder convs that consume res2..res5 read the pairs and write the fp32 tensors the transformer / FPN arithmetic below works on
            res2, res3, res4, res5 = self._run_backbone_pair(images)
            in_conv = lambda conv, f: self._pc(conv, f, out_pair=False)  # noqa: E731
        else:
            res2, res3, res4, res5 = self._run_backbone(images)
            in_conv = lambda conv, f: conv(f, algo=A)  # noqa: E731
        d, nh = self.d, self.nhead
        scale = 1.0 / math.sqrt(d // nh)
        # ---- pixel decoder (TransformerFPN.forward_features)
        x = in_conv(self.pd_in, res5)
        h, w = x.shape[1], x.shape[2]
        pos = self._pos(h, w)
        src = x.reshape(B, h * w, d)
        for blk in self.enc:  # pre-norm encoder layer (nn/layers/transformer.py:583-601 with normalize_before)
            s2 = ops.layernorm(src, *blk["n_attn"])
            qk = blk["qk"](ops.add(s2, pos), algo=A)
            a = ops.attention(qk[..., :d], qk[..., d:], blk["v"](s2, algo=A), nh, scale, split=self.precision == "fp32_tc")
            src = blk["out"](a, residual=src, algo=A)
            s2 = ops.layernorm(src, *blk["n_ffn"])
            src = blk["l2"](blk["l1"](s2, act=ops.ACT_RELU, algo=A), residual=src, algo=A)
        src = ops.layernorm(src, *self.enc_norm)
        y = self.layer[4](src.reshape(B, h, w, d), algo=A)
        ms = [y]
        # fp32_tc: the 1/4-resolution layer feeds only the mask_features conv, whose output is only ever read as a tensor-core operand (the per-image mask product):
        # both stay in the pair format - no fp32 copy of the two largest activations of the pixel decoder, no split pass over them
        chain = pair and getattr(self.layer[1], "w3", None) is not None and getattr(self.mask_features, "w3", None) is not None
        for idx, f in ((3, res4), (2, res3), (1, res2)):
            u = ops.upsample_nearest_add(y, in_conv(self.adapter[idx], f))
            y = self._pc(self.layer[idx], u) if (chain and idx == 1) else self.layer[idx](u, algo=A)
            if len(ms) < 3:
                ms.append(y)
        mask_features = self._pc(self.mask_features, y) if chain else self.mask_features(y, algo=A)
        if taps is not None:
            taps.update(res5=res5.float() if pair else res5, enc_memory=src.reshape(B, h, w, d), mask_features=mask_features.float() if chain else mask_features, multi_scale=ms)
        return self._run_decoder(ms, mask_features, B, H, W, taps)

    def _run_decoder(self, ms, mask_features, B, H, W, taps=None):
        """MultiScaleMaskedTransformerDecoder.forward (fai_mf/modelling.py:467-550) + head + final upsample; `ms` = the decoder's
        feature levels (3 for fai_mf, 2 for bisenetformer)."""
        cfg, A = self.cfg, self.algo
        d, nh = self.d, self.nhead
        scale = 1.0 / math.sqrt(d // nh)
        nl = len(ms)
        self._mf_pair = None
        if isinstance(mask_features, ops.Pair):  # written as a pair by its conv (MFEngine.forward)
            self._mf_pair = (mask_features, mask_features.buf)
        elif (self.precision == "fp32_tc" and A == ops.ALGO_AUTO and mask_features.dtype == torch.float32 and mask_features.shape[-1] % 64 == 0
                and (ops._backend is not None or ops.supports_tcgen05_cached())):
            self._mf_pair = (mask_features, ops.split_pair(mask_features))  # consumed by every _heads call of this forward
        srcs, kpos, sizes = [], [], []
        for i in range(nl):
            hh, ww = ms[i].shape[1], ms[i].shape[2]
            s = self.dec_in[i](ms[i], algo=A).reshape(B, hh * ww, d)
            srcs.append(s)
            kpos.append(ops.add(s, self._pos(hh, ww)))
            sizes.append((hh, ww))
        # fp32_tc: masked cross-attention on the tensor cores (fb200_attention_masked_split); the per-level key / value inputs are split ONCE (they are the same for the
        # three layers of a level) and the K / V projections run pair -> pair
        split_attn = self.precision == "fp32_tc" and A == ops.ALGO_AUTO
        pair_kv = split_attn and self.pair_capable() and all(b_["ck"].w3 is not None and b_["cv"].w3 is not None for b_ in self.dec)
        if pair_kv:
            kpos_p, srcs_p = [ops.to_pair(t) for t in kpos], [ops.to_pair(t) for t in srcs]
        Q = cfg.num_queries
        out = self.query_feat.unsqueeze(0).expand(B, Q, d).contiguous()
        qpos = self.query_embed
        # fused row glue (csrc/head_fused.cu, the kernels of the fai-detr head): every LayerNorm writes the pair operand(s) of the linears behind it - LN(x) and
        # LN(x) + query_pos in one launch - and the FFN / mask-MLP hidden layers stay in the pair format: no add / split launches between two tensor-core linears
        lins = [b_[k] for b_ in self.dec for k in ("cq", "cout", "sqk", "sv", "sout", "l1", "l2")] + list(self.mask_mlp)
        self._fused_glue = bool(pair_kv and self.fused_glue and all(getattr(l_, "w3", None) is not None for l_ in lins))
        _, masks, attn = self._heads(out, mask_features, sizes[0], False)
        L = len(self.dec)
        cls = None
        for i, blk in enumerate(self.dec):
            lvl = i % nl
            if self._fused_glue:
                _, _, tq = ops.layernorm_ex(out, *blk["cn"], pos=qpos, want_f32=False, want_pair=False, want_pair_pos=True)
                q = self._plin(blk["cq"], tq)
                kk, vv = self._plin(blk["ck"], kpos_p[lvl], out_pair=True), self._plin(blk["cv"], srcs_p[lvl], out_pair=True)
                a = ops.attention_masked(q, kk, vv, attn[0], attn[1], nh, scale, split=True)
                out = self._plin(blk["cout"], ops.to_pair(a), residual=out)
                _, t2p, t2pp = ops.layernorm_ex(out, *blk["sn"], pos=qpos, want_f32=False, want_pair=True, want_pair_pos=True)
                qk = self._plin(blk["sqk"], t2pp)
                a = ops.attention(qk[..., :d], qk[..., d:], self._plin(blk["sv"], t2p), nh, scale, split=True, out_pair=True)
                out = self._plin(blk["sout"], a, residual=out)
                _, t2p, _ = ops.layernorm_ex(out, *blk["fn"], want_f32=False)
                out = self._plin(blk["l2"], self._plin(blk["l1"], t2p, act=ops.ACT_RELU, out_pair=True), residual=out)
                last = i == L - 1
                cls, masks, attn = self._heads(out, mask_features, None if last else sizes[(i + 1) % nl], last)
                if taps is not None:
                    taps[f"dec{i}_out"] = out
                continue
            t2 = ops.layernorm(out, *blk["cn"])
            q = blk["cq"](ops.add(t2, qpos), algo=A)
            if pair_kv:  # K / V projections write the fp16 [hi | lo] pairs the attention kernel stages with plain 16-byte copies
                kk, vv = self._plin(blk["ck"], kpos_p[lvl], out_pair=True), self._plin(blk["cv"], srcs_p[lvl], out_pair=True)
            else:
                kk, vv = blk["ck"](kpos[lvl], algo=A), blk["cv"](srcs[lvl], algo=A)
            a = ops.attention_masked(q, kk, vv, attn[0], attn[1], nh, scale, split=split_attn)
            out = blk["cout"](a, residual=out, algo=A)
            t2 = ops.layernorm(out, *blk["sn"])
            qk = blk["sqk"](ops.add(t2, qpos), algo=A)
            a = ops.attention(qk[..., :d], qk[..., d:], blk["sv"](t2, algo=A), nh, scale, split=self.precision == "fp32_tc")
            out = blk["sout"](a, residual=out, algo=A)
            t2 = ops.layernorm(out, *blk["fn"])
            out = blk["l2"](blk["l1"](t2, act=ops.ACT_RELU, algo=A), residual=out, algo=A)
            last = i == L - 1
            cls, masks, attn = self._heads(out, mask_features, None if last else sizes[(i + 1) % nl], last)
            if taps is not None:
                taps[f"dec{i}_out"] = out
        if taps is not None:
            taps.update(pred_logits=cls, pred_masks=masks)  # masks: NHWC [B,h4,w4,Qp] pre-sigmoid logits
        probs = ops.softmax_drop_last(cls)
        lazy = LazyMasks(masks, Q, (H, W))
        self._mf_pair, self._fused_glue = None, False
        return probs, (lazy if self.lazy_masks else lazy.materialize())


class FAIMaskFormer(nn.Module):
    """Drop-in for the reference `FAIMaskFormer(BaseModelNN)` (fai_mf/modelling.py:633)."""

    lazy_masks = False  # True: forward() returns fai_mf.LazyMasks (low-resolution logits) instead of the upsampled [B,Q,H,W] probabilities

    def __init__(self, config: MaskFormerConfig, precision: str = "fp16"):
        super().__init__()
        self.config = c = config
        if c.postprocessing_type not in ("semantic", "instance"):
            raise ValueError(f"Invalid postprocessing type: {c.postprocessing_type}. Must be one of: ['semantic', 'instance']")
        self.pixel_decoder = TransformerFPN(ResNet(c.backbone_config), c.pixel_decoder_feat_dim, c.pixel_decoder_out_dim, c.pixel_decoder_transformer_layers,
                                            c.pixel_decoder_transformer_nheads, c.pixel_decoder_transformer_dim_feedforward)
        self.head = MaskFormerHead(MultiScaleMaskedTransformerDecoder(c.pixel_decoder_out_dim, c.transformer_predictor_out_dim, c.num_classes,
                                                                      c.transformer_predictor_hidden_dim, c.num_queries, 8,
                                                                      c.transformer_predictor_dim_feedforward, c.transformer_predictor_dec_layers), c.num_classes)
        self.register_buffer("pixel_mean", torch.tensor(c.pixel_mean, dtype=torch.float32).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor(c.pixel_std, dtype=torch.float32).view(-1, 1, 1), False)
        self.num_classes, self.precision, self.algo, self._engine = c.num_classes, precision, ops.ALGO_AUTO, None
        self.eval()

    device = property(lambda self: self.pixel_mean.device)
    dtype = property(lambda self: self.pixel_mean.dtype)

    def load_state_dict(self, state_dict, strict: bool = False, assign: bool = False):
        if "model" in state_dict and isinstance(state_dict["model"], dict):
            state_dict = state_dict["model"]
        own = self.state_dict()
        filtered = {k: v for k, v in state_dict.items() if k in own and tuple(own[k].shape) == tuple(v.shape)}
        res = super().load_state_dict(filtered, strict=False)
        self._engine = None
        if strict and (res.missing_keys or len(filtered) != len(state_dict)):
            raise RuntimeError(f"load_state_dict(strict): missing {res.missing_keys[:5]} / dropped {len(state_dict) - len(filtered)}")
        return res

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self) -> MFEngine:
        e = self._engine
        if e is None or e.device != self.device or e.precision != self.precision or e.algo != self.algo:
            self._engine = MFEngine(self.state_dict(), self.config, self.device, self.precision, self.algo)
        return self._engine

    def forward(self, images: torch.Tensor, targets: list = [], taps: Optional[dict] = None) -> MaskFormerModelOutput:
        if self.training or (targets is not None and len(targets) > 0):
            raise NotImplementedError("focoos_b200: losses / fine-tuning are not part of the inference hot path")
        if ops._backend is None and not images.is_cuda:
            raise RuntimeError("focoos_b200.FAIMaskFormer runs on CUDA (sm_100a) only — no CPU fallback")
        eng = self.engine()
        eng.lazy_masks = bool(getattr(self, "lazy_masks", False))
        probs, masks = eng.forward(images if images.dtype == torch.uint8 else images.to(torch.float32), taps)
        return MaskFormerModelOutput(masks=masks, logits=probs, loss=None)
