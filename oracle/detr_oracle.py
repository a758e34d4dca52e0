"""CPU oracle: functional fp32 restatement of the reference's FAIDetr inference path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import this module.  The product path
(`focoos_b200`) never routes through it and fails loudly when its CUDA library is missing.

It restates, function by function, what `/root/reference/focoos` computes for
`ModelManager.get("fai-detr-*") -> model.forward -> processor.postprocess`, operating directly on
a reference-keyed `state_dict` in the reference's own NCHW fp32
formulation (unfused BatchNorm, `F.grid_sample` deformable attention, `nn.MultiheadAttention`
maths), so it is independent of the fused NHWC formulation the CUDA path uses.

Pinned: `tests/test_oracle.py` checks it (a) against outputs of the UNMODIFIED reference imported
in the build container (`oracle/ref_import.py`, skipped where /root/reference is absent) and
(b) against the committed fixtures in `tests/golden/` that `oracle/gen_golden.py` produced from
the reference.  The reference's own test-suite holds no numeric fixture for this path
 — "parity unpinned" by the reference's suite, pinned here by side-by-side runs.

Each function cites the reference file:line it follows (paths relative to /root/reference/focoos).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass
class DetrOracleConfig:
    """Subset of DETRConfig (models/fai_detr/config.py:9-61) + ResnetConfig (nn/backbone/resnet.py:152-161)."""

    num_classes: int = 365
    num_queries: int = 300
    depth: int = 50
    feat_dim: int = 256
    hidden_dim: int = 256
    nhead: int = 8
    enc_dim_feedforward: int = 1024
    dec_dim_feedforward: int = 1024
    dec_layers: int = 6
    num_points: int = 4
    pixel_mean: Sequence[float] = (123.675, 116.28, 103.53)
    pixel_std: Sequence[float] = (58.395, 57.12, 57.375)
    threshold: float = 0.5
    top_k: int = 300
    compute_dead_mask_features: bool = False  # modelling.py:347 computes it, :381 drops it


RESNET_BLOCKS = {50: [3, 4, 6, 3], 101: [3, 4, 23, 3]}  # nn/backbone/resnet.py:19-25


# ----------------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------------
def batchnorm_eval(x: Tensor, sd: SD, p: str) -> Tensor:
    """nn.BatchNorm2d in eval, eps 1e-5 (nn/layers/conv.py:89, nn/layers/norm.py get_norm('BN'))."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def act_fn(x: Tensor, act: Optional[str]) -> Tensor:
    """nn/layers/base.py:8-28."""
    if act is None:
        return x
    if act == "relu":
        return F.relu(x)
    if act == "silu":
        return F.silu(x)
    if act == "gelu":
        return F.gelu(x)
    raise ValueError(act)


def conv_norm_layer(x: Tensor, sd: SD, p: str, stride: int = 1, act: Optional[str] = None) -> Tensor:
    """ConvNormLayer.forward (nn/layers/conv.py:78-98): conv(no bias, pad=(k-1)//2) -> BN -> act."""
    w = sd[p + ".conv.weight"]
    k = w.shape[-1]
    x = F.conv2d(x, w, None, stride, (k - 1) // 2)
    x = batchnorm_eval(x, sd, p + ".norm")
    return act_fn(x, act)


def linear(x: Tensor, sd: SD, p: str) -> Tensor:
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def layer_norm(x: Tensor, sd: SD, p: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def mlp(x: Tensor, sd: SD, p: str, num_layers: int) -> Tensor:
    """MLP.forward (nn/layers/base.py:51-62): Linear->ReLU ... ->Linear."""
    for i in range(num_layers):
        x = linear(x, sd, f"{p}.layers.{i}")
        if i < num_layers - 1:
            x = F.relu(x)
    return x


def multihead_attention(q_in: Tensor, k_in: Tensor, v_in: Tensor, sd: SD, p: str, nhead: int) -> Tensor:
    """nn.MultiheadAttention(batch_first=True) maths: packed in_proj [3d,d],
    scale 1/sqrt(d/heads), softmax over keys, out_proj.  Inputs [B,L,d]."""
    d = q_in.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(q_in, w[:d], b[:d])
    k = F.linear(k_in, w[d : 2 * d], b[d : 2 * d])
    v = F.linear(v_in, w[2 * d :], b[2 * d :])
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    hd = d // nhead
    q = q.view(B, Lq, nhead, hd).transpose(1, 2)
    k = k.view(B, Lk, nhead, hd).transpose(1, 2)
    v = v.view(B, Lk, nhead, hd).transpose(1, 2)
    attn = torch.softmax((q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd)), dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B, Lq, d)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


# ----------------------------------------------------------------------------------------------
# backbone: ResNet-vd (nn/backbone/resnet.py)
# ----------------------------------------------------------------------------------------------
def bottleneck(x: Tensor, sd: SD, p: str, stride: int, shortcut: bool) -> Tensor:
    """BottleNeck.forward variant 'd' (nn/backbone/resnet.py:72-121): stride on branch2b (:81);
    stride-2 shortcut = AvgPool2d(2,2,0,ceil) + 1x1 ConvNormLayer (:91-102); add then ReLU (:118-119)."""
    out = conv_norm_layer(x, sd, p + ".branch2a", 1, "relu")
    out = conv_norm_layer(out, sd, p + ".branch2b", stride, "relu")
    out = conv_norm_layer(out, sd, p + ".branch2c", 1, None)
    if shortcut:
        short = x
    elif stride == 2:
        short = F.avg_pool2d(x, 2, 2, 0, ceil_mode=True)
        short = conv_norm_layer(short, sd, p + ".short.conv", 1, None)
    else:
        short = conv_norm_layer(x, sd, p + ".short", stride, None)
    return F.relu(out + short)


def resnet_vd(x: Tensor, sd: SD, p: str, depth: int) -> Dict[str, Tensor]:
    """ResNet.forward (nn/backbone/resnet.py:252-266): 3-conv stem (:181-186), max_pool2d(3,2,1) (:254),
    stages; Blocks stride rule `2 if i == 0 and stage_num != 2 else 1` (:133)."""
    x = conv_norm_layer(x, sd, p + ".conv1.conv1_1", 2, "relu")
    x = conv_norm_layer(x, sd, p + ".conv1.conv1_2", 1, "relu")
    x = conv_norm_layer(x, sd, p + ".conv1.conv1_3", 1, "relu")
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = {}
    for si, count in enumerate(RESNET_BLOCKS[depth]):
        for bi in range(count):
            stride = 2 if (bi == 0 and si != 0) else 1
            x = bottleneck(x, sd, f"{p}.res_layers.{si}.blocks.{bi}", stride, shortcut=(bi != 0))
        outs[f"res{si + 2}"] = x
    return outs


# ----------------------------------------------------------------------------------------------
# hybrid encoder (models/fai_detr/modelling.py:195-347)
# ----------------------------------------------------------------------------------------------
def aifi_position_embedding(h: int, w: int, num_pos_feats: int = 128, temperature: float = 10000.0) -> Tensor:
    """File-local PositionEmbeddingSine, normalize=False (modelling.py:110-179): 0-based cumsum coords
    (:161-162), dim_t (:167-168), cat(y_sin, y_cos, x_sin, x_cos) (:174-178) -> [1, h*w, 4*num_pos_feats/2]."""
    not_mask = torch.ones(1, h, w, dtype=torch.bool)
    y_embed = not_mask.cumsum(1, dtype=torch.float32) - 1
    x_embed = not_mask.cumsum(2, dtype=torch.float32) - 1
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x_sin = pos_x[:, :, :, 0::2].sin().view(1, h * w, -1)
    pos_x_cos = pos_x[:, :, :, 1::2].cos().view(1, h * w, -1)
    pos_y_sin = pos_y[:, :, :, 0::2].sin().view(1, h * w, -1)
    pos_y_cos = pos_y[:, :, :, 1::2].cos().view(1, h * w, -1)
    return torch.cat((pos_y_sin, pos_y_cos, pos_x_sin, pos_x_cos), dim=2)


def aifi_layer(src: Tensor, pos: Tensor, sd: SD, p: str, nhead: int) -> Tensor:
    """TransformerEncoderLayer.forward post-norm (nn/layers/transformer.py:583-601), GELU FFN."""
    q = k = src + pos
    a = multihead_attention(q, k, src, sd, p + ".self_attn", nhead)
    src = layer_norm(src + a, sd, p + ".norm1")
    f = linear(F.gelu(linear(src, sd, p + ".linear1")), sd, p + ".linear2")
    return layer_norm(src + f, sd, p + ".norm2")


def repvgg_block(x: Tensor, sd: SD, p: str) -> Tensor:
    """RepVggBlock.forward, UNFUSED (modelling.py:39-45)."""
    y = conv_norm_layer(x, sd, p + ".conv1", 1, None) + conv_norm_layer(x, sd, p + ".conv2", 1, None)
    return F.silu(y)


def csp_rep_layer(x: Tensor, sd: SD, p: str, num_blocks: int = 3) -> Tensor:
    """CSPRepLayer.forward (modelling.py:103-107); conv3 = Identity for expansion 1.0 (:98-101)."""
    x1 = conv_norm_layer(x, sd, p + ".conv1", 1, "silu")
    for i in range(num_blocks):
        x1 = repvgg_block(x1, sd, f"{p}.bottlenecks.{i}")
    x2 = conv_norm_layer(x, sd, p + ".conv2", 1, "silu")
    y = x1 + x2
    if (p + ".conv3.conv.weight") in sd:
        y = conv_norm_layer(y, sd, p + ".conv3", 1, "silu")
    return y


def hybrid_encoder(images_norm: Tensor, sd: SD, cfg: DetrOracleConfig, taps: Optional[dict] = None) -> List[Tensor]:
    """Encoder.forward (modelling.py:297-347). Returns outs[::-1] = [1/32, 1/16, 1/8] maps (:347)."""
    p = "pixel_decoder"
    feats_d = resnet_vd(images_norm, sd, p + ".backbone", cfg.depth)
    feats = [feats_d["res3"], feats_d["res4"], feats_d["res5"]]
    if taps is not None:
        taps.update({k: v for k, v in feats_d.items()})
    proj = []
    for i, f in enumerate(feats):  # input_proj: Conv2d 1x1 no bias + BN (:230-237, :302)
        y = F.conv2d(f, sd[f"{p}.input_proj.{i}.0.weight"])
        proj.append(batchnorm_eval(y, sd, f"{p}.input_proj.{i}.1"))
    # AIFI on the 1/32 map (:315-324)
    B, C, h, w = proj[2].shape
    src = proj[2].flatten(2).permute(0, 2, 1)
    pos = aifi_position_embedding(h, w, cfg.feat_dim // 2)
    mem = aifi_layer(src, pos, sd, f"{p}.encoder.0.layers.0", cfg.nhead)
    proj[2] = mem.permute(0, 2, 1).reshape(B, C, h, w).contiguous()
    if taps is not None:
        taps["aifi"] = proj[2]
    # top-down FPN (:328-336)
    inner = [proj[2]]
    for idx in (2, 1):
        hi = conv_norm_layer(inner[0], sd, f"{p}.lateral_convs.{2 - idx}", 1, "silu")
        inner[0] = hi
        lo = proj[idx - 1]
        up = F.interpolate(hi, size=lo.shape[-2:], mode="bilinear")
        inner.insert(0, csp_rep_layer(torch.cat([up, lo], 1), sd, f"{p}.fpn_blocks.{2 - idx}"))
    # bottom-up PAN (:338-345): bilinear resize THEN stride-1 3x3 conv
    outs = [inner[0]]
    for idx in range(2):
        hi = inner[idx + 1]
        down = F.interpolate(outs[-1], size=hi.shape[-2:], mode="bilinear")
        down = conv_norm_layer(down, sd, f"{p}.downsample_convs.{idx}", 1, "silu")
        outs.append(csp_rep_layer(torch.cat([down, hi], 1), sd, f"{p}.pan_blocks.{idx}"))
    if cfg.compute_dead_mask_features and taps is not None:  # (:347) result unused by DETRHead (:381)
        taps["mask_features"] = F.conv2d(outs[0], sd[p + ".mask_features.weight"], sd[p + ".mask_features.bias"], 1, 1)
    if taps is not None:
        taps["enc_outs"] = outs[::-1]
    return outs[::-1]


# ----------------------------------------------------------------------------------------------
# transformer predictor (models/fai_detr/modelling.py:1023-1263)
# ----------------------------------------------------------------------------------------------
def generate_anchors(spatial_shapes: Sequence[Tuple[int, int]], grid_size: float = 0.05, eps: float = 1e-2):
    """_generate_anchors (modelling.py:1169-1189). Returns logit-space anchors [1,S,4], valid mask [1,S,1]."""
    anchors = []
    for lvl, (h, w) in enumerate(spatial_shapes):
        gy, gx = torch.meshgrid(torch.arange(end=h, dtype=torch.float32), torch.arange(end=w, dtype=torch.float32), indexing="ij")
        grid_xy = torch.stack([gx, gy], -1)
        valid_wh = torch.tensor([w, h]).to(torch.float32)
        grid_xy = (grid_xy.unsqueeze(0) + 0.5) / valid_wh
        wh = torch.ones_like(grid_xy) * grid_size * (2.0 ** (2 - lvl))
        anchors.append(torch.concat([grid_xy, wh], -1).reshape(-1, h * w, 4))
    anchors = torch.concat(anchors, 1)
    valid = ((anchors > eps) * (anchors < 1 - eps)).all(-1, keepdim=True)
    anchors = torch.log(anchors / (1 - anchors))
    anchors = torch.where(valid, anchors, torch.zeros(()))
    return anchors, valid


def topk_stable(scores: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """torch.topk(sorted=True) with the tie rule made explicit: descending value, ties by ascending
    index (torch leaves tie order unspecified, SURVEY §7 'hard parts'). scores [..., N]."""
    vals, idx = torch.sort(scores, dim=-1, descending=True, stable=True)
    return vals[..., :k], idx[..., :k]


def inverse_sigmoid(x: Tensor, eps: float = 1e-5) -> Tensor:
    """nn/layers/functional.py:4-6."""
    x = x.clip(min=0.0, max=1.0)
    return torch.log(x.clip(min=eps) / (1 - x).clip(min=eps))


def ms_deform_attn_core(value: Tensor, shapes: Sequence[Tuple[int, int]], loc: Tensor, w: Tensor) -> Tensor:
    """ms_deform_attn_core_pytorch (nn/layers/deformable.py:10-35), verbatim semantics:
    grid = 2*loc-1; grid_sample(bilinear, zeros, align_corners=False) per level; weighted sum."""
    N_, S_, M_, D_ = value.shape
    _, Lq_, _, L_, P_, _ = loc.shape
    value_list = value.split([h * w_ for h, w_ in shapes], dim=1)
    grids = 2 * loc - 1
    sampled = []
    for lid, (H_, W_) in enumerate(shapes):
        v = value_list[lid].flatten(2).transpose(1, 2).reshape(N_ * M_, D_, H_, W_)
        g = grids[:, :, :, lid].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    aw = w.transpose(1, 2).reshape(N_ * M_, 1, Lq_, L_ * P_)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1).view(N_, M_ * D_, Lq_)
    return out.transpose(1, 2).contiguous()


def ms_deformable_attention(query: Tensor, ref: Tensor, memory: Tensor, shapes, sd: SD, p: str, nhead: int, npoints: int) -> Tensor:
    """MSDeformableAttention.forward, 4-d reference branch (modelling.py:831-884)."""
    bs, Lq, C = query.shape
    nl = len(shapes)
    value = linear(memory, sd, p + ".value_proj").reshape(bs, memory.shape[1], nhead, C // nhead)
    off = linear(query, sd, p + ".sampling_offsets").reshape(bs, Lq, nhead, nl, npoints, 2)
    aw = linear(query, sd, p + ".attention_weights").reshape(bs, Lq, nhead, nl * npoints)
    aw = F.softmax(aw, dim=-1).reshape(bs, Lq, nhead, nl, npoints)
    ref_in = ref.unsqueeze(2)  # [bs,Lq,1,4] broadcast over levels (modelling.py:989)
    loc = ref_in[:, :, None, :, None, :2] + off / npoints * ref_in[:, :, None, :, None, 2:] * 0.5
    out = ms_deform_attn_core(value, shapes, loc, aw)
    return linear(out, sd, p + ".output_proj")


def decoder_layer(tgt: Tensor, ref: Tensor, memory: Tensor, shapes, pos: Tensor, sd: SD, p: str, cfg: DetrOracleConfig) -> Tensor:
    """TransformerDecoderLayer.forward (modelling.py:924-958)."""
    q = k = tgt + pos
    a = multihead_attention(q, k, tgt, sd, p + ".self_attn", cfg.nhead)
    tgt = layer_norm(tgt + a, sd, p + ".norm1")
    c = ms_deformable_attention(tgt + pos, ref, memory, shapes, sd, p + ".cross_attn", cfg.nhead, cfg.num_points)
    tgt = layer_norm(tgt + c, sd, p + ".norm2")
    f = linear(F.relu(linear(tgt, sd, p + ".linear1")), sd, p + ".linear2")
    return layer_norm(tgt + f, sd, p + ".norm3")


def transformer_predictor(feats: List[Tensor], sd: SD, cfg: DetrOracleConfig, taps: Optional[dict] = None):
    """TransformerPredictor.forward in eval (modelling.py:1234-1263). Returns (pred_logits raw, pred_boxes cxcywh)."""
    p = "head.predictor"
    # _get_encoder_input (:1145-1167)
    flat, shapes = [], []
    for i, f in enumerate(feats):
        y = F.conv2d(f, sd[f"{p}.input_proj.{i}.conv.weight"])
        y = batchnorm_eval(y, sd, f"{p}.input_proj.{i}.norm")
        shapes.append((y.shape[2], y.shape[3]))
        flat.append(y.flatten(2).permute(0, 2, 1))
    memory = torch.concat(flat, 1)
    # _get_decoder_input (:1191-1232)
    anchors, valid = generate_anchors(shapes)
    mem_v = valid.to(memory.dtype) * memory
    output_memory = layer_norm(linear(mem_v, sd, p + ".enc_output.0"), sd, p + ".enc_output.1")
    enc_class = linear(output_memory, sd, p + ".enc_score_classifier")
    enc_coord_unact = mlp(output_memory, sd, p + ".enc_bbox_classifier", 3) + anchors
    scores = enc_class.max(-1).values
    _, topk_ind = topk_stable(scores, cfg.num_queries)
    ref_unact = enc_coord_unact.gather(1, topk_ind.unsqueeze(-1).repeat(1, 1, 4))
    target = output_memory.gather(1, topk_ind.unsqueeze(-1).repeat(1, 1, output_memory.shape[-1]))
    if taps is not None:
        taps.update(memory=memory, enc_scores=scores, topk_ind=topk_ind, target=target, ref_unact=ref_unact)
    # TransformerDecoder.forward eval (:969-1020), eval_idx = last layer
    out = target
    ref = torch.sigmoid(ref_unact)
    logits = boxes = None
    for i in range(cfg.dec_layers):
        pos = mlp(ref, sd, p + ".query_pos_head", 2)
        out = decoder_layer(out, ref, memory, shapes, pos, sd, f"{p}.decoder.layers.{i}", cfg)
        new_ref = torch.sigmoid(mlp(out, sd, f"{p}.dec_bbox_classifier.{i}", 3) + inverse_sigmoid(ref))
        if taps is not None:
            taps[f"dec{i}_out"] = out
            taps[f"dec{i}_ref"] = new_ref
        if i == cfg.dec_layers - 1:
            logits = linear(out, sd, f"{p}.dec_score_classifier.{i}")
            boxes = new_ref
            break
        ref = new_ref
    return logits, boxes


def box_cxcywh_to_xyxy(x: Tensor) -> Tensor:
    """utils/box.py:14-17."""
    xc, yc, w, h = x.unbind(-1)
    return torch.stack([xc - 0.5 * w, yc - 0.5 * h, xc + 0.5 * w, yc + 0.5 * h], dim=-1)


def detr_forward(sd: SD, images: Tensor, cfg: DetrOracleConfig, taps: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """FAIDetr.forward eval (modelling.py:1344-1358) + DETRHead.forward (:386-401).
    images [B,3,H,W] fp32 0..255 -> (scores = sigmoid(logits) [B,Q,C], boxes xyxy [B,Q,4])."""
    mean = torch.tensor(list(cfg.pixel_mean), dtype=torch.float32).view(-1, 1, 1)
    std = torch.tensor(list(cfg.pixel_std), dtype=torch.float32).view(-1, 1, 1)
    x = (images - mean) / std
    feats = hybrid_encoder(x, sd, cfg, taps)
    logits, boxes = transformer_predictor(feats, sd, cfg, taps)
    if taps is not None:
        taps["pred_logits"] = logits
        taps["pred_boxes_cxcywh"] = boxes
    return torch.sigmoid(logits), box_cxcywh_to_xyxy(boxes)


# ----------------------------------------------------------------------------------------------
# processor (models/fai_detr/processor.py, processor/base_processor.py)
# ----------------------------------------------------------------------------------------------
def detr_preprocess(images: Sequence, target_size: Optional[Tuple[int, int]]) -> Tensor:
    """Processor.get_torch_batch (processor/base_processor.py:223-296) for a list of HWC uint8 arrays /
    CHW tensors: ->float32 CHW, optional bilinear resize (align_corners=False) per image, stack."""
    import numpy as np

    outs = []
    for im in images:
        if isinstance(im, np.ndarray):
            im = torch.from_numpy(np.ascontiguousarray(im))
        if im.dim() == 3:
            im = im.unsqueeze(0)
        if im.shape[1] != 3 and im.shape[-1] == 3:
            im = im.permute(0, 3, 1, 2)
        im = im.to(torch.float32)
        if target_size is not None:
            im = F.interpolate(im, size=target_size, mode="bilinear", align_corners=False)
        outs.append(im.squeeze(0))
    return torch.stack(outs, 0)


@dataclass
class OracleDetections:
    boxes: List[List[int]] = field(default_factory=list)  # int xyxy in original-image pixels
    scores: List[float] = field(default_factory=list)
    labels: List[int] = field(default_factory=list)
    query_index: List[int] = field(default_factory=list)


def detr_postprocess(scores: Tensor, boxes: Tensor, image_sizes: Sequence[Tuple[int, int]], threshold: float = 0.5, top_k: int = 300) -> List[OracleDetections]:
    """DETRProcessor.postprocess / _get_predictions (models/fai_detr/processor.py:146-217):
    per image topk over flattened [Q*C] scores, label = i % C, query = i // C, keep score > thr (strict),
    x*W, y*H of the ORIGINAL image, torch.round (half-to-even) -> int32; order = descending score."""
    B, Q, C = scores.shape
    res = []
    for i in range(B):
        s, idx = topk_stable(scores[i].flatten(0), top_k)
        labels = idx % C
        q = idx // C
        b = boxes[i].gather(0, q.unsqueeze(-1).repeat(1, 4))
        m = s > threshold
        b, s, labels, q = b[m].clone(), s[m], labels[m], q[m]
        b[:, 0::2] = b[:, 0::2] * image_sizes[i][1]
        b[:, 1::2] = b[:, 1::2] * image_sizes[i][0]
        b = b.round().to(torch.int32)
        res.append(OracleDetections(b.tolist(), s.tolist(), labels.tolist(), q.tolist()))
    return res
