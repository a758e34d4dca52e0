"""CPU oracle: functional fp32 restatement of the reference's FAIMaskFormer inference path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE (same import rules as oracle/detr_oracle.py).
Restates `focoos/models/fai_mf/modelling.py` (TransformerFPN pixel decoder, MultiScaleMaskedTransformerDecoder,
PredictionHeads, MaskFormerHead, FAIMaskFormer.forward) and the tensor part of `MaskFormerProcessor.postprocess`
on a reference-keyed state_dict, in the reference's own NCHW / [B,Q,H,W] formulation.  Pinned by tests/test_oracle_mf.py
against fixtures produced from the unmodified reference (oracle/gen_golden_mf.py).  Paths cited relative to /root/reference/focoos.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .detr_oracle import batchnorm_eval, layer_norm, linear, mlp, multihead_attention, resnet_vd

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass
class MFOracleConfig:
    """Subset of MaskFormerConfig (models/fai_mf/config.py) for fai-mf-l-coco-ins."""

    num_classes: int = 80
    num_queries: int = 100
    depth: int = 101
    feat_dim: int = 256
    hidden_dim: int = 256
    nhead: int = 8
    enc_layers: int = 6
    dec_layers: int = 9
    pixel_mean: Sequence[float] = (123.675, 116.28, 103.53)
    pixel_std: Sequence[float] = (58.395, 57.12, 57.375)
    mask_threshold: float = 0.5
    threshold: float = 0.5
    use_mask_score: bool = True
    predict_all_pixels: bool = False


def position_embedding_sine_normalized(h: int, w: int, num_pos_feats: int = 128, temperature: float = 10000.0) -> Tensor:
    """nn/layers/position_encoding.py:45-74 with normalize=True: 1-based cumsum, /(last+eps)*2pi, interleaved sin/cos,
    cat(pos_y, pos_x) -> [1, 2*num_pos_feats, h, w]."""
    not_mask = torch.ones(1, h, w, dtype=torch.bool)
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps, scale = 1e-6, 2 * math.pi
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).view(1, h, w, -1)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).view(1, h, w, -1)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def conv2d_norm(x: Tensor, sd: SD, p: str, pad: int, relu: bool) -> Tensor:
    """nn/layers/conv.py:22-75 `Conv2d` wrapper: conv (+bias if present) -> child `norm` (BN) if present -> activation."""
    y = F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), 1, pad)
    if (p + ".norm.weight") in sd:
        y = batchnorm_eval(y, sd, p + ".norm")
    return F.relu(y) if relu else y


def encoder_layer_pre_norm(src: Tensor, pos: Tensor, sd: SD, p: str, nhead: int) -> Tensor:
    """TransformerEncoderLayer.forward with normalize_before=True (nn/layers/transformer.py:583-601), ReLU FFN."""
    s2 = layer_norm(src, sd, p + ".norm1")
    q = k = s2 + pos
    src = src + multihead_attention(q, k, s2, sd, p + ".self_attn", nhead)
    s2 = layer_norm(src, sd, p + ".norm2")
    return src + linear(F.relu(linear(s2, sd, p + ".linear1")), sd, p + ".linear2")


def transformer_fpn(images_norm: Tensor, sd: SD, cfg: MFOracleConfig, taps: Optional[dict] = None):
    """TransformerFPN.forward_features (fai_mf/modelling.py:348-369). Returns (mask_features [B,C,H/4,W/4], [1/32,1/16,1/8] maps)."""
    p = "pixel_decoder"
    feats = resnet_vd(images_norm, sd, p + ".backbone", cfg.depth)
    if taps is not None:
        taps.update(feats)
    x = F.conv2d(feats["res5"], sd[p + ".input_proj.weight"], sd[p + ".input_proj.bias"])
    B, C, h, w = x.shape
    pos = position_embedding_sine_normalized(h, w, cfg.feat_dim // 2).flatten(2).permute(0, 2, 1)  # [1,hw,C]
    src = x.flatten(2).permute(0, 2, 1)
    for i in range(cfg.enc_layers):
        src = encoder_layer_pre_norm(src, pos, sd, f"{p}.transformer.encoder.layers.{i}", cfg.nhead)
    src = layer_norm(src, sd, p + ".transformer.encoder.norm")  # final norm (transformer.py:495-496)
    x = src.permute(0, 2, 1).reshape(B, C, h, w)
    if taps is not None:
        taps["enc_memory"] = x
    y = conv2d_norm(x, sd, p + ".layer_4", 1, True)
    ms = [y]
    for idx, name in ((3, "res4"), (2, "res3"), (1, "res2")):
        cur = conv2d_norm(feats[name], sd, f"{p}.adapter_{idx}", 0, False)
        y = cur + F.interpolate(y, size=cur.shape[-2:], mode="nearest")
        y = conv2d_norm(y, sd, f"{p}.layer_{idx}", 1, True)
        if len(ms) < 3:
            ms.append(y)
    mask_features = F.conv2d(y, sd[p + ".mask_features.weight"], sd[p + ".mask_features.bias"], 1, 1)
    if taps is not None:
        taps["mask_features"] = mask_features
        taps["multi_scale"] = ms
    return mask_features, ms


def prediction_heads(x: Tensor, mask_features: Tensor, sd: SD, p: str, size: Optional[Tuple[int, int]]):
    """PredictionHeads.forward (fai_mf/modelling.py:69-112); x [B,Q,C]."""
    d = layer_norm(x, sd, p + ".decoder_norm")
    cls = linear(d, sd, p + ".classifier")
    me = mlp(d, sd, p + ".mask_classifier", 3)
    masks = torch.einsum("bqc,bchw->bqhw", me, mask_features)
    attn = None
    if size is not None:
        attn = F.interpolate(masks, size=size, mode="bilinear", align_corners=False).flatten(2) < 0  # True = not allowed
    return cls, masks, attn


def masked_decoder(ms: List[Tensor], mask_features: Tensor, sd: SD, cfg: MFOracleConfig, taps: Optional[dict] = None):
    """MultiScaleMaskedTransformerDecoder.forward (fai_mf/modelling.py:467-550), pre-norm layers
    (nn/layers/transformer.py:83-106,206-238,365-378)."""
    p = "head.predictor"
    B = ms[0].shape[0]
    nl = len(ms)  # 3 levels for fai_mf, 2 for bisenetformer (x[:-1], bisenetformer/modelling.py:378)
    src, pos, sizes = [], [], []
    for i in range(nl):
        h, w = ms[i].shape[-2:]
        sizes.append((h, w))
        pos.append(position_embedding_sine_normalized(h, w, cfg.hidden_dim // 2).flatten(2).permute(0, 2, 1))
        y = F.conv2d(ms[i], sd[f"{p}.input_proj.{i}.weight"], sd[f"{p}.input_proj.{i}.bias"])
        src.append(y.flatten(2).permute(0, 2, 1))
    qpos = sd[p + ".query_embed.weight"].unsqueeze(0)
    out = sd[p + ".query_feat.weight"].unsqueeze(0).repeat(B, 1, 1)
    hp = p + ".forward_prediction_heads"
    cls, masks, attn = prediction_heads(out, mask_features, sd, hp, sizes[0])
    for i in range(cfg.dec_layers):
        lvl = i % nl
        # rows that mask everything are un-masked (:510-512)
        keep = (attn.sum(-1) != attn.shape[-1]).unsqueeze(-1)
        attn = attn & keep
        # cross attention, pre-norm (transformer.py:206-238)
        c = f"{p}.transformer_cross_attention_layers.{i}"
        t2 = layer_norm(out, sd, c + ".norm")
        out = out + _mha_masked(t2 + qpos, src[lvl] + pos[lvl], src[lvl], attn, sd, c + ".multihead_attn", cfg.nhead)
        s = f"{p}.transformer_self_attention_layers.{i}"
        t2 = layer_norm(out, sd, s + ".norm")
        out = out + multihead_attention(t2 + qpos, t2 + qpos, t2, sd, s + ".self_attn", cfg.nhead)
        f = f"{p}.transformer_ffn_layers.{i}"
        t2 = layer_norm(out, sd, f + ".norm")
        out = out + linear(F.relu(linear(t2, sd, f + ".linear1")), sd, f + ".linear2")
        cls, masks, attn = prediction_heads(out, mask_features, sd, hp, sizes[(i + 1) % nl])
        if taps is not None:
            taps[f"dec{i}_out"] = out
    return cls, masks


def _mha_masked(q_in, k_in, v_in, mask_bqk, sd, p, nhead):
    """nn.MultiheadAttention with a boolean attn_mask [B,Q,K] shared by all heads (True = -inf)."""
    d = q_in.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(q_in, w[:d], b[:d])
    k = F.linear(k_in, w[d: 2 * d], b[d: 2 * d])
    v = F.linear(v_in, w[2 * d:], b[2 * d:])
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    hd = d // nhead
    q = q.view(B, Lq, nhead, hd).transpose(1, 2)
    k = k.view(B, Lk, nhead, hd).transpose(1, 2)
    v = v.view(B, Lk, nhead, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd))
    s = s.masked_fill(mask_bqk.unsqueeze(1), float("-inf"))
    o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, Lq, d)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def mf_forward(sd: SD, images: Tensor, cfg: MFOracleConfig, taps: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """FAIMaskFormer.forward eval (fai_mf/modelling.py:712-725) + MaskFormerHead.forward (:603-621):
    -> (class probs [B,Q,K] = softmax[..., :-1], mask probabilities [B,Q,H,W] = bilinear(sigmoid(mask logits)))."""
    mean = torch.tensor(list(cfg.pixel_mean), dtype=torch.float32).view(-1, 1, 1)
    std = torch.tensor(list(cfg.pixel_std), dtype=torch.float32).view(-1, 1, 1)
    x = (images - mean) / std
    mf, ms = transformer_fpn(x, sd, cfg, taps)
    cls, masks = masked_decoder(ms, mf, sd, cfg, taps)
    if taps is not None:
        taps["pred_logits"] = cls
        taps["pred_masks"] = masks  # pre-sigmoid mask logits at 1/4 resolution: the 1e-3 parity point
    probs = F.softmax(cls, dim=-1)[..., :-1]
    m = F.interpolate(masks.sigmoid(), size=images.shape[2:], mode="bilinear", align_corners=False)
    return probs, m


def mf_postprocess_tensors(logits: Tensor, masks: Tensor, cfg: MFOracleConfig, threshold: Optional[float] = None):
    """Tensor part of MaskFormerProcessor.postprocess for ONE image (the reference only works for B=1, SURVEY A.25;
    fai_mf/processor.py:204-262), instance mode: returns (kept query indices, scores, labels, boolean masks [n,H,W])."""
    assert logits.shape[0] == 1
    thr = threshold or cfg.threshold
    scores, labels = logits.max(-1)
    if cfg.predict_all_pixels:
        out = (scores.view(1, -1, 1, 1) * masks).argmax(dim=1)
        binm = torch.stack([out[0] == q for q in range(masks.shape[1])]).unsqueeze(0)
    else:
        binm = masks >= cfg.mask_threshold
    nz = (binm.sum(dim=(-2, -1)) > 1)[0].nonzero()[:, 0]
    scores, labels, binm, mp = scores[0, nz], labels[0, nz], binm[0, nz], masks[0, nz]
    if cfg.use_mask_score:
        bf = binm.int() * 1e-3
        scores = scores * ((bf * mp).sum(-1).sum(-1) / (bf.sum(-1).sum(-1) + 1e-5))
    keep = (scores > thr).nonzero()[:, 0] if thr > 0 else torch.arange(len(scores))
    return nz[keep], scores[keep], labels[keep], binm[keep]
