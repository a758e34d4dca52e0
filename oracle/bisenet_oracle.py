"""CPU oracle: functional fp32 restatement of the reference's BisenetFormer inference path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Restates `focoos/nn/backbone/stdc.py` (STDC, CatBottleneck, ConvX) and
`focoos/models/bisenetformer/modelling.py` (ContextPath, AttentionRefinementModule, FeatureFusionModule, BiseNet, the 2-level masked
TransformerDecoder, head, final interpolate) plus the tensor part of the semantic post-process
(`focoos/models/bisenetformer/processor.py:163-301`, identical to fai_mf's).  Pinned by tests against fixtures from the unmodified reference
(oracle/gen_golden_bisenet.py)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .detr_oracle import batchnorm_eval
from .mf_oracle import masked_decoder

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass
class BisenetOracleConfig:
    num_classes: int = 150
    num_queries: int = 100
    layers: Sequence[int] = (4, 5, 3)
    base: int = 64
    feat_dim: int = 128
    hidden_dim: int = 256
    nhead: int = 8
    dec_layers: int = 6
    pixel_mean: Sequence[float] = (123.675, 116.28, 103.53)
    pixel_std: Sequence[float] = (58.395, 57.12, 57.375)
    threshold: float = 0.5
    mask_threshold: float = 0.5
    predict_all_pixels: bool = True
    use_mask_score: bool = False


def conv_bn_relu(x: Tensor, sd: SD, p: str, stride: int = 1, bn: str = "bn", relu: bool = True) -> Tensor:
    """ConvX (nn/backbone/stdc.py:20-38) / ConvBNReLU (bisenetformer/modelling.py:122-146): conv(no bias, pad k//2) + BN + ReLU."""
    w = sd[p + ".conv.weight"]
    y = batchnorm_eval(F.conv2d(x, w, None, stride, w.shape[-1] // 2), sd, f"{p}.{bn}")
    return F.relu(y) if relu else y


def cat_bottleneck(x: Tensor, sd: SD, p: str, stride: int) -> Tensor:
    """CatBottleneck.forward (nn/backbone/stdc.py:153-172)."""
    out1 = conv_bn_relu(x, sd, p + ".conv_list.0")
    outs = []
    out = out1
    for idx in range(1, 4):
        if idx == 1 and stride == 2:
            w = sd[p + ".avd_layer.0.weight"]
            out = batchnorm_eval(F.conv2d(out1, w, None, 2, 1, 1, w.shape[0]), sd, p + ".avd_layer.1")
        out = conv_bn_relu(out, sd, f"{p}.conv_list.{idx}")
        outs.append(out)
    if stride == 2:
        out1 = F.avg_pool2d(out1, 3, 2, 1)
    return torch.cat([out1] + outs, dim=1)


def stdc(x: Tensor, sd: SD, p: str, layers: Sequence[int]) -> Dict[str, Tensor]:
    """STDC.forward (nn/backbone/stdc.py:314-321): features[0,1] ConvX 3x3 s2; out_ids per `layers`."""
    x = conv_bn_relu(x, sd, p + ".features.0", 2)
    x = conv_bn_relu(x, sd, p + ".features.1", 2)
    outs = {"res2": x}
    idx = 2
    for i, n in enumerate(layers):
        for j in range(n):
            x = cat_bottleneck(x, sd, f"{p}.features.{idx}", 2 if j == 0 else 1)
            idx += 1
        outs[f"res{i + 3}"] = x
    return outs


def arm(x: Tensor, sd: SD, p: str) -> Tensor:
    """AttentionRefinementModule.forward (bisenetformer/modelling.py:159-167)."""
    feat = conv_bn_relu(F.conv2d(x, sd[p + ".proj.weight"]), sd, p + ".conv")
    att = feat.mean(dim=(2, 3), keepdim=True)
    att = torch.sigmoid(batchnorm_eval(F.conv2d(att, sd[p + ".conv_atten.weight"]), sd, p + ".bn_atten"))
    return feat * att


def bisenet_pixel_decoder(images_norm: Tensor, sd: SD, cfg: BisenetOracleConfig, taps: Optional[dict] = None):
    """BiseNet.forward_features (bisenetformer/modelling.py:276-282) with ContextPath (:186-210) and FFM (:224-235)."""
    p = "pixel_decoder"
    f = stdc(images_norm, sd, p + ".backbone", cfg.layers)
    res3, res4, res5 = f["res3"], f["res4"], f["res5"]
    avg = conv_bn_relu(res5.mean(dim=(2, 3), keepdim=True), sd, p + ".cp.conv_avg")
    f32 = arm(res5, sd, p + ".cp.arm32") + avg
    up = conv_bn_relu(F.interpolate(f32, size=res4.shape[-2:], mode="bilinear"), sd, p + ".cp.conv_head32")
    f16 = arm(res4, sd, p + ".cp.arm16") + up
    f8 = conv_bn_relu(F.interpolate(f16, size=res3.shape[-2:], mode="bilinear"), sd, p + ".cp.conv_head16")
    # FFM
    q = p + ".ffm"
    feat = F.conv2d(res3, sd[q + ".proj1.weight"], sd[q + ".proj1.bias"]) + F.conv2d(f8, sd[q + ".proj2.weight"], sd[q + ".proj2.bias"])
    feat = conv_bn_relu(feat, sd, q + ".convblk")
    att = F.adaptive_avg_pool2d(feat, 1)
    att = torch.sigmoid(F.conv2d(F.relu(F.conv2d(att, sd[q + ".conv1.weight"])), sd[q + ".conv2.weight"]))
    fuse = feat * att + feat
    out = conv_bn_relu(fuse, sd, p + ".conv_out")
    if taps is not None:
        taps.update(f)
        taps.update(cp32=f32, cp16=f16, cp8=f8, mask_features=out)
    return out, [f32, f16, f8]


def bisenet_forward(sd: SD, images: Tensor, cfg: BisenetOracleConfig, taps: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """BisenetFormer.forward eval (bisenetformer/modelling.py:609-622): -> (class probs [B,Q,K], mask probabilities [B,Q,H,W])."""
    mean = torch.tensor(list(cfg.pixel_mean), dtype=torch.float32).view(-1, 1, 1)
    std = torch.tensor(list(cfg.pixel_std), dtype=torch.float32).view(-1, 1, 1)
    x = (images - mean) / std
    mf, ms = bisenet_pixel_decoder(x, sd, cfg, taps)
    cls, masks = masked_decoder(ms[:-1], mf, sd, cfg, taps)  # F1 and F2 only (:378)
    if taps is not None:
        taps["pred_logits"], taps["pred_masks"] = cls, masks
    probs = F.softmax(cls, dim=-1)[..., :-1]
    return probs, F.interpolate(masks.sigmoid(), size=images.shape[2:], mode="bilinear", align_corners=False)


def semantic_postprocess_tensors(logits: Tensor, masks: Tensor, cfg: BisenetOracleConfig, threshold: Optional[float] = None):
    """bisenetformer/processor.py:204-262 for ONE image, predict_all_pixels=True: per-pixel argmax_q(score_q * prob_q) -> one-hot masks,
    drop masks with <= 1 pixel, score threshold.  Returns (kept query idx, scores, labels, bool masks)."""
    assert logits.shape[0] == 1
    thr = threshold or cfg.threshold
    scores, labels = logits.max(-1)
    out = (scores.view(1, -1, 1, 1) * masks).argmax(dim=1)
    binm = torch.stack([out[0] == q for q in range(masks.shape[1])])
    nz = (binm.sum(dim=(-2, -1)) > 1).nonzero()[:, 0]
    s, l, bm = scores[0, nz], labels[0, nz], binm[nz]
    keep = (s > thr).nonzero()[:, 0] if thr > 0 else torch.arange(len(s))
    return nz[keep], s[keep], l[keep], bm[keep]
