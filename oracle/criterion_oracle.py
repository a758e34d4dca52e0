"""TEST INFRASTRUCTURE ONLY - CPU restatement (torch fp32 + scipy) of the reference's DETR training criterion.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Pinned against the unmodified reference by oracle/gen_golden_criterion.py -> tests/golden/detr_criterion_b4.npz.

Follows (file:line in /root/reference):
  box_cxcywh_to_xyxy / box_iou / generalized_box_iou   focoos/utils/box.py:14-17,27-64
  BoxHungarianMatcher.forward                          focoos/models/fai_detr/modelling.py:693-758
  SetCriterion.loss_labels_vfl / loss_boxes / forward  focoos/models/fai_detr/modelling.py:464-499,513-531,553-598
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment


def cxcywh_to_xyxy(x):
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def pair_iou(a, b):
    """IoU / union of matching rows of two xyxy sets (the diagonal the reference extracts with torch.diag)."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = (torch.min(a[:, 2:], b[:, 2:]) - torch.max(a[:, :2], b[:, :2])).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    union = area_a + area_b - inter
    return inter / union, union


def pair_giou(a, b):
    iou, union = pair_iou(a, b)
    wh = (torch.max(a[:, 2:], b[:, 2:]) - torch.min(a[:, :2], b[:, :2])).clamp(min=0)
    area = wh[:, 0] * wh[:, 1]
    return iou - (area - union) / (area + 1e-5)


def match_cost(logits, boxes, labels, tboxes, w_class=2.0, w_bbox=5.0, w_giou=2.0, alpha=0.25, gamma=2.0):
    """one image: logits [Q,C], boxes [Q,4], labels [n], tboxes [n,4] -> cost [Q,n]   (modelling.py:722-741)"""
    p = torch.sigmoid(logits)[:, labels]
    neg = (1 - alpha) * (p**gamma) * (-(1 - p + 1e-8).log())
    pos = alpha * ((1 - p) ** gamma) * (-(p + 1e-8).log())
    Q, n = p.shape
    l1 = (boxes[:, None, :] - tboxes[None, :, :]).abs().sum(-1)
    a = cxcywh_to_xyxy(boxes)[:, None, :].expand(Q, n, 4).reshape(-1, 4)
    b = cxcywh_to_xyxy(tboxes)[None, :, :].expand(Q, n, 4).reshape(-1, 4)
    giou = pair_giou(a, b).reshape(Q, n)
    return w_bbox * l1 + w_class * (pos - neg) + w_giou * (-giou)


def hungarian(cost):
    """[Q,n] -> (query idx sorted ascending, target idx)   (scipy LSA, modelling.py:747)"""
    i, j = linear_sum_assignment(cost.detach().cpu().numpy())
    return torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)


def layer_losses(logits, boxes, targets, indices, num_boxes, alpha=0.75, gamma=2.0, weights=(1.0, 5.0, 2.0)):
    """logits [B,Q,C], boxes [B,Q,4]; targets = [(labels, boxes)]; indices = [(qi, tj)] -> (vfl, bbox, giou) weighted scalars."""
    B, Q, C = logits.shape
    bi = torch.cat([torch.full_like(q, b) for b, (q, _) in enumerate(indices)])
    qi = torch.cat([q for q, _ in indices])
    src = boxes[bi, qi]
    tgt = torch.cat([t[1][j] for t, (_, j) in zip(targets, indices)], 0)
    lab = torch.cat([t[0][j] for t, (_, j) in zip(targets, indices)], 0)
    iou, _ = pair_iou(cxcywh_to_xyxy(src), cxcywh_to_xyxy(tgt))
    iou = iou.detach()
    onehot = torch.zeros((B, Q, C), dtype=logits.dtype)
    onehot[bi, qi, lab] = 1.0
    tscore = torch.zeros((B, Q), dtype=logits.dtype)
    tscore[bi, qi] = iou
    tscore = tscore[..., None] * onehot
    p = torch.sigmoid(logits).detach()
    w = alpha * p.pow(gamma) * (1 - onehot) + tscore
    vfl = F.binary_cross_entropy_with_logits(logits, tscore, weight=w, reduction="none")
    vfl = vfl.mean(1).sum() * Q / num_boxes
    l1 = (src - tgt).abs().sum() / num_boxes
    giou = (1 - pair_giou(cxcywh_to_xyxy(src), cxcywh_to_xyxy(tgt))).sum() / num_boxes
    return weights[0] * vfl, weights[1] * l1, weights[2] * giou


def criterion(logits_l, boxes_l, targets, world_size=1):
    """all supervised layers: logits_l [L,B,Q,C], boxes_l [L,B,Q,4] (layer 0 = the final prediction).  Returns
    (losses [L,3], indices per layer) with the reference's key order: main, aux_0.., (modelling.py:563-598)."""
    n = sum(len(t[0]) for t in targets)
    num_boxes = max(float(n) / world_size, 1.0)
    out, all_idx = [], []
    for lg, bx in zip(logits_l, boxes_l):
        idx = [hungarian(match_cost(lg[b], bx[b], t[0], t[1])) for b, t in enumerate(targets)]
        all_idx.append(idx)
        out.append(torch.stack(layer_losses(lg, bx, targets, idx, num_boxes)))
    return torch.stack(out), all_idx


def synth_case(seed=4, B=4, Q=300, C=80, L=7):
    """SURVEY 8(d) train inputs: per image n in [1,20] boxes, cxcy in [0.2,0.8], wh in [0.05,0.35], labels in [0,C)."""
    g = torch.Generator().manual_seed(seed)
    targets = []
    for _ in range(B):
        n = int(torch.randint(1, 21, (1,), generator=g))
        cxcy = 0.2 + 0.6 * torch.rand((n, 2), generator=g)
        wh = 0.05 + 0.30 * torch.rand((n, 2), generator=g)
        targets.append((torch.randint(0, C, (n,), generator=g), torch.cat([cxcy, wh], 1)))
    logits = torch.randn((L, B, Q, C), generator=g) * 2.0 - 3.0
    boxes = torch.cat([0.1 + 0.8 * torch.rand((L, B, Q, 2), generator=g), 0.02 + 0.5 * torch.rand((L, B, Q, 2), generator=g)], -1)
    # plant near-hits so that IoUs are not all tiny
    for l in range(L):
        for b, (lab, tb) in enumerate(targets):
            for j in range(len(lab)):
                q = int(torch.randint(0, Q, (1,), generator=g))
                boxes[l, b, q] = (tb[j] + 0.03 * torch.randn(4, generator=g)).clamp(0.02, 0.98)
                logits[l, b, q, lab[j]] += 4.0
    return logits, boxes, targets
