"""Generates tests/golden/detr_criterion_b4.npz from the UNMODIFIED reference (SetCriterion + BoxHungarianMatcher,
focoos/models/fai_detr/modelling.py) on seeded synthetic predictions/targets; run in the build container only:

    python -m oracle.gen_golden_criterion
"""
import os

import numpy as np
import torch

from oracle import criterion_oracle as CO
from oracle import ref_import


def main():
    ref_import.install()
    from focoos.models.fai_detr.modelling import BoxHungarianMatcher, SetCriterion
    from focoos.models.fai_detr.ports import DETRTargets

    logits, boxes, targets = CO.synth_case()
    L = logits.shape[0]
    crit = SetCriterion(num_classes=80, matcher=BoxHungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2, use_focal_loss=True, alpha=0.25, gamma=2.0),
                        weight_dict={"loss_vfl": 1, "loss_bbox": 5, "loss_giou": 2}, losses=["vfl", "boxes"], eos_coef=0.1, focal_alpha=0.75, focal_gamma=2.0)
    lg = logits.clone().requires_grad_(True)
    bx = boxes.clone().requires_grad_(True)
    outputs = {"pred_logits": lg[0], "pred_boxes": bx[0], "aux_outputs": [{"pred_logits": lg[i], "pred_boxes": bx[i]} for i in range(1, L)]}
    tg = [DETRTargets(labels=t[0], boxes=t[1]) for t in targets]
    losses = crit(outputs, tg)
    keys = ["", *[f"_{i}" for i in range(L - 1)]]
    table = torch.stack([torch.stack([losses["loss_vfl" + k], losses["loss_bbox" + k], losses["loss_giou" + k]]) for k in keys])
    sum(losses.values()).backward()
    # the matcher's indices per layer, as "query assigned to each target"
    match = []
    for l in range(L):
        idx = crit.matcher({"pred_logits": logits[l], "pred_boxes": boxes[l]}, tg)
        row = []
        for (qi, tj), t in zip(idx, targets):
            m = torch.empty(len(t[0]), dtype=torch.int64)
            m[tj] = qi
            row.append(m)
        match.append(torch.cat(row))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "detr_criterion_b4.npz")
    np.savez_compressed(out, losses=table.detach().numpy(), match_q=torch.stack(match).numpy().astype(np.int32),
                        grad_logits=lg.grad.numpy().astype(np.float32), grad_boxes=bx.grad.numpy().astype(np.float32), loss_keys=np.array(sorted(losses.keys())))
    print("wrote", out, os.path.getsize(out) // 1024, "KiB; total loss", float(sum(losses.values())), "n_targets", sum(len(t[0]) for t in targets))


if __name__ == "__main__":
    main()
