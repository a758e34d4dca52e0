"""GPU parity of the training criterion through the C-ABI: matching cost, device Hungarian assignment,
VFL/L1/GIoU losses and their gradients, against the CPU oracle and the golden fixture of the unmodified reference."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from focoos_b200 import ops
from focoos_b200 import criterion as K
from oracle import criterion_oracle as CO
from tests.parity_utils import load_golden

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _targets(targets):
    return [K.DETRTargets(labels=t[0].to(DEV), boxes=t[1].to(DEV)) for t in targets]


def _criterion():
    return K.SetCriterion(num_classes=80, matcher=K.BoxHungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2, use_focal_loss=True, alpha=0.25, gamma=2.0),
                          weight_dict={"loss_vfl": 1, "loss_bbox": 5, "loss_giou": 2}, losses=["vfl", "boxes"], focal_alpha=0.75, focal_gamma=2.0)


def test_match_cost_and_assignment_match_reference():
    g = load_golden("detr_criterion_b4")
    logits, boxes, targets = CO.synth_case()
    mq, cost = K.match(logits.to(DEV), boxes.to(DEV), _targets(targets), return_cost=True)
    o = 0
    for b, t in enumerate(targets):
        n = len(t[0])
        for l in (0, 3, 6):
            ref = CO.match_cost(logits[l, b], boxes[l, b], t[0], t[1]).T
            np.testing.assert_allclose(cost[l, o:o + n].cpu().numpy(), ref.numpy(), rtol=1e-5, atol=2e-5)
        o += n
    assert np.array_equal(mq.cpu().numpy(), g["match_q"]), "device assignment differs from the reference's scipy assignment"


@pytest.mark.parametrize("n,Q,seed", [(1, 300, 0), (20, 300, 1), (100, 300, 2), (300, 300, 3), (7, 8, 4), (64, 1000, 5)])
def test_hungarian_is_optimal(n, Q, seed):
    """random rectangular costs (targets x queries): same assignment as scipy (unique optimum for continuous costs)."""
    g = torch.Generator().manual_seed(seed)
    L, B = 2, 3
    cost = torch.rand((L, B * n, Q), generator=g) * 10 - 5
    toff = torch.arange(0, (B + 1) * n, n, dtype=torch.int32)
    mq = torch.full((L, B * n), -7, dtype=torch.int32, device=DEV)
    ops._be().hungarian(cost.to(DEV), toff.to(DEV), B, n, mq)
    mq = mq.cpu().numpy()
    for l in range(L):
        for b in range(B):
            blk = cost[l, b * n:(b + 1) * n].double().numpy()
            r, c = linear_sum_assignment(blk)
            got = mq[l, b * n:(b + 1) * n]
            assert len(set(got.tolist())) == n and got.min() >= 0 and got.max() < Q
            assert abs(blk[np.arange(n), got].sum() - blk[r, c].sum()) < 1e-9
            assert np.array_equal(got, c)


def test_losses_and_gradients_match_reference_golden():
    g = load_golden("detr_criterion_b4")
    logits, boxes, targets = CO.synth_case()
    L = logits.shape[0]
    lg, bx = logits.to(DEV).requires_grad_(True), boxes.to(DEV).requires_grad_(True)
    outputs = {"pred_logits": lg[0], "pred_boxes": bx[0], "aux_outputs": [{"pred_logits": lg[i], "pred_boxes": bx[i]} for i in range(1, L)]}
    losses = _criterion()(outputs, _targets(targets))
    assert sorted(losses.keys()) == sorted(g["loss_keys"].tolist())
    for l in range(L):
        sfx = "" if l == 0 else f"_{l - 1}"
        got = [float(losses[k + sfx].detach()) for k in ("loss_vfl", "loss_bbox", "loss_giou")]
        np.testing.assert_allclose(got, g["losses"][l], rtol=1e-5, atol=1e-6)
    sum(losses.values()).backward()
    np.testing.assert_allclose(lg.grad.cpu().numpy(), g["grad_logits"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(bx.grad.cpu().numpy(), g["grad_boxes"], rtol=1e-4, atol=1e-6)


def test_obj365_width_and_images_without_targets():
    """C=365 (the bench model's class count), one image with no boxes, one with a single box."""
    logits, boxes, targets = CO.synth_case(seed=11, B=3, Q=300, C=365, L=2)
    targets[1] = (torch.zeros(0, dtype=torch.int64), torch.zeros((0, 4)))
    targets[2] = (targets[2][0][:1], targets[2][1][:1])
    lg, bx = logits.clone().requires_grad_(True), boxes.clone().requires_grad_(True)
    ref, _ = CO.criterion(lg, bx, targets)
    ref.sum().backward()
    lgd, bxd = logits.to(DEV).requires_grad_(True), boxes.to(DEV).requires_grad_(True)
    out = _criterion()({"pred_logits": lgd[0], "pred_boxes": bxd[0], "aux_outputs": [{"pred_logits": lgd[1], "pred_boxes": bxd[1]}]}, _targets(targets))
    got = torch.stack([torch.stack([out["loss_vfl" + s], out["loss_bbox" + s], out["loss_giou" + s]]) for s in ("", "_0")])
    np.testing.assert_allclose(got.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    got.sum().backward()
    np.testing.assert_allclose(lgd.grad.cpu().numpy(), lg.grad.numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(bxd.grad.cpu().numpy(), bx.grad.numpy(), rtol=1e-4, atol=1e-6)


def test_all_images_empty():
    logits, boxes, _ = CO.synth_case(seed=12, B=2, Q=50, C=10, L=1)
    empty = [(torch.zeros(0, dtype=torch.int64), torch.zeros((0, 4)))] * 2
    out = _criterion()({"pred_logits": logits[0].to(DEV), "pred_boxes": boxes[0].to(DEV)}, _targets(empty))
    p = torch.sigmoid(logits[0])
    vfl = (0.75 * p.pow(2) * torch.nn.functional.binary_cross_entropy_with_logits(logits[0], torch.zeros_like(p), reduction="none")).sum()  # num_boxes clamps to 1
    assert abs(float(out["loss_vfl"]) - float(vfl)) <= 1e-5 * float(vfl)
    assert float(out["loss_bbox"]) == 0.0 and float(out["loss_giou"]) == 0.0
