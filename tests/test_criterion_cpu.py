"""Training criterion without a GPU: (1) the CPU oracle against the golden fixture produced by the
unmodified reference SetCriterion/BoxHungarianMatcher (oracle/gen_golden_criterion.py); (2) the host logic of
focoos_b200.criterion (target packing, layer stacking, loss keys, autograd hand-over) with the per-op CPU references
installed as the backend."""
import numpy as np
import pytest
import torch

from focoos_b200 import ops
from focoos_b200.criterion import BoxHungarianMatcher, DETRTargets, SetCriterion
from oracle import criterion_oracle as CO
from oracle.ops_ref import RefBackend
from tests.parity_utils import load_golden


@pytest.fixture()
def ref_backend():
    ops._backend = RefBackend()
    yield
    ops._backend = None


def _match_from(idx, targets):
    rows = []
    for (qi, tj), t in zip(idx, targets):
        m = torch.empty(len(t[0]), dtype=torch.int64)
        m[tj] = qi
        rows.append(m)
    return torch.cat(rows).numpy()


def test_oracle_matches_reference_golden():
    g = load_golden("detr_criterion_b4")
    logits, boxes, targets = CO.synth_case()
    lg, bx = logits.clone().requires_grad_(True), boxes.clone().requires_grad_(True)
    table, idx = CO.criterion(lg, bx, targets)
    np.testing.assert_allclose(table.detach().numpy(), g["losses"], rtol=2e-6, atol=1e-6)
    for l in range(logits.shape[0]):
        assert np.array_equal(_match_from(idx[l], targets), g["match_q"][l]), f"layer {l}: assignment differs from the reference"
    table.sum().backward()
    np.testing.assert_allclose(lg.grad.numpy(), g["grad_logits"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(bx.grad.numpy(), g["grad_boxes"], rtol=1e-5, atol=1e-7)


def _criterion():
    return SetCriterion(num_classes=80, matcher=BoxHungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2, use_focal_loss=True, alpha=0.25, gamma=2.0),
                        weight_dict={"loss_vfl": 1, "loss_bbox": 5, "loss_giou": 2}, losses=["vfl", "boxes"], focal_alpha=0.75, focal_gamma=2.0)


def test_set_criterion_host_logic(ref_backend):
    g = load_golden("detr_criterion_b4")
    logits, boxes, targets = CO.synth_case()
    L = logits.shape[0]
    lg, bx = logits.clone().requires_grad_(True), boxes.clone().requires_grad_(True)
    outputs = {"pred_logits": lg[0], "pred_boxes": bx[0], "aux_outputs": [{"pred_logits": lg[i], "pred_boxes": bx[i]} for i in range(1, L)]}
    losses = _criterion()(outputs, [DETRTargets(labels=t[0], boxes=t[1]) for t in targets])
    assert sorted(losses.keys()) == sorted(g["loss_keys"].tolist())
    for l in range(L):
        sfx = "" if l == 0 else f"_{l - 1}"
        got = [float(losses[k + sfx]) for k in ("loss_vfl", "loss_bbox", "loss_giou")]
        np.testing.assert_allclose(got, g["losses"][l], rtol=2e-6, atol=1e-6)
    sum(losses.values()).backward()
    np.testing.assert_allclose(lg.grad.numpy(), g["grad_logits"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(bx.grad.numpy(), g["grad_boxes"], rtol=1e-5, atol=1e-7)


def test_matcher_return_contract(ref_backend):
    logits, boxes, targets = CO.synth_case(seed=9, B=3, Q=50, C=12, L=1)
    out = _criterion().matcher({"pred_logits": logits[0], "pred_boxes": boxes[0]}, [DETRTargets(labels=t[0], boxes=t[1]) for t in targets])
    for b, (qi, tj) in enumerate(out):
        ri, rj = CO.hungarian(CO.match_cost(logits[0, b], boxes[0, b], targets[b][0], targets[b][1]))
        assert qi.dtype == torch.int64 and torch.equal(qi, ri) and torch.equal(tj, rj)


def test_no_cpu_fallback_and_unsupported_configs():
    logits, boxes, targets = CO.synth_case(seed=9, B=1, Q=20, C=5, L=1)
    with pytest.raises(RuntimeError):
        _criterion()({"pred_logits": logits[0], "pred_boxes": boxes[0]}, [DETRTargets(labels=targets[0][0], boxes=targets[0][1])])
    with pytest.raises(NotImplementedError):
        BoxHungarianMatcher(use_focal_loss=False)
    with pytest.raises(NotImplementedError):
        SetCriterion(80, BoxHungarianMatcher(), {}, losses=["labels"])
