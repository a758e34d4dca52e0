"""GPU parity of the optimiser step through the C-ABI: grad_stats + optim_finalize + adamw_step against the
reference's own torch call sequence replayed on the CPU (oracle/optim_oracle.py) with identical gradients."""
import pytest
import torch
import torch.nn as nn

from focoos_b200 import DETRConfig, FAIDetr
from focoos_b200.train_step import FlatAdamW, get_optimizer_params
from oracle.optim_oracle import ReferenceStepper

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
SHAPES = [(7,), (65, 1027), (3, 3, 64, 33), (1,), (300_001,), (256, 256)]


def _params(seed, device):
    g = torch.Generator().manual_seed(seed)
    return [nn.Parameter((torch.randn(s, generator=g) * 0.1).to(device)) for s in SHAPES]


def _groups(ps):
    return [{"params": [p], "lr": 1e-3 * (0.1 if i % 2 else 1.0), "weight_decay": 0.0 if i == 3 else 0.02, "name": f"p{i}"} for i, p in enumerate(ps)]


@pytest.mark.parametrize("amp", [True, False])
def test_adamw_clip_scaler_match_reference_sequence(amp):
    ours, ref = _params(0, DEV), _params(0, "cpu")
    opt = FlatAdamW(_groups(ours), growth_interval=3, amp=amp)
    stepper = ReferenceStepper(_groups(ref), lr=1e-3, weight_decay=0.02, growth_interval=3, amp=amp)
    g = torch.Generator().manual_seed(1)
    for it in range(9):
        mag = [1.0, 1e-3, 50.0][it % 3]  # below / above the clip threshold
        grads = [torch.randn(s, generator=g) * mag for s in SHAPES]
        if it == 5 and amp:
            grads[2][0, 1, 2, 3] = float("nan")
        scale = float(opt.loss_scale)
        assert abs(scale - (stepper.scaler.get_scale() if amp else 1.0)) <= 1e-6 * scale
        stepper.step(lambda: sum((p * gr).sum() for p, gr in zip(ref, grads)))
        opt.zero_grad()
        for p, gr in zip(ours, grads):
            p.grad.copy_((gr * scale).to(DEV))
        opt.step()
        st = opt.stats()
        assert st["found_inf"] == (1 if (it == 5 and amp) else 0)
        for i, (p, q) in enumerate(zip(ours, ref)):
            d = float((p.detach().cpu() - q.detach()).abs().max())
            assert d <= 2e-6 * max(1.0, float(q.abs().max())), f"step {it} tensor {i}: {d:.3e}"
    assert opt.stats()["step"] == (8 if amp else 9)


def test_real_model_flat_buffer_and_step_time():
    """fai-detr-l: 501 tensors / 44.0 M elements in one flat buffer; parameters stay views of it; one step is three launches."""
    m = FAIDetr(DETRConfig(), precision="fp16").to(DEV)
    groups = get_optimizer_params(m, base_lr=5e-4, weight_decay=0.02, backbone_multiplier=0.1)
    opt = FlatAdamW(groups)
    assert len(groups) == 501 and opt.total >= 44_026_679
    before = opt.flat_params.clone()
    opt.flat_grads.normal_(generator=torch.Generator(device=DEV).manual_seed(0))
    opt.flat_grads.mul_(float(opt.loss_scale) * 1e-3)
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = opt.total * (4 + 28) / (ms * 1e-3) / 1e9  # stats reads g; adamw reads g,p,m,v and writes p,m,v
    print(f"optimizer step: {ms:.3f} ms for {opt.total / 1e6:.1f} M parameters = {gbs:.0f} GB/s")
    assert not torch.equal(before, opt.flat_params)
    p0 = groups[0]["params"][0]
    assert p0.data_ptr() == opt.flat_params.data_ptr() and p0.grad.data_ptr() == opt.flat_grads.data_ptr()
    st = opt.stats()
    assert st["found_inf"] == 0 and st["step"] == 13 and 0 < st["clip_coef"] < 1
    assert ms < 5.0
