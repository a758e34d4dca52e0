"""SyncBatchNorm / FrozenBatchNorm of the training graph (reference: trainer/trainer.py:330-334, nn/backbone/resnet.py:226-250) on a
GPU-less machine: host logic + collectives over a world_size-2 gloo group with the CPU reference operators installed as the ops backend.
Two ranks x B rows with statistics exchange must equal ONE process x 2B rows of plain train-mode BatchNorm - forward, input gradient, parameter
gradients (summed over ranks) and running statistics."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from focoos_b200 import autograd_ops as A
from focoos_b200 import ops
from oracle.ops_ref import RefBackend


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(act, with_res):
    g = torch.Generator().manual_seed(7)
    x = torch.randn((2, 6, 5, 8), generator=g) * 2 + 0.5          # [2B, H, W, C] NHWC, rank r owns image r
    res = torch.randn((2, 6, 5, 8), generator=g) if with_res else None
    dy = torch.randn((2, 6, 5, 8), generator=g)
    return x, res, dy


def _bn():
    torch.manual_seed(3)
    bn = nn.BatchNorm2d(8)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.3, 0.3)
        bn.running_mean.uniform_(-0.2, 0.2)
        bn.running_var.uniform_(0.5, 1.5)
    return bn


def _worker(rank, world, port, q, act, with_res):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ops._backend = RefBackend()
    x, res, dy = _data(act, with_res)
    bn = _bn()
    xs = x[rank:rank + 1].clone().requires_grad_(True)
    rs = None if res is None else res[rank:rank + 1].clone().requires_grad_(True)
    y = A.batch_norm_train(xs, bn, rs, act, sync_group=True)
    y.backward(dy[rank:rank + 1])
    q.put((rank, y.detach().numpy(), xs.grad.numpy(), None if rs is None else rs.grad.numpy(), bn.weight.grad.numpy(), bn.bias.grad.numpy(),
           bn.running_mean.numpy().copy(), bn.running_var.numpy().copy(), int(bn.num_batches_tracked)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("act,with_res", [(ops.ACT_NONE, False), (ops.ACT_RELU, True), (ops.ACT_SILU, False)])
def test_two_ranks_with_statistics_exchange_equal_one_process_on_the_full_batch(act, with_res):
    ops._backend = RefBackend()
    try:
        x, res, dy = _data(act, with_res)
        bn = _bn()
        xf = x.clone().requires_grad_(True)
        rf = None if res is None else res.clone().requires_grad_(True)
        yf = A.batch_norm_train(xf, bn, rf, act)
        yf.backward(dy)
    finally:
        ops._backend = None
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, act, with_res)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    for r, (rank, y, dx, dres, dg, db, rm, rv, nbt) in enumerate(out):
        assert abs(y - yf.detach().numpy()[r:r + 1]).max() < 1e-5
        assert abs(dx - xf.grad.numpy()[r:r + 1]).max() < 1e-5
        if with_res:
            assert abs(dres - rf.grad.numpy()[r:r + 1]).max() < 1e-6
        assert abs(rm - bn.running_mean.numpy()).max() < 1e-6 and abs(rv - bn.running_var.numpy()).max() < 1e-5, "running statistics = the full-batch ones on every rank"
        assert nbt == 1
    assert abs(out[0][4] + out[1][4] - bn.weight.grad.numpy()).max() < 1e-4, "dgamma: the gradient exchange's SUM over ranks is the full-batch gradient"
    assert abs(out[0][5] + out[1][5] - bn.bias.grad.numpy()).max() < 1e-4


def test_frozen_batch_norm_uses_the_running_statistics_and_trains_nothing():
    """FrozenBatchNorm2d semantics (resnet.py:226-250): y = (x - running_mean) / sqrt(running_var + eps) * w + b in TRAINING mode, buffers untouched,
    no weight / bias gradient, dx = dy * w / sqrt(running_var + eps) through the fused activation."""
    ops._backend = RefBackend()
    try:
        x, res, dy = _data(ops.ACT_RELU, True)
        bn = _bn()
        rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
        xs, rs = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
        y = A.batch_norm_train(xs, bn, rs, ops.ACT_RELU, frozen=True)
        y.backward(dy)
        xr, rr = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
        scale = bn.weight.detach() / torch.sqrt(rv0 + bn.eps)
        yr = torch.relu((xr - rm0) * scale + bn.bias.detach() + rr)
        yr.backward(dy)
        assert torch.allclose(y, yr, atol=1e-6) and torch.allclose(xs.grad, xr.grad, atol=1e-6) and torch.allclose(rs.grad, rr.grad, atol=1e-6)
        assert torch.equal(bn.running_mean, rm0) and torch.equal(bn.running_var, rv0) and int(bn.num_batches_tracked) == 0
        assert bn.weight.grad is None and bn.bias.grad is None
    finally:
        ops._backend = None


def test_model_flags_select_the_variant():
    from focoos_b200 import DETRConfig, FAIDetr, ResnetConfig
    m = FAIDetr(DETRConfig(backbone_config=ResnetConfig(freeze_norm=True, freeze_at=1)))
    bb = m.pixel_decoder.backbone
    assert all(getattr(mod, "frozen_stats", False) for mod in bb.modules() if isinstance(mod, nn.BatchNorm2d))
    assert not any(getattr(mod, "frozen_stats", False) for n, mod in m.named_modules() if isinstance(mod, nn.BatchNorm2d) and "backbone" not in n)
    assert not any(p.requires_grad for p in bb.conv1.parameters()) and not any(p.requires_grad for p in bb.res_layers[0].parameters())
    assert any(p.requires_grad for p in bb.res_layers[1].parameters() if p.dim() == 4)
    assert m.sync_bn is False and m.freeze_bn is False
