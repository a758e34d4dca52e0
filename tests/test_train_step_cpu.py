"""Optimiser / data-parallel step without a GPU: parameter groups against the unmodified reference's,
the flat-buffer AdamW + clipping + loss-scaling host logic against the reference's own torch call sequence
(oracle/optim_oracle.py), and the bucketed gradient all-reduce on a world_size-2 gloo group."""
import json
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from focoos_b200 import DETRConfig, FAIDetr, ops
from focoos_b200.train_step import FlatAdamW, GradBucketReducer, get_optimizer_params
from oracle.ops_ref import RefBackend
from oracle.optim_oracle import ReferenceStepper

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture()
def ref_backend():
    ops._backend = RefBackend()
    yield
    ops._backend = None


def test_param_groups_match_reference():
    g = json.load(open(os.path.join(GOLDEN, "fai_detr_l_param_groups.json")))
    groups = get_optimizer_params(FAIDetr(DETRConfig()), base_lr=g["base_lr"], weight_decay=g["weight_decay"], weight_decay_norm=0.0, weight_decay_embed=0.0,
                                  backbone_multiplier=0.1)
    got = [[x["name"], x["lr"], x["weight_decay"], x["params"][0].numel()] for x in groups]
    assert len(got) == len(g["groups"]) == 501
    assert [r[0] for r in got] == [r[0] for r in g["groups"]], "same tensors in the same order"
    for a, b in zip(got, g["groups"]):
        assert a[3] == b[3] and abs(a[1] - b[1]) < 1e-12 and abs(a[2] - b[2]) < 1e-12, (a, b)
    assert sum(r[3] for r in got) == g["total_elements"]


class _Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.backbone = nn.Sequential(nn.Linear(7, 13), nn.ReLU())
        self.pixel_decoder = nn.Sequential(nn.Linear(13, 10), nn.LayerNorm(10))
        self.head = nn.Linear(10, 3)

    def forward(self, x):
        return self.head(self.pixel_decoder(self.backbone(x)))


def _pair(seed=0):
    torch.manual_seed(seed)
    a = _Tiny()
    b = _Tiny()
    b.load_state_dict(a.state_dict())
    return a, b


def _groups(m):
    return get_optimizer_params(m, base_lr=5e-2, weight_decay=0.02, weight_decay_norm=0.0, backbone_multiplier=0.1)


def test_flat_adamw_follows_the_reference_step_sequence(ref_backend):
    ours, ref = _pair()
    opt = FlatAdamW(_groups(ours), growth_interval=3, chunk_elems=16)
    stepper = ReferenceStepper(_groups(ref), lr=5e-2, weight_decay=0.02, growth_interval=3)
    g = torch.Generator().manual_seed(1)
    for it in range(8):
        x, y = torch.randn((5, 7), generator=g), torch.randn((5, 3), generator=g)
        if it == 4:
            y[0, 0] = float("inf")  # -> non-finite gradients: the step must be skipped and the loss scale halved
        scale_ref = stepper.scaler.get_scale()
        assert abs(float(opt.loss_scale) - scale_ref) < 1e-6 * scale_ref, f"step {it}: loss scale"
        stepper.step(lambda: ((ref(x) - y) ** 2).mean() * (10.0 if it < 6 else 1e-3))
        opt.zero_grad()
        opt.scale_loss(((ours(x) - y) ** 2).mean() * (10.0 if it < 6 else 1e-3)).backward()
        opt.step()
        st = opt.stats()
        assert st["found_inf"] == (1 if it == 4 else 0), f"step {it}: inf detection"
        for (n, p), q in zip(ours.named_parameters(), ref.parameters()):
            assert torch.allclose(p, q, rtol=2e-6, atol=1e-7), f"step {it}: {n} differs by {float((p - q).abs().max()):.3e}"
    assert opt.stats()["step"] == 7, "the skipped step must not advance Adam's step count"
    sd = opt.state_dict()
    opt.load_state_dict(sd)


def test_no_cpu_fallback():
    m = _Tiny()
    with pytest.raises(RuntimeError):
        FlatAdamW(_groups(m))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ddp_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)  # two workers on a shared host: do not oversubscribe the cores
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ops._backend = RefBackend()
    torch.manual_seed(0)
    m = _Tiny()
    opt = FlatAdamW(_groups(m), world_size=world, chunk_elems=16)
    red = GradBucketReducer(opt, bucket_bytes=256)  # tiny buckets: several all-reduces per step
    red.attach_hooks()
    assert len(red.buckets) > 2
    for it in range(3):
        g = torch.Generator().manual_seed(100 * it + rank)
        x, y = torch.randn((4, 7), generator=g), torch.randn((4, 3), generator=g)
        opt.zero_grad()
        opt.scale_loss(((m(x) - y) ** 2).mean()).backward()
        red.finish()
        opt.step()
    q.put((rank, [p.detach().numpy().copy() for p in m.parameters()]))  # numpy: pickled by value (tensors go through fd passing and race with worker exit)
    dist.destroy_process_group()


def test_two_rank_gradient_exchange_equals_full_batch_reference():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r: [torch.from_numpy(a) for a in ps] for r, ps in (q.get(timeout=150) for _ in range(world))}
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    torch.manual_seed(0)
    ref = _Tiny()
    stepper = ReferenceStepper(_groups(ref), lr=5e-2, weight_decay=0.02)
    for it in range(3):
        data = []
        for rank in range(world):
            g = torch.Generator().manual_seed(100 * it + rank)
            data.append((torch.randn((4, 7), generator=g), torch.randn((4, 3), generator=g)))
        stepper.step(lambda: sum(((ref(x) - y) ** 2).mean() for x, y in data) / world)  # DDP averages the per-rank gradients
    for rank in range(world):
        for p, r in zip(res[rank], ref.parameters()):
            assert torch.allclose(p, r, rtol=2e-6, atol=1e-7), float((p - r).abs().max())
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b), "replicas must stay bit-identical"


def test_bucket_layout_and_unused_parameter_tracking(ref_backend):
    """GradBucketReducer: buckets tile the flat gradient buffer exactly once, in reverse parameter order, never splitting a tensor; hooks launch a bucket
    when its last tensor is ready.  FlatAdamW: tensors the loss never reaches are skipped entirely (torch.optim semantics for p.grad is None); state round-trips."""
    torch.manual_seed(0)
    m = _Tiny()
    dead = nn.Linear(3, 3)  # never used in forward
    m.add_module("dead", dead)
    opt = FlatAdamW(_groups(m), amp=False, chunk_elems=16)
    opt.track_unused_parameters()
    red = GradBucketReducer(opt, bucket_bytes=200)
    spans = sorted((s, e) for s, e, _, _ in red.buckets)
    assert spans[0][0] == 0 and spans[-1][1] == opt.total and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert [b[3] for b in red.buckets] == sorted([b[3] for b in red.buckets], reverse=True), "reverse parameter order"
    for s, e, lo, hi in red.buckets:
        assert s == opt.offsets[lo] and e == (opt.offsets[hi + 1] if hi + 1 < len(opt.offsets) else opt.total)
    launched = []
    red._launch = lambda b, _orig=red._launch: (launched.append(b), _orig(b))[1]
    red.attach_hooks()
    before = opt.flat_params.clone()
    x, y = torch.randn(4, 7), torch.randn(4, 3)
    opt.zero_grad()
    ((m(x) - y) ** 2).mean().backward()
    # first pass: collectives go out STRICTLY in bucket order on every rank, so the unused tensors (the LAST parameters = bucket 0) hold everything back
    # until finish(); finish() learns which tensors are structurally unused ...
    assert launched == [], launched
    red.finish()
    assert launched == list(range(len(red.buckets))), "finish() launches in index order"
    assert red._static_unused == {i for i, n in enumerate(opt.names) if n.startswith("dead")}
    opt.step()
    # ... and from the second pass on they count as ready: every bucket leaves from the hooks, in order, overlapped with backward
    launched.clear()
    opt.zero_grad()
    ((m(x) - y) ** 2).mean().backward()
    assert launched == list(range(len(red.buckets))), launched
    with pytest.raises(RuntimeError):  # a second backward before finish() would add onto already-exchanged slices
        ((m(x) - y) ** 2).mean().backward()
    red.reset()
    opt.zero_grad()
    ((m(x) - y) ** 2).mean().backward()
    red.finish()
    before = before.clone()
    for i, n in enumerate(opt.names):
        o, cnt = opt.offsets[i], opt.params[i].numel()
        moved = not torch.equal(before[o:o + cnt], opt.flat_params[o:o + cnt])
        assert moved == (not n.startswith("dead")), n
    sd = opt.state_dict()
    opt2 = FlatAdamW(_groups(m), amp=False, chunk_elems=16)
    opt2.load_state_dict(sd)
    assert torch.equal(opt2.exp_avg, opt.exp_avg) and opt2.stats()["step"] == 1
