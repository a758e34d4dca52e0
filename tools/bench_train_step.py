"""Gradient exchange + optimiser step of the fai-detr-l fine-tune (SURVEY §8 a21 / BASELINE configs[4]) on N GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/bench_train_step.py
Times, on the device (max over ranks): the bucketed NCCL all-reduce of the 176 MB flat gradient buffer, and the three-launch
AdamW step.  Gradients are synthetic (the backward pass of the network is not built yet - DESIGN.md)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from focoos_b200 import DETRConfig, FAIDetr  # noqa: E402
from focoos_b200 import distributed as D  # noqa: E402
from focoos_b200.train_step import FlatAdamW, GradBucketReducer, get_optimizer_params  # noqa: E402


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    D.init_from_env("nccl", dev)
    torch.manual_seed(0)  # every rank must start from the same parameters (DDP broadcasts rank 0's; here: same seed)
    m = FAIDetr(DETRConfig(), precision="fp16").to(dev)
    opt = FlatAdamW(get_optimizer_params(m, base_lr=5e-4, weight_decay=0.02, backbone_multiplier=0.1), world_size=world)
    red = GradBucketReducer(opt, bucket_bytes=25 << 20)
    gen = torch.Generator(device=dev).manual_seed(rank)

    def fill():
        opt.flat_grads.normal_(generator=gen)
        opt.flat_grads.mul_(1e-3)

    def timed(fn, n=20):
        for _ in range(3):
            fill(); fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        tot = 0.0
        for _ in range(n):
            fill()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return D.max_over_ranks(tot / n, dev)

    def checksum_spread():
        chk = opt.flat_params.double().sum()
        lo, hi = chk.clone(), chk.clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return float(lo), float(hi)

    # ---- consistency phase (before any timing leg): rank-specific gradients -> exchange -> step, 5 times; replicas must stay bit-identical
    lo0, hi0 = checksum_spread()
    for _ in range(5):
        fill(); red.finish(); opt.step()
    torch.cuda.synchronize()
    lo1, hi1 = checksum_spread()
    consistent = {"identical_at_start": lo0 == hi0, "identical_after_5_exchanged_steps": lo1 == hi1, "params_moved": lo1 != lo0}

    ms_ar = timed(red.finish)
    ms_opt = timed(opt.step)
    ms_both = timed(lambda: (red.finish(), opt.step()))
    # NOTE: the optimizer-only timing leg above steps on rank-local gradients WITHOUT the exchange, so replicas legitimately differ
    # from here on; the replica check is the dedicated phase before the timing legs.
    nbytes = opt.total * 4
    if rank == 0:
        print(json.dumps({"what": "fai-detr-l gradient all-reduce + AdamW step", "n_gpus": world, "params": opt.total, "grad_bytes": nbytes, "buckets": len(red.buckets),
                          "allreduce_ms": ms_ar, "allreduce_busbw_GBps": (2 * (world - 1) / world * nbytes / (ms_ar * 1e-3) / 1e9) if world > 1 else None,
                          "optimizer_ms": ms_opt, "optimizer_GBps": opt.total * 32 / (ms_opt * 1e-3) / 1e9, "exchange_plus_step_ms": ms_both,
                          "replica_consistency": consistent, "stats": opt.stats()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
