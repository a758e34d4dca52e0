"""Micro-benchmark of the tcgen05 conv kernel on representative layer shapes (CUDA events, inputs > L2 where the layer is).
    python tools/conv_micro.py [names...]      (env FB200_TC_BN=64|128|256 forces the N tile)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from focoos_b200 import ops

SHAPES = {  # name: (B,H,W,Cin,Cout,k,stride,res,act)
    "s0_2c_res": (32, 160, 160, 64, 256, 1, 1, True, 1),
    "s0_2a": (32, 160, 160, 256, 64, 1, 1, False, 1),
    "s0_2b": (32, 160, 160, 64, 64, 3, 1, False, 1),
    "s1_2c_res": (32, 80, 80, 128, 512, 1, 1, True, 1),
    "s2_2c_res": (32, 40, 40, 256, 1024, 1, 1, True, 1),
    "s2_2b": (32, 40, 40, 256, 256, 3, 1, False, 1),
    "s3_2b": (32, 20, 20, 512, 512, 3, 1, False, 1),
    "csp_1x1_80": (32, 80, 80, 512, 512, 1, 1, False, 2),
    "csp_1x1_40": (32, 40, 40, 512, 512, 1, 1, False, 2),
    "rep_3x3_80": (32, 80, 80, 256, 256, 3, 1, False, 2),
    "rep_3x3_40": (32, 40, 40, 256, 256, 3, 1, False, 2),
    "value_all": (1, 1, 268800, 256, 1536, 1, 1, False, 0),
    "stem2": (32, 320, 320, 32, 32, 3, 1, False, 1),
    "stem3": (32, 320, 320, 32, 64, 3, 1, False, 1),
    "s1_2b": (32, 80, 80, 128, 128, 3, 1, False, 1),
    "s0_2c": (32, 160, 160, 64, 256, 1, 1, False, 1),
    "s1_2a": (32, 80, 80, 512, 128, 1, 1, False, 1),
}
PAIR = "--pair" in sys.argv    # fp32-accurate mode with pair-format input, residual and output (fb200_conv2d_pair): what the fp32_tc engines run
SPLIT = "--split" in sys.argv  # fp32-accurate mode: [hi|lo] pair input, [W_hi|W_lo|W_hi] weights, fp32 output (TF/s = algorithmic flops)
names = [a for a in sys.argv[1:] if a in SHAPES] or list(SHAPES)
reps = 10
print(f"{'name':12} {'us':>8} {'TF/s':>7} {'GB/s':>7}  shape   (FB200_TC_BN={os.environ.get('FB200_TC_BN','auto')})")
for n in names:
    B, H, W, Cin, Cout, k, s, res, act = SHAPES[n]
    bi = torch.zeros(Cout, device="cuda")
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    if PAIR:
        if Cin % 32:
            continue
        from focoos_b200.fai_detr import _split3_weights
        xp = ops.Pair(ops.split_pair(torch.randn((B, H, W, Cin), device="cuda")))
        w = _split3_weights(torch.randn((Cout, k, k, Cin), device="cuda") * 0.05)
        rp = ops.Pair(ops.split_pair(torch.randn((B, Ho, Wo, Cout), device="cuda"))) if res else None
        yp = ops.Pair.empty((B, Ho, Wo, Cout), "cuda")
        run = lambda: ops.conv2d_pair(xp, w, None, bi, stride=s, pad=(k - 1) // 2, act=act, residual=rp, out=yp)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        M = B * Ho * Wo
        fl = 2.0 * M * Cout * k * k * Cin
        by = 4.0 * (B * H * W * Cin + Cout * k * k * Cin + M * Cout * (2 if res else 1))
        print(f"{n:12} {us:8.1f} {fl/us/1e6:7.0f} {by/us/1e3:7.0f}  {H}x{W} {Cin}->{Cout} k{k} s{s}{' +res' if res else ''} pair")
        continue
    if SPLIT:
        if Cin % 32:
            continue
        from focoos_b200.fai_detr import _split3_weights
        x = ops.split_pair(torch.randn((B, H, W, Cin), device="cuda"))
        w = _split3_weights(torch.randn((Cout, k, k, Cin), device="cuda") * 0.05)
        r = torch.randn((B, Ho, Wo, Cout), device="cuda") if res else None
        y = torch.empty((B, Ho, Wo, Cout), device="cuda", dtype=torch.float32)
        ALGO = ops.ALGO_TCGEN05_SPLIT3
    else:
        x = torch.randn((B, H, W, Cin), device="cuda").half()
        w = (torch.randn((Cout, k, k, Cin), device="cuda") * 0.05).half()
        r = torch.randn((B, Ho, Wo, Cout), device="cuda").half() if res else None
        y = torch.empty((B, Ho, Wo, Cout), device="cuda", dtype=torch.float16)
        ALGO = ops.ALGO_TCGEN05
    for _ in range(2):
        ops.conv2d(x, w, None, bi, stride=s, pad=(k - 1) // 2, act=act, residual=r, out=y, algo=ALGO)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.conv2d(x, w, None, bi, stride=s, pad=(k - 1) // 2, act=act, residual=r, out=y, algo=ALGO)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    M = B * Ho * Wo
    fl = 2.0 * M * Cout * k * k * Cin
    by = (4.0 if SPLIT else 2.0) * (B * H * W * Cin + Cout * k * k * Cin + M * Cout * (2 if res else 1))
    print(f"{n:12} {us:8.1f} {fl/us/1e6:7.0f} {by/us/1e3:7.0f}  {H}x{W} {Cin}->{Cout} k{k} s{s}{' +res' if res else ''}")
