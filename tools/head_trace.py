"""Per-CTA timeline of the small pair-format linears of the decoder (M = 9600 rows): where do the ~19 us of a 3-10 us GEMM go?
    python tools/head_trace.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from focoos_b200 import ops
from focoos_b200.fai_detr import _split3_weights

for M, K, N, pair_out in ((9600, 256, 256, False), (9600, 256, 256, True), (9600, 256, 1024, True), (9600, 1024, 256, False)):
    x = ops.Pair(ops.split_pair(torch.randn((1, 1, M, K), device="cuda")))
    w3 = _split3_weights(torch.randn((N, 1, 1, K), device="cuda") * 0.05)
    bias = torch.zeros(N, device="cuda")
    run = lambda: ops.conv2d_pair(x, w3, None, bias, act=0, out_pair=pair_out)  # noqa: E731
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record(); torch.cuda.synchronize()
    tr = torch.zeros((296 * 128,), dtype=torch.int64, device="cuda")
    ops.set_conv_trace(tr)
    run()
    torch.cuda.synchronize()
    ops.set_conv_trace(None)
    t = tr.cpu().numpy().reshape(296, 128)
    used = np.nonzero(t[:, 1])[0]
    print(f"== M={M} K={K} N={N} pair_out={pair_out}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us back-to-back; {len(used)} CTAs traced; kernel-entry spread {t[used, 0].max() - t[used, 0].min()} ns")
    for cta in list(used[:2]) + list(used[-1:]):
        r = t[cta]
        base = r[1]
        rel = lambda v: int(v - base) if v else -1  # noqa: E731
        line = f"  cta {cta}:"
        for kk in range(4):
            a, f, i, d, s_, p0 = (r[2 + 6 * kk + j] for j in range(6))
            if a == 0 and d == 0 and p0 == 0:
                break
            line += f" | tile{kk}: prod {rel(p0)} acc_free {rel(a)} first_ops {rel(f)} issued {rel(i)} acc_done {rel(d)} stored {rel(s_)}"
        print(line)
