"""Per-CTA timeline of one tcgen05 conv launch (clock64 stamps written by conv_tc_kernel when fb200_set_conv_trace is armed).
    python tools/conv_trace.py [shape names from tools/conv_micro.py ...]     env FB200_TC_CTA2=0|1|2"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from focoos_b200 import ops

SHAPES = {"rep_3x3_80": (32, 80, 80, 256, 256, 3, 1), "rep_3x3_40": (32, 40, 40, 256, 256, 3, 1), "s3_2b": (32, 20, 20, 512, 512, 3, 1),
          "csp_1x1_80": (32, 80, 80, 512, 512, 1, 1), "s2_2a": (32, 40, 40, 1024, 256, 1, 1), "s0_2c": (32, 160, 160, 64, 256, 1, 1)}
for n in [a for a in sys.argv[1:] if a in SHAPES] or ["rep_3x3_80", "rep_3x3_40"]:
    B, H, W, Cin, Cout, k, s = SHAPES[n]
    x = torch.randn((B, H, W, Cin), device="cuda").half()
    w = (torch.randn((Cout, k, k, Cin), device="cuda") * 0.05).half()
    bi = torch.zeros(Cout, device="cuda")
    y = torch.empty((B, H // s, W // s, Cout), device="cuda", dtype=torch.float16)
    if os.environ.get("FB200_GRID_CAP"):
        print("   (grid capped to", os.environ["FB200_GRID_CAP"], "CTAs)")
    for _ in range(3):
        ops.conv2d(x, w, None, bi, stride=s, pad=(k - 1) // 2, act=2, out=y, algo=ops.ALGO_TCGEN05)
    tr = torch.zeros((296 * 128,), dtype=torch.int64, device="cuda")
    ops.set_conv_trace(tr)
    ops.conv2d(x, w, None, bi, stride=s, pad=(k - 1) // 2, act=2, out=y, algo=ops.ALGO_TCGEN05)
    torch.cuda.synchronize()
    ops.set_conv_trace(None)
    t = tr.cpu().numpy().reshape(296, 128)
    used = np.nonzero(t[:, 1])[0]
    print(f"== {n}: {len(used)} CTAs traced; kernel entry spread {(t[used, 0].max() - t[used, 0].min())} ns")
    for cta in list(used[:2]) + list(used[-1:]):
        r = t[cta]
        base = r[1]
        line = [f"cta {cta}:"]
        for kk in range(20):
            a, f, i, e0, e1, p0 = (r[2 + 6 * kk + j] for j in range(6))
            if a == 0 and e0 == 0 and p0 == 0:
                break
            rel = lambda v: (v - base) if v else -1
            line.append(f"\n   tile{kk}: prod_start {rel(p0):7d} acc_free {rel(a):7d} first_ops {rel(f):7d} issued {rel(i):7d} | acc_done {rel(e0):7d} stored {rel(e1):7d}")
        print("".join(line))
    # aggregate: mean mainloop issue span, mean gap between consecutive tiles' issue end, epilogue span
    spans, epis, gaps = [], [], []
    for cta in used:
        r = t[cta]
        prev = None
        for kk in range(20):
            a, f, i, e0, e1 = (r[2 + 6 * kk + j] for j in range(5))
            if i:
                spans.append(i - a)
                if prev is not None:
                    gaps.append(i - prev)
                prev = i
            if e0 and e1:
                epis.append(e1 - e0)
    mm = [c for c in used if t[c, 125]]
    if mm:
        tot = np.array([t[c, 125] - t[c, 1] for c in mm], dtype=np.float64)
        print(f"   issuer CTAs: kernel span {tot.mean():.0f} cyc; waiting for operands {np.mean([t[c,122] for c in mm]):.0f} ({100*np.mean([t[c,122] for c in mm])/tot.mean():.0f}%), for a free accumulator {np.mean([t[c,123] for c in mm]):.0f};"
              f" producer waiting for a free stage {np.mean([t[c,124] for c in used]):.0f}")
    if spans:
        print(f"   MMA-thread issue span per tile: mean {np.mean(spans):.0f} cyc; tile-to-tile period {np.mean(gaps) if gaps else 0:.0f} cyc; epilogue (acc_done->stored) {np.mean(epis) if epis else 0:.0f} cyc")
