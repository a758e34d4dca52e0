"""A/B of two builds of libfocoos_b200.so on the SAME box: fb200_conv2d (fp16, tcgen05) on representative layer shapes, interleaved launches.
    python tools/ab_lib.py focoos_b200/lib/libfocoos_b200_r01.so focoos_b200/lib/libfocoos_b200.so"""
import ctypes, sys, os
import torch

paths = [a for a in sys.argv[1:] if a.endswith(".so")]
libs = [ctypes.CDLL(os.path.abspath(p)) for p in paths]
SHAPES = {"s0_2c_res": (32, 160, 160, 64, 256, 1, 1, True), "s0_2b": (32, 160, 160, 64, 64, 3, 1, False), "s1_2c_res": (32, 80, 80, 128, 512, 1, 1, True),
          "s2_2b": (32, 40, 40, 256, 256, 3, 1, False), "s2_2a": (32, 40, 40, 1024, 256, 1, 1, False), "rep_3x3_80": (32, 80, 80, 256, 256, 3, 1, False),
          "csp_1x1_80": (32, 80, 80, 512, 512, 1, 1, False), "value_all": (1, 1, 268800, 256, 1536, 1, 1, False), "stem3": (32, 320, 320, 32, 64, 3, 1, False),
          "dec_lin": (1, 1, 9600, 256, 256, 1, 1, False)}
P = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
print(f"{'shape':12} " + " ".join(f"{os.path.basename(p)[15:-3][-22:]:>22}" for p in paths) + "   (us; ratios vs the first)")
for name, (B, H, W, Cin, Cout, k, s, res) in SHAPES.items():
    x = torch.randn((B, H, W, Cin), device="cuda").half()
    w = (torch.randn((Cout, k, k, Cin), device="cuda") * 0.05).half()
    bi = torch.zeros(Cout, device="cuda")
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    r = torch.randn((B, Ho, Wo, Cout), device="cuda").half() if res else None
    y = torch.empty((B, Ho, Wo, Cout), device="cuda", dtype=torch.float16)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    def call(lib):
        rc = lib.fb200_conv2d(P(x), 1, B, H, W, Cin, Cin, P(w), k, k, s, (k - 1) // 2, None, P(bi), P(r), Cout if res else 0, 1, P(y), 1, Cout, ctypes.c_int64(0), Cout, 2, st)
        assert rc == 0, rc
    ts = []
    for rep in range(3):       # interleave the two libraries so that clock / thermal drift hits both
        row = []
        for lib in libs:
            for _ in range(2): call(lib)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): call(lib)
            e1.record(); torch.cuda.synchronize()
            row.append(e0.elapsed_time(e1) * 100)
        ts.append(row)
    best = [min(t[i] for t in ts) for i in range(len(libs))]
    print(f"{name:12} " + " ".join(f"{b:22.1f}" for b in best) + "   " + " ".join(f"{b / best[0]:.3f}" for b in best[1:]))
