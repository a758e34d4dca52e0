"""-m gpu: every CUDA kernel (through the C ABI) against its CPU reference in oracle/ops_ref.py on seeded inputs.
fp32 kernels: tolerance = fp32 reassociation (1e-4 relative to the tensor scale).  fp16 kernels: inputs are the same
fp16-rounded values on both sides, the reference accumulates in fp32, tolerance = one fp16 output rounding (2e-3)."""
import math

import numpy as np
import pytest
import torch

from focoos_b200 import ops
from oracle.ops_ref import RefBackend

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
REF = RefBackend()
DEV = "cuda"


def tol(dtype):
    return 1e-4 if dtype == torch.float32 else 3e-3


def close(a_gpu, b_cpu, dtype, what=""):
    a, b = a_gpu.detach().float().cpu(), b_cpu.detach().float()
    scale = max(1.0, float(b.abs().max()))
    err = float((a - b).abs().max())
    assert err <= tol(dtype) * scale, f"{what}: max|d|={err:.3e} scale={scale:.3e}"


def rnd(shape, dtype, seed, s=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * s).to(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_stem_conv(dtype):
    img = torch.rand((2, 3, 64, 96), generator=torch.Generator().manual_seed(0)) * 255
    w = rnd((32, 3, 3, 3), torch.float32, 1, 0.3)
    sc, bi = torch.rand(32) + 0.5, rnd((32,), torch.float32, 2, 0.1)
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    ref = torch.empty((2, 32, 48, 32), dtype=dtype)
    REF.stem_conv(img, w, sc, bi, mean, std, 1, ref)
    out = ops.stem_conv(img.to(DEV), w.to(DEV), sc.to(DEV), bi.to(DEV), mean, std, 1, dtype)
    close(out, ref, dtype, "stem")


CONV_CASES = [  # B,H,W,Cin,Cout,k,stride,act,res,scale
    (2, 20, 20, 32, 64, 3, 1, 1, True, True),
    (2, 20, 20, 64, 64, 1, 1, 2, False, True),
    (1, 17, 23, 32, 48, 3, 2, 0, False, False),
    (1, 1, 300, 4, 512, 1, 1, 1, False, False),     # query_pos_head layer 0 (K = 4)
    (1, 1, 77, 256, 365, 1, 1, 0, False, False),    # class head: Cout not a multiple of 4
    (2, 10, 10, 128, 128, 3, 1, 2 | 16, True, True),  # post-activation residual (CSPRepLayer add)
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_simt(case, dtype):
    B, H, W, Cin, Cout, k, stride, act, use_res, use_scale = case
    x = rnd((B, H, W, Cin), dtype, 3)
    w = rnd((Cout, k, k, Cin), dtype, 4, 1.0 / math.sqrt(k * k * Cin))
    sc = (torch.rand(Cout) + 0.5) if use_scale else None
    bi = rnd((Cout,), torch.float32, 5, 0.2)
    pad = (k - 1) // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    res = rnd((B, Ho, Wo, Cout), dtype, 6) if use_res else None
    ref = torch.empty((B, Ho, Wo, Cout), dtype=dtype)
    REF.conv2d(x, w, sc, bi, stride, pad, act, res, ref, 0)
    out = ops.conv2d(x.to(DEV), w.to(DEV), None if sc is None else sc.to(DEV), bi.to(DEV), stride=stride, pad=pad, act=act,
                     residual=None if res is None else res.to(DEV), algo=ops.ALGO_SIMT)
    close(out, ref, dtype, f"conv {case}")


def test_conv2d_slices_and_mixed_dtype():
    # input = channel slice of a wider buffer, output = slice of a concat buffer, fp16 in / fp32 out
    xb = rnd((2, 8, 8, 96), torch.float16, 7)
    w = rnd((32, 1, 1, 64), torch.float16, 8, 0.1)
    bi = rnd((32,), torch.float32, 9)
    ref = torch.empty((2, 8, 8, 32), dtype=torch.float32)
    REF.conv2d(xb[..., 32:], w, None, bi, 1, 0, 0, None, ref, 0)
    xg = xb.to(DEV)
    ob = torch.zeros((2, 8, 8, 64), dtype=torch.float32, device=DEV)
    ops.conv2d(xg[..., 32:], w.to(DEV), None, bi.to(DEV), out=ob[..., 32:], algo=ops.ALGO_SIMT)
    close(ob[..., 32:], ref, torch.float16, "slices")
    assert float(ob[..., :32].abs().max()) == 0.0
    # batch-strided output (one level of the [B, S, C] memory)
    mem = torch.zeros((2, 100, 32), dtype=torch.float32, device=DEV)
    ops.conv2d(xg[..., 32:], w.to(DEV), None, bi.to(DEV), out=mem[:, 20:84].unflatten(1, (8, 8)), algo=ops.ALGO_SIMT)
    close(mem[:, 20:84].reshape(2, 8, 8, 32), ref, torch.float16, "batch-strided")
    assert float(mem[:, :20].abs().max()) == 0.0 and float(mem[:, 84:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_pools_resize_add(dtype):
    x = rnd((2, 21, 30, 64), dtype, 10)
    for name, shape in (("maxpool3x3s2", (2, 11, 15, 64)), ("avgpool2x2", (2, 11, 15, 64))):
        ref = torch.empty(shape, dtype=dtype)
        getattr(REF, name)(x, ref)
        close(getattr(ops, name)(x.to(DEV)), ref, dtype, name)
    for size in ((42, 60), (10, 15), (33, 17)):
        ref = torch.empty((2, size[0], size[1], 64), dtype=dtype)
        REF.resize_bilinear(x, ref)
        close(ops.resize_bilinear(x.to(DEV), size), ref, dtype, f"resize {size}")
    buf = torch.zeros((2, 42, 60, 128), dtype=dtype, device=DEV)
    ops.resize_bilinear(x.to(DEV), (42, 60), out=buf[..., 64:])
    ref = torch.empty((2, 42, 60, 64), dtype=dtype)
    REF.resize_bilinear(x, ref)
    close(buf[..., 64:], ref, dtype, "resize into slice")
    a, b = rnd((3, 50, 64), dtype, 11), rnd((50, 64), dtype, 12)
    ref = torch.empty_like(a)
    REF.add(a, b, ref)
    close(ops.add(a.to(DEV), b.to(DEV)), ref, dtype, "add bcast")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_layernorm_attention(dtype):
    x, r = rnd((5, 300, 256), dtype, 13, 3.0), rnd((5, 300, 256), dtype, 14)
    g, b = torch.rand(256) + 0.5, rnd((256,), torch.float32, 15, 0.1)
    ref = torch.empty_like(x)
    REF.layernorm(x, r, g, b, ref, 1e-5)
    close(ops.layernorm(x.to(DEV), g.to(DEV), b.to(DEV), residual=r.to(DEV)), ref, dtype, "layernorm+res")
    x2 = rnd((7, 1024), dtype, 16)
    g2, b2 = torch.rand(1024) + 0.5, rnd((1024,), torch.float32, 17, 0.1)
    ref = torch.empty_like(x2)
    REF.layernorm(x2, None, g2, b2, ref, 1e-5)
    close(ops.layernorm(x2.to(DEV), g2.to(DEV), b2.to(DEV)), ref, dtype, "layernorm 1024")
    for L in (300, 400, 37):
        qk, v = rnd((2, L, 512), dtype, 18), rnd((2, L, 256), dtype, 19)
        ref = torch.empty((2, L, 256), dtype=dtype)
        REF.attention(qk[..., :256], qk[..., 256:], v, ref, 8, 1 / math.sqrt(32))
        qg = qk.to(DEV)
        close(ops.attention(qg[..., :256], qg[..., 256:], v.to(DEV), 8, 1 / math.sqrt(32)), ref, dtype, f"attention L={L}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_msda(dtype):
    shapes = [(5, 7), (10, 14), (20, 28)]
    S = sum(h * w for h, w in shapes)
    B, Q, heads, P = 2, 50, 8, 4
    val_all = rnd((B, S, 512), dtype, 20)
    oa = rnd((B, Q, heads * 3 * P * 3), torch.float32, 21, 2.0)
    g = torch.Generator().manual_seed(22)
    ref_pts = torch.cat([torch.rand((B, Q, 2), generator=g) * 1.2 - 0.1, torch.rand((B, Q, 2), generator=g) * 0.5], -1)
    ref = torch.empty((B, Q, 256), dtype=dtype)
    REF.msda(val_all[..., 256:], oa, ref_pts, shapes, P, heads, ref)
    vg = val_all.to(DEV)
    out = ops.msda(vg[..., 256:], oa.to(DEV), ref_pts.to(DEV), shapes, P, heads)
    close(out, ref, dtype, "msda")


def test_selection_ops():
    x = rnd((2, 40, 64), torch.float32, 23)
    valid = (torch.arange(40) % 3 != 0).to(torch.uint8)
    fill = rnd((64,), torch.float32, 24)
    ref = torch.empty_like(x)
    REF.row_select(x, valid, fill, ref)
    close(ops.row_select(x.to(DEV), valid.to(DEV), fill.to(DEV)), ref, torch.float32, "row_select")
    buf = rnd((3, 100, 368), torch.float32, 25)
    ref = torch.empty((3, 100))
    REF.rowmax(buf[..., :365], ref)
    close(ops.rowmax(buf.to(DEV)[..., :365]), ref, torch.float32, "rowmax")
    src = rnd((2, 90, 256), torch.float16, 26)
    idx = torch.stack([torch.randperm(90, generator=torch.Generator().manual_seed(s))[:30] for s in (1, 2)]).to(torch.int32)
    ref = torch.empty((2, 30, 256), dtype=torch.float16)
    REF.gather_rows(src, idx, ref)
    assert torch.equal(ops.gather_rows(src.to(DEV), idx.to(DEV)).cpu(), ref)


@pytest.mark.parametrize("N,K", [(8400, 300), (109500, 300), (1000, 1), (512, 512), (300, 17)])
def test_topk_exact(N, K):
    g = torch.Generator().manual_seed(N + K)
    x = torch.randn((3, N), generator=g)
    x[0, ::7] = x[0, 3]          # many exact ties, some straddling the K-th value
    x[1] = torch.round(x[1] * 4) / 4   # heavy quantisation: long tie runs
    x[2, 5:50] = -0.0
    x[2, 60:90] = 0.0
    ri, rv = torch.empty((3, K), dtype=torch.int32), torch.empty((3, K))
    REF.topk(x, K, ri, rv)
    v, i = ops.topk(x.to(DEV), K)
    assert torch.equal(v.cpu(), rv)
    # -0.0 and +0.0 compare equal in torch.sort but have distinct keys here: compare indices where values are non-zero
    nz = rv != 0
    assert torch.equal(i.cpu()[nz], ri[nz])


def test_box_ops():
    x = rnd((2, 300, 4), torch.float32, 30, 2.0)
    r = torch.rand((2, 300, 4), generator=torch.Generator().manual_seed(31))
    r[0, 0] = torch.tensor([0.0, 1.0, 1e-7, 0.5])
    for mode, ref_in in ((0, None), (1, r), (3, None)):
        ref = torch.empty_like(x)
        REF.box_op(mode, x, ref_in, None, ref)
        fn = {0: lambda: ops.box_sigmoid(x.to(DEV)), 1: lambda: ops.box_refine(x.to(DEV), r.to(DEV)), 3: lambda: ops.box_cxcywh_to_xyxy(x.to(DEV))}[mode]
        close(fn(), ref, torch.float32, f"box_op {mode}")
    anchors = rnd((8400, 4), torch.float32, 32)
    idx = torch.randint(0, 8400, (2, 300), generator=torch.Generator().manual_seed(33), dtype=torch.int32)
    ref = torch.empty_like(x)
    REF.box_op(2, x, anchors, idx, ref)
    assert torch.equal(ops.box_add_anchors(x.to(DEV), anchors.to(DEV), idx.to(DEV)).cpu(), ref)


def test_detr_postprocess_vs_ref():
    g = torch.Generator().manual_seed(40)
    B, Q, C, K = 3, 300, 365, 300
    scores = torch.sigmoid(torch.randn((B, Q, C), generator=g) * 2 - 4)
    cxcywh = torch.rand((B, Q, 4), generator=g)
    boxes = torch.stack([cxcywh[..., 0] - cxcywh[..., 2] / 2, cxcywh[..., 1] - cxcywh[..., 3] / 2, cxcywh[..., 0] + cxcywh[..., 2] / 2, cxcywh[..., 1] + cxcywh[..., 3] / 2], -1)
    sizes = torch.tensor([[375, 500], [720, 1280], [640, 640]], dtype=torch.int32)
    outs_ref = (torch.empty((B, K)), torch.empty((B, K), dtype=torch.int32), torch.empty((B, K, 4), dtype=torch.int32), torch.empty((B, K), dtype=torch.int32),
                torch.empty((B,), dtype=torch.int32))
    REF.detr_postprocess(scores, boxes, sizes, K, 0.4, *outs_ref)
    outs = ops.detr_postprocess(scores.to(DEV), boxes.to(DEV), sizes.to(DEV), K, 0.4)
    for a, b, name in zip(outs, outs_ref, ("scores", "labels", "boxes", "query", "count")):
        assert torch.equal(a.cpu(), b), name


def test_stem_conv_u8_nhwc_matches_float_path():
    g = torch.Generator().manual_seed(77)
    u8 = torch.randint(0, 256, (2, 64, 96, 3), generator=g, dtype=torch.uint8)
    w = rnd((32, 3, 3, 3), torch.float32, 1, 0.3)
    sc, bi = torch.rand(32) + 0.5, rnd((32,), torch.float32, 2, 0.1)
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    a = ops.stem_conv(u8.to(DEV), w.to(DEV), sc.to(DEV), bi.to(DEV), mean, std, 1, torch.float32)
    b = ops.stem_conv(u8.permute(0, 3, 1, 2).float().contiguous().to(DEV), w.to(DEV), sc.to(DEV), bi.to(DEV), mean, std, 1, torch.float32)
    assert torch.equal(a, b)


@pytest.mark.parametrize("B,Lq,Lk,heads", [(2, 400, 400, 8), (3, 300, 300, 8), (2, 100, 37, 4), (1, 65, 129, 2)])
def test_attention_split_precision(B, Lq, Lk, heads):
    """fp32 attention on the tensor cores (hi/lo fp16 splits of Q, K, V and of the softmax numerators) vs the fp32 reference, column-sliced inputs."""
    C = heads * 32
    if Lq == Lk:  # q and k as column slices of one fused projection buffer (how the engines call it; batch stride = L * pitch)
        qk = rnd((B, Lq, 2 * C), torch.float32, 1, 1.5)
        q, k = qk[..., :C], qk[..., C:]
    else:
        q, k = rnd((B, Lq, C), torch.float32, 1, 1.5), rnd((B, Lk, C), torch.float32, 3, 1.5)
    v = rnd((B, Lk, C), torch.float32, 2)
    scale = 1.0 / math.sqrt(32)
    ref = torch.empty((B, Lq, C))
    REF.attention(q, k, v, ref, heads, scale)
    if Lq == Lk:
        qd = qk.to(DEV)
        qg, kg = qd[..., :C], qd[..., C:]
    else:
        qg, kg = q.to(DEV), k.to(DEV)
    out = ops.attention(qg, kg, v.to(DEV), heads, scale, split=True)
    err = float((out.cpu() - ref).abs().max())
    assert err <= 2e-5 * max(1.0, float(ref.abs().max())), err
    # pair rows for the out_proj linear: exactly the split of the fp32 result
    op = ops.attention(qg, kg, v.to(DEV), heads, scale, split=True, out_pair=True)
    assert torch.equal(op.buf, ops.split_pair(out))


def _pair_eq(p, f32):
    """a Pair holds exactly the [hi | lo] split of an fp32 tensor"""
    return torch.equal(p.buf, ops.split_pair(f32.contiguous()))


def test_fused_row_glue_equals_the_separate_launches():
    """csrc/head_fused.cu against the launches it replaces (bit for bit): layernorm_ex (mask fill, top-k gather, residual, fp32 / pair / pair + pos outputs),
    split_pair_ex (GELU, + pos), box_refine_qpos (refinement + query_pos layer 0), sigmoid_rows (pitched), msda pair rows."""
    B, S, C, K = 3, 500, 256, 40
    x = rnd((B, S, C), torch.float32, 1, 2.0).to(DEV)
    res = rnd((B, S, C), torch.float32, 2).to(DEV)
    g, b_ = (torch.rand(C) + 0.5).to(DEV), rnd((C,), torch.float32, 3, 0.2).to(DEV)
    valid = (torch.rand(S, generator=torch.Generator().manual_seed(4)) > 0.2).to(torch.uint8).to(DEV)
    fill = rnd((C,), torch.float32, 5).to(DEV)
    pos = rnd((B, S, C), torch.float32, 6).to(DEV)
    # plain LayerNorm(x + res): fp32, pair, pair(y + pos)
    y_ref = ops.layernorm(x + res, g, b_)
    y, yp, ypp = ops.layernorm_ex(x, g, b_, residual=res, pos=pos, want_pair_pos=True)
    assert torch.equal(y, y_ref) and _pair_eq(yp, y_ref) and _pair_eq(ypp, ops.add(y_ref, pos))
    # masked rows (memory * valid_mask behind enc_output.0) and gathered rows
    sel = ops.row_select(x, valid, fill)
    full = ops.layernorm(sel, g, b_)
    _, fp, _ = ops.layernorm_ex(x, g, b_, valid=valid, fill=fill, want_f32=False)
    assert _pair_eq(fp, full)
    idx = torch.stack([torch.randperm(S, generator=torch.Generator().manual_seed(7 + i))[:K] for i in range(B)]).to(torch.int32).to(DEV)
    gat = ops.gather_rows(full, idx)
    gy, gp, _ = ops.layernorm_ex(x, g, b_, gather=idx, valid=valid, fill=fill)
    assert torch.equal(gy, gat) and _pair_eq(gp, gat)
    # broadcast positional term (AIFI: pos [L, C] for every image), GELU
    pos1 = rnd((S, C), torch.float32, 8).to(DEV)
    sp, spp = ops.split_pair_ex(x, pos=pos1, want_pair=True, want_pair_pos=True)
    assert _pair_eq(sp, x) and _pair_eq(spp, ops.add(x, pos1))
    hp, _ = ops.split_pair_ex(x, act=ops.ACT_GELU)
    ref_gelu = torch.nn.functional.gelu(x.cpu())
    assert float((hp.float().cpu() - ref_gelu).abs().max()) < 1e-5
    # box refinement + query_pos_head layer 0
    M = 700
    ref = torch.rand((2, M // 2, 4), generator=torch.Generator().manual_seed(9)).to(DEV)
    delta = rnd((2, M // 2, 4), torch.float32, 10, 0.5).to(DEV)
    w0, b0 = rnd((512, 4), torch.float32, 11, 0.7).to(DEV), rnd((512,), torch.float32, 12, 0.3).to(DEV)
    new_ref, qp = ops.box_refine_qpos(delta, ref, w0, b0)
    want_ref = ops.box_refine(delta, ref)
    assert torch.equal(new_ref, want_ref)
    want_q = ops.linear(want_ref, w0, b0, act=ops.ACT_RELU, out_dtype=torch.float32, algo=ops.ALGO_SIMT)
    assert _pair_eq(qp, want_q)
    same_ref, qp0 = ops.box_refine_qpos(None, ref, w0, b0)
    assert same_ref is ref or torch.equal(same_ref, ref)
    assert _pair_eq(qp0, ops.linear(ref, w0, b0, act=ops.ACT_RELU, out_dtype=torch.float32, algo=ops.ALGO_SIMT))
    only_ref, none = ops.box_refine_qpos(delta, ref)
    assert none is None and torch.equal(only_ref, want_ref)
    # pitched sigmoid
    buf = rnd((2, 300, 368), torch.float32, 13, 3.0).to(DEV)
    sg = ops.sigmoid_rows(buf[..., :365])
    assert sg.is_contiguous() and torch.equal(sg, ops.box_sigmoid(buf[..., :365].contiguous()))


def test_msda_pair_rows_equal_the_split_fp32_output():
    shapes = [(20, 20), (40, 40), (80, 80)]
    Sv = sum(h * w for h, w in shapes)
    B, Q, heads, P = 2, 300, 8, 4
    value = rnd((B, Sv, 6 * 256), torch.float32, 1).to(DEV)[..., 256:512]   # a layer's column slice of the fused value projection
    oa = rnd((B, Q, heads * 3 * P * 3), torch.float32, 2).to(DEV)
    ref = torch.rand((B, Q, 4), generator=torch.Generator().manual_seed(3)).to(DEV)
    out = ops.msda(value, oa, ref, shapes, P, heads)
    op = ops.msda(value, oa, ref, shapes, P, heads, out_pair=True)
    assert _pair_eq(op, out)


@pytest.mark.parametrize("B,H,W,size", [(3, 480, 640, (640, 640)), (2, 720, 1280, (640, 640)), (1, 333, 500, (800, 800)), (2, 64, 48, (96, 160))])
def test_image_resize_matches_interpolate(B, H, W, size):
    """fb200_image_resize (the pre-processing resize of a whole batch in one launch) vs F.interpolate(bilinear, align_corners=False) on the float image - what the
    reference runs (processor/base_processor.py:284-294) - for uint8 NHWC and float NCHW inputs, up- and down-scaling."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    ref = F.interpolate(u8.permute(0, 3, 1, 2).float(), size=size, mode="bilinear", align_corners=False)
    got = ops.image_resize(u8.cuda(), size)
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert float((got.cpu() - ref).abs().max()) <= 2e-4, float((got.cpu() - ref).abs().max())
    f32 = u8.permute(0, 3, 1, 2).float().contiguous()
    got2 = ops.image_resize(f32.cuda(), size)
    assert torch.equal(got2, got), "uint8 NHWC and float NCHW inputs of the same image must give the same result"
