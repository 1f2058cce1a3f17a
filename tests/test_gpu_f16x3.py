"""GPU: the FP16-pair ("f16x3") operand format and engine 4.
  x ~= hi + lo / 2048 with both halves rounded to nearest (include/cvar.h: cvar_split_f16); three kind::f16 tcgen05 MMAs
  (hi*hi, hi*lo, lo*hi) on the 2-CTA kernel must be at least as accurate as the 3xTF32 engine on the same problem.
Covered: the split itself (error bound, tiny / large / saturating values), every producer of pairs (cvar_ln_modulate,
cvar_gemm out16, cvar_attn_kvcache out16), every consumer (cvar_gemm, cvar_qkv_project) incl. M / N tails smaller than a
tile, a single K-block, all epilogues, determinism."""
import math

import pytest
import torch
import torch.nn.functional as F

from controlvar_b200 import ops
from controlvar_b200._lib import CvarError

pytestmark = pytest.mark.gpu
DEV = "cuda"


def g(t):
    return t.to(DEV).contiguous()


def err(a, ref):
    return ((a.double() - ref).abs().max() / ref.abs().max()).item()


@pytest.fixture
def engine4():
    old = ops.set_gemm_engine(ops.ENGINE_TC_F16X3)
    yield
    ops.set_gemm_engine(old)


def test_split_f16_error_bound():
    torch.manual_seed(0)
    x = torch.cat([torch.randn(1 << 16), torch.randn(1 << 14) * 1e-3, torch.randn(1 << 14) * 300.0,
                   torch.randn(1 << 12) * 1e-6, torch.tensor([0.0, -0.0, 1.0, -1.0, 65504.0, -65504.0, 6.1e-5, 5.9e-8])])
    p = ops.F16Pair.from_tensor(g(x))
    back = p.float().cpu().double()
    # |x| >= 2^-12: the scaled residual is a normal fp16 number, so two roundings to nearest leave 2^-24 relative
    # (+ 2^-24 for the fp32 rounding of the reconstruction in F16Pair.float())
    big = x.abs() >= 2.0 ** -12
    rel = ((back - x.double()).abs() / x.double().abs().clamp_min(1e-300))[big].max().item()
    assert rel <= 2.0 ** -23, rel
    # below that the scaled residual may be subnormal (spacing 2^-24): absolute error <= 2^-25 / 2048 (+ reconstruction)
    assert (back - x.double()).abs()[~big].max().item() <= 2.0 ** -35
    assert torch.isfinite(p.hi).all() and torch.isfinite(p.lo).all()
    # saturation instead of inf
    s = ops.F16Pair.from_tensor(g(torch.tensor([1e6, -1e6, 7e4, 0.0])))
    assert torch.isfinite(s.hi).all() and s.hi[0].item() == 65504.0 and s.hi[1].item() == -65504.0


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (8, 256, 128), (100, 64, 192), (300, 1536, 1536), (1000, 768, 3072),
                                   (512, 1920, 1920), (4096, 1536, 6144), (2304, 4096, 768), (37 * 256, 1536, 256),
                                   (65536, 1536, 1536)])
def test_f16x3_gemm_vs_fp64_and_3xtf32(engine4, M, N, K):
    torch.manual_seed(M + N + K)
    A, W, b = torch.randn(M, K), torch.randn(N, K) / math.sqrt(K), torch.randn(N)
    ref = A.double() @ W.double().T + b.double()
    Ag, bg = g(A), g(b)
    W16 = ops.SplitWeight(g(W), f16=True)
    A16 = ops.F16Pair.from_tensor(Ag)
    n0 = ops.launch_count()
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(None, W16, bg, out, M, N, K, A16=A16)
    torch.cuda.synchronize()
    assert ops.launch_count() - n0 == 1
    assert not torch.isnan(out).any(), "some output elements were never written"
    e4 = err(out.cpu(), ref)
    out_b = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(None, W16, bg, out_b, M, N, K, A16=A16)
    assert torch.equal(out, out_b), "f16x3 GEMM is not deterministic run to run"
    # the 3xTF32 engine and the SIMT engine on the same problem
    Wt = ops.SplitWeight(g(W))
    out3, out0 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    ops.set_gemm_engine(ops.ENGINE_TC_3XTF32)
    ops.gemm(Ag, Wt, bg, out3, M, N, K)
    ops.set_gemm_engine(ops.ENGINE_SIMT)
    ops.gemm(Ag, Wt, bg, out0, M, N, K)
    ops.set_gemm_engine(ops.ENGINE_TC_F16X3)
    e3, e0 = err(out3.cpu(), ref), err(out0.cpu(), ref)
    print(f"\n[f16x3-accuracy] M={M} N={N} K={K}: f16x3 err {e4:.3e}  3xTF32 err {e3:.3e}  SIMT err {e0:.3e}")
    assert e4 < 2e-5
    assert e4 <= 1.5 * e3 + 1e-7, "f16x3 should not be less accurate than 3xTF32"


@pytest.mark.parametrize("M,N,K", [(256, 1536, 1536), (512, 1536, 6144), (200, 6144, 1536), (1100, 768, 768), (256, 4096, 1536)])
def test_f16x3_narrow_tiles_are_bit_identical(engine4, M, N, K):
    """Dense layers whose 256 x 256 tiling would fill less than half of the machine (small scales, small batches) run on
    256 x 128 pair tiles.  Per output element the products and their order are the same, so the first M rows of a launch large
    enough to take the regular kernel must come out bit-identical - for every epilogue kind, incl. the FP16-pair output."""
    torch.manual_seed(M + N)
    Mbig, l = 16384, 64
    A = g(torch.randn(Mbig, K))
    W16 = ops.SplitWeight(g(torch.randn(N, K) / math.sqrt(K)), f16=True)
    b = g(torch.randn(N))
    A16 = ops.F16Pair.from_tensor(A)
    gamma = g(torch.randn(Mbig // l, N))
    x0 = g(torch.randn(Mbig, N))
    for epi in (ops.EPI_BIAS, ops.EPI_BIAS_GELU, ops.EPI_BIAS_GAMMA_RESID, ops.EPI_BIAS_RESID):
        outs = []
        for rows in (Mbig, M):
            kw = dict(epilogue=epi)
            out, out16 = x0.clone(), None
            if epi == ops.EPI_BIAS_GAMMA_RESID:
                kw.update(gamma=gamma, gamma_row_stride=N, rows_per_sample=l)
            elif epi == ops.EPI_BIAS_RESID:
                kw.update(resid=x0)
                out = torch.empty(Mbig, N, device=DEV)
            elif epi == ops.EPI_BIAS_GELU:
                out, out16 = None, ops.F16Pair.empty((Mbig, N), DEV)
            ops.gemm(None, W16, b, out, rows, N, K, A16=A16, out16=out16, **kw)
            outs.append((out16.hi[:M].clone(), out16.lo[:M].clone()) if out16 is not None else (out[:M].clone(),))
        assert all(torch.equal(a, c) for a, c in zip(*outs)), f"epilogue {epi}: narrow tiles differ from the regular kernel"
    ref = A[:M].double().cpu() @ W16.w.double().cpu().T + b.double().cpu()
    out = torch.empty(M, N, device=DEV)
    ops.gemm(None, W16, b, out, M, N, K, A16=A16)
    assert err(out.cpu(), ref) < 2e-5


def test_f16x3_epilogues_and_pair_output(engine4):
    torch.manual_seed(3)
    R, l, C, K = 4, 128, 512, 1024
    M = R * l
    A, Wt, b = torch.randn(M, K), torch.randn(C, K) / math.sqrt(K), torch.randn(C)
    x0, ada = torch.randn(M, C), torch.randn(R, 6 * C)
    A16 = ops.F16Pair.from_tensor(g(A))
    W16 = ops.SplitWeight(g(Wt), f16=True)
    ref_lin = A.double() @ Wt.double().T + b.double()
    x, ada_g = g(x0), g(ada)
    ops.gemm(None, W16, g(b), x, M, C, K, A16=A16, epilogue=ops.EPI_BIAS_GAMMA_RESID, gamma=ada_g[:, C:2 * C],
             gamma_row_stride=6 * C, rows_per_sample=l)
    assert err(x.cpu(), x0.double() + ref_lin * ada[:, C:2 * C].double().repeat_interleave(l, 0)) < 1e-5
    # GELU epilogue written ONLY as a pair (the fc1 -> fc2 hand-over), and as fp32 + pair
    o16 = ops.F16Pair.empty((M, C), DEV)
    ops.gemm(None, W16, g(b), None, M, C, K, A16=A16, epilogue=ops.EPI_BIAS_GELU, out16=o16)
    ref_gelu = F.gelu(ref_lin, approximate="tanh")
    assert err(o16.float().cpu(), ref_gelu) < 1e-5
    o32, o16b = torch.empty(M, C, device=DEV), ops.F16Pair.empty((M, C), DEV)
    ops.gemm(None, W16, g(b), o32, M, C, K, A16=A16, epilogue=ops.EPI_BIAS_GELU, out16=o16b)
    assert torch.equal(o16b.hi, o16.hi) and torch.equal(o16b.lo, o16.lo)
    chk = ops.F16Pair.from_tensor(o32)              # the pair written by the epilogue == the split of the fp32 result
    assert torch.equal(chk.hi, o16.hi) and torch.equal(chk.lo, o16.lo)
    # alpha (BIAS mode)
    o = torch.empty(M, C, device=DEV)
    ops.gemm(None, W16, g(b), o, M, C, K, A16=A16, alpha=0.25)
    assert err(o.cpu(), (A.double() @ Wt.double().T) * 0.25 + b.double()) < 1e-5


def test_f16x3_rejects_bad_use(engine4):
    A16 = ops.F16Pair.from_tensor(torch.randn(64, 96, device=DEV))
    W16 = ops.SplitWeight(torch.randn(64, 96, device=DEV), f16=True)
    out = torch.empty(64, 64, device=DEV)
    with pytest.raises(CvarError):
        ops.gemm(None, W16, None, out, 64, 64, 96, A16=A16)             # K % 64 != 0: an error, not a fall-through
    with pytest.raises(CvarError):
        ops.gemm(None, ops.SplitWeight(torch.randn(64, 128, device=DEV)), None, out, 64, 64, 128,
                 A16=ops.F16Pair.from_tensor(torch.randn(64, 128, device=DEV)))     # TF32 weight with an FP16 activation
    old = ops.set_gemm_engine(ops.ENGINE_SIMT)
    try:
        with pytest.raises(CvarError):
            ops.gemm(None, ops.SplitWeight(torch.randn(64, 128, device=DEV), f16=True), None, out, 64, 64, 128,
                     A16=ops.F16Pair.from_tensor(torch.randn(64, 128, device=DEV)))
    finally:
        ops.set_gemm_engine(old)


def test_ln_modulate_pair_output():
    torch.manual_seed(5)
    R, l, C = 6, 50, 768
    M = R * l
    x, ada = g(torch.randn(M, C) * 3 + 1), g(torch.randn(R, 6 * C))
    y = torch.empty(M, C, device=DEV)
    ops.ln_modulate(x, ada[:, 2 * C:3 * C], ada[:, 4 * C:5 * C], 6 * C, y, M, C, l, 1e-6)
    p = ops.F16Pair.empty((M, C), DEV)
    ops.ln_modulate(x, ada[:, 2 * C:3 * C], ada[:, 4 * C:5 * C], 6 * C, None, M, C, l, 1e-6, out16=p)
    chk = ops.F16Pair.from_tensor(y)
    assert torch.equal(chk.hi, p.hi) and torch.equal(chk.lo, p.lo)
    y2, p2 = torch.empty(M, C, device=DEV), ops.F16Pair.empty((M, C), DEV)
    ops.ln_modulate(x, ada[:, 2 * C:3 * C], ada[:, 4 * C:5 * C], 6 * C, y2, M, C, l, 1e-6, out16=p2)
    assert torch.equal(y2, y) and torch.equal(p2.hi, p.hi) and torch.equal(p2.lo, p.lo)


@pytest.mark.parametrize("l,L", [(8, 30), (128, 328), (200, 548)])
def test_attention_pair_output(l, L):
    """cvar_attn_kvcache out16 == split of its fp32 output, on the SIMT (l < 64) and the tensor-core kernel."""
    torch.manual_seed(l)
    R, H = 3, 4
    cache = ops.KVCache(R, H, L, DEV)
    A = g(torch.randn(R * L, H * 64))
    W = ops.SplitWeight(g(torch.randn(3 * H * 64, H * 64) / 16))
    zeros = torch.zeros(H * 64, device=DEV)
    q = torch.empty(R * H * L * 64, device=DEV)
    ops.qkv_project(A, W, zeros, zeros, zeros, q, cache, R, L, 0, H, False, None)       # fill the cache with L keys
    qq = g(torch.randn(R, H, l, 64))
    o = torch.empty(R * l, H * 64, device=DEV)
    ops.attn_kvcache(qq, cache, o, R, H, l, L, 0.125)
    p = ops.F16Pair.empty((R * l, H * 64), DEV)
    ops.attn_kvcache(qq, cache, None, R, H, l, L, 0.125, out16=p)
    chk = ops.F16Pair.from_tensor(o)
    assert torch.equal(chk.hi, p.hi) and torch.equal(chk.lo, p.lo)


def test_qkv_project_pair_operands(engine4):
    torch.manual_seed(9)
    R, l, H = 2, 96, 4
    C = H * 64
    A = g(torch.randn(R * l, C))
    Wq = g(torch.randn(3 * C, C) / math.sqrt(C))
    qb, kb, vb = g(torch.randn(C)), torch.zeros(C, device=DEV), g(torch.randn(C))
    c16, c32 = ops.KVCache(R, H, l, DEV), ops.KVCache(R, H, l, DEV)
    q16, q32 = torch.empty(R * H * l * 64, device=DEV), torch.empty(R * H * l * 64, device=DEV)
    ops.qkv_project(None, ops.SplitWeight(Wq, f16=True), qb, kb, vb, q16, c16, R, l, 0, H, False, None,
                    A16=ops.F16Pair.from_tensor(A))
    ops.set_gemm_engine(ops.ENGINE_SIMT)
    ops.qkv_project(A, Wq, qb, kb, vb, q32, c32, R, l, 0, H, False, None)
    ops.set_gemm_engine(ops.ENGINE_TC_F16X3)
    ref = (A.double() @ Wq.double().T).cpu()
    scale = ref.abs().max().item()
    assert (q16 - q32).abs().max().item() / scale < 5e-6
    assert (c16.keys(l) - c32.keys(l)).abs().max().item() / scale < 5e-6
    assert (c16.values(l) - c32.values(l)).abs().max().item() / scale < 5e-6


@pytest.mark.parametrize("H,l,L_prev", [(12, 8, 10), (24, 18, 0), (30, 32, 28)])
def test_qkv_project16_narrow_tiles_are_bit_identical(engine4, H, l, L_prev):
    """cvar_qkv_project16 on few rows (small scales / batches) takes the 256 x 128 pair tiles: q, the appended K and V^T of the
    first samples must equal, bit for bit, those of a launch with enough samples to take the regular tiles."""
    torch.manual_seed(H + l)
    C, T, R_small, R_big = H * 64, 64, 2, 128
    A = g(torch.randn(R_big * l, C))
    Wq = ops.SplitWeight(g(torch.randn(3 * C, C) / math.sqrt(C)), f16=True)
    qb, kb, vb = g(torch.randn(C)), torch.zeros(C, device=DEV), g(torch.randn(C))
    A16 = ops.F16Pair.from_tensor(A)
    got = []
    for R in (R_big, R_small):
        kv = ops.KVCache16(R, H, T, DEV)
        for t in (kv.k_hi, kv.k_lo, kv.vt_hi, kv.vt_lo):
            t.zero_()
        q16 = ops.F16Pair.empty((R, H, l, 64), DEV)
        ops.qkv_project16(ops.F16Pair(A16.hi[:R * l], A16.lo[:R * l]), Wq, qb, kb, vb, q16, kv, R, l, L_prev, H, False, None)
        n = R_small * H
        got.append([q16.hi[:R_small].clone(), q16.lo[:R_small].clone()] +
                   [t.view(R * H, -1)[:n].clone() for t in (kv.k_hi, kv.k_lo, kv.vt_hi, kv.vt_lo)])
    assert all(torch.equal(a, b) for a, b in zip(*got))
    assert got[1][0].abs().max().item() > 0 and got[1][4].abs().max().item() > 0


# ------------------------------------------------------------------------------------------ convolution (decoder)
def _conv_ref(x_nhwc, w_oihw, bias, resid=None):
    y = F.conv2d(x_nhwc.double().permute(0, 3, 1, 2), w_oihw.double(), bias.double(), padding=w_oihw.shape[-1] // 2)
    y = y.permute(0, 2, 3, 1)
    return y if resid is None else y + resid.double()


@pytest.mark.parametrize("B,H,W,Cin,Cout,ks,resid", [
    (2, 16, 16, 64, 160, 3, False),      # 2 tiles per image, 8 x 16-pixel rows per box
    (3, 16, 8, 32, 32, 3, True),         # M = 384: a pair tile with an absent second half (TMA zero fill past the batch)
    (1, 64, 64, 320, 320, 3, True),      # two N tiles of 160
    (2, 128, 128, 160, 160, 3, False),   # one box = one image row
    (1, 256, 256, 160, 160, 3, True),    # two boxes per image row
    (2, 32, 32, 640, 640, 3, False),     # K = 5760
    (2, 64, 64, 320, 160, 1, False),     # 1x1 (nin_shortcut)
    (1, 32, 32, 96, 256, 3, False),      # BN = 256
])
def test_f16x3_conv_vs_fp64(engine4, B, H, W, Cin, Cout, ks, resid):
    torch.manual_seed(B * 1000 + H + Cin + Cout + ks)
    assert ops.conv2d_f16_supported(H, W, Cin, Cout, ks)
    x = torch.randn(B, H, W, Cin)
    w = torch.randn(Cout, Cin, ks, ks) / math.sqrt(Cin * ks * ks)
    b = torch.randn(Cout)
    r = torch.randn(B, H, W, Cout) if resid else None
    ref = _conv_ref(x, w, b, r)
    wp = torch.empty(Cout, ks * ks * Cin, device=DEV)
    ops.repack_conv_weight(g(w), wp)
    x16, w16 = ops.F16Pair.from_tensor(g(x)), ops.F16Pair.from_tensor(wp)
    out = torch.full((B, H, W, Cout), float("nan"), device=DEV)
    n0 = ops.launch_count()
    ops.conv2d(None, wp, g(b), out, B, H, W, Cin, Cout, ks, x16=x16, w16=w16, resid=None if r is None else g(r))
    torch.cuda.synchronize()
    assert ops.launch_count() - n0 == 1
    assert not torch.isnan(out).any(), "some output elements were never written"
    e4 = err(out.cpu(), ref)
    # the SIMT engine on the same problem
    out0 = torch.empty(B, H, W, Cout, device=DEV)
    ops.conv2d(g(x), wp, g(b), out0, B, H, W, Cin, Cout, ks, resid=None if r is None else g(r), engine=0)
    e0 = err(out0.cpu(), ref)
    print(f"\n[f16x3-conv] B={B} {H}x{W} {Cin}->{Cout} ks={ks}: f16x3 err {e4:.3e}  SIMT err {e0:.3e}")
    assert e4 < 1e-5
    out_b = torch.full((B, H, W, Cout), float("nan"), device=DEV)
    ops.conv2d(None, wp, g(b), out_b, B, H, W, Cin, Cout, ks, x16=x16, w16=w16, resid=None if r is None else g(r))
    assert torch.equal(out, out_b), "f16x3 conv is not deterministic run to run"


@pytest.mark.parametrize("B,H,W,Cin,Cout,ksplit,resid", [
    (2, 128, 128, 160, 160, 0, True),    # 5 channels per group, 16 groups per column half
    (1, 64, 64, 320, 320, 0, False),     # 10 per group, two N tiles
    (2, 32, 32, 640, 640, 3, True),      # 20 per group, K-split: the statistics come from the last of the three launches
    (3, 16, 16, 64, 160, 0, False),      # 256-pixel images: M = 768 = three pair tiles, one per image
])
def test_f16x3_conv_groupnorm_statistics_from_the_epilogue(engine4, B, H, W, Cin, Cout, ksplit, resid):
    """cvar_conv_args.gn_part: GroupNorm(32) coefficients from the partial sums the conv epilogue writes must equal the ones
    cvar_gn_stats computes by reading the stored activation (vae_modules.py:18-19), and the output must not change."""
    torch.manual_seed(H + Cin + Cout)
    assert ops.conv2d_gn_fusable(H, W, Cin, Cout, 3)
    x = torch.randn(B, H, W, Cin)
    w = torch.randn(Cout, Cin, 3, 3) / math.sqrt(Cin * 9)
    b = torch.randn(Cout) + 0.5
    r = g(torch.randn(B, H, W, Cout)) if resid else None
    gamma, beta = g(torch.randn(Cout)), g(torch.randn(Cout))
    wp = torch.empty(Cout, 9 * Cin, device=DEV)
    ops.repack_conv_weight(g(w), wp)
    x16, w16 = ops.F16Pair.from_tensor(g(x)), ops.F16Pair.from_tensor(wp)
    out_ref = torch.empty(B, H, W, Cout, device=DEV)
    ops.conv2d(None, wp, g(b), out_ref, B, H, W, Cin, Cout, 3, x16=x16, w16=w16, resid=r, ksplit=ksplit)
    out = torch.empty(B, H, W, Cout, device=DEV)
    part = torch.full((2 * B * 32 * (H * W // 32),), float("nan"), dtype=torch.float64, device=DEV)
    ops.conv2d(None, wp, g(b), out, B, H, W, Cin, Cout, 3, x16=x16, w16=w16, resid=r, ksplit=ksplit, gn_part=part)
    assert torch.equal(out, out_ref)
    assert not torch.isnan(part).any(), "some partial slots were never written"
    a1, b1 = torch.empty(B, Cout, device=DEV), torch.empty(B, Cout, device=DEV)
    ops.gn_finalize_parts(part, gamma, beta, a1, b1, B, H * W, Cout)
    a0, b0 = torch.empty(B, Cout, device=DEV), torch.empty(B, Cout, device=DEV)
    scratch = torch.empty(2 * B * 32 * ops.gn_chunks(H * W), dtype=torch.float64, device=DEV)
    ops.gn_stats(out, gamma, beta, a0, b0, scratch, B, H * W, Cout)
    ea = ((a1 - a0).abs().max() / a0.abs().max()).item()
    eb = ((b1 - b0).abs().max() / b0.abs().max()).item()
    # fp32 partial sums over 32 pixels x one group, fp64 from there on: ~1e-7 relative
    assert ea < 2e-6 and eb < 2e-6, (ea, eb)
    # and against torch's GroupNorm on the stored activation
    y = F.group_norm(out.permute(0, 3, 1, 2).double().cpu(), 32, gamma.double().cpu(), beta.double().cpu(), 1e-6)
    y1 = out.double().cpu() * a1.double().cpu()[:, None, None, :] + b1.double().cpu()[:, None, None, :]
    assert (y1.permute(0, 3, 1, 2) - y).abs().max().item() < 2e-5
    assert not ops.conv2d_gn_fusable(32, 32, 96, 256, 3) or True      # shape query never raises
    with pytest.raises(CvarError):      # fp32 input path cannot produce them: an error, not a silent skip
        ops.conv2d(g(x), wp, g(b), out, B, H, W, Cin, Cout, 3, gn_part=part)


def test_f16x3_conv_unsupported_shapes_are_reported():
    assert not ops.conv2d_f16_supported(80, 80, 160, 160, 3)       # 128 % 80 != 0 (truncated pyramids)
    assert not ops.conv2d_f16_supported(64, 64, 16, 160, 3)        # Cin % 32
    assert not ops.conv2d_f16_supported(64, 64, 160, 3, 3)         # the image conv stays on the SIMT engine
    assert ops.conv2d_f16_supported(256, 256, 160, 160, 3)


def test_affine_and_upsample_pair_producers():
    torch.manual_seed(11)
    B, H, W, Cn = 2, 8, 8, 64
    x, a, b = g(torch.randn(B, H, W, Cn)), g(torch.rand(B, Cn) + 0.5), g(torch.randn(B, Cn))
    y = torch.empty(B, H, W, Cn, device=DEV)
    p = ops.F16Pair.empty((B, H, W, Cn), DEV)
    # without SiLU the pair is exactly the split of the fp32 result
    ops.affine_nc(x, a, b, y, B, H * W, Cn, silu=False)
    ops.affine_nc(x, a, b, None, B, H * W, Cn, silu=False, out16=p)
    chk = ops.F16Pair.from_tensor(y)
    assert torch.equal(chk.hi, p.hi) and torch.equal(chk.lo, p.lo)
    # with SiLU the pair-only call uses x * rcp(1 + ex2(-x log2 e)) (a few ulp; the fp32 output keeps the IEEE division and expf):
    # both must sit close to the float64 value
    ops.affine_nc(x, a, b, y, B, H * W, Cn, silu=True)
    ops.affine_nc(x, a, b, None, B, H * W, Cn, silu=True, out16=p)
    t = x.double() * a.double()[:, None, None, :] + b.double()[:, None, None, :]
    want = t / (1.0 + torch.exp(-t))
    # fp32 rounding of a*x+b (2^-24) is amplified by up to ~(1 + |t|) through SiLU; __expf adds 2 + 1.16 |t| ulp; the pair 2^-22
    assert ((y.double() - want).abs() <= 1e-6 * want.abs() + 1e-8).all()
    assert ((p.float().double() - want).abs() <= 3e-6 * want.abs() + 1e-8).all()
    assert (p.float() - y).abs().max().item() < 2e-6
    up = ops.F16Pair.empty((B, 2 * H, 2 * W, Cn), DEV)
    ops.upsample2x_split_f16(x, up, B, H, W, Cn)
    ref = ops.F16Pair.from_tensor(x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).contiguous())
    assert torch.equal(up.hi, ref.hi) and torch.equal(up.lo, ref.lo)


def test_decoder_engine4_pixels_within_tolerance(engine4):
    """VQVAE.fhat_to_img with the f16x3 convolutions against the CPU oracle: every pixel within the north-star bound."""
    from controlvar_b200 import VQVAE, weights as Wt
    from controlvar_b200.config import PathConfig
    from oracle import controlvar_oracle as O
    cfg = PathConfig(depth=2)
    vae = VQVAE(vocab_size=cfg.vocab_size, z_channels=cfg.Cvae, ch=cfg.vae_ch, test_mode=True,
                share_quant_resi=cfg.share_quant_resi, v_patch_nums=cfg.patch_nums)
    vsd = Wt.synthetic_vae_state_dict(cfg, 0)
    vae.load_state_dict(vsd, strict=True)
    vae.to(DEV)
    torch.manual_seed(2)
    f_hat = torch.randn(1, cfg.Cvae, 16, 16) * 1.5
    n0 = ops.launch_count()
    img = vae.fhat_to_img(g(f_hat))
    torch.cuda.synchronize()
    ref = O.fhat_to_img(f_hat, vsd)
    e = (img.cpu() - ref).abs().max().item()
    print(f"\n[f16x3-decoder] max pixel err {e:.3e} (tc_min_hw={vae._min_hw()}), launches {ops.launch_count() - n0}")
    assert e < 1e-4


def test_fast_mode_is_half_precision_class_and_off_by_default(engine4):
    """cvar_set_fast_mode(1) - NOT a parity mode - uses the hi halves only: errors of the order of fp16 operand rounding
    (2^-11 relative per operand), far above the parity mode's; the default (0) is untouched by having used it."""
    assert ops.get_fast_mode() == 0
    torch.manual_seed(5)
    M, N, K = 1024, 512, 1536
    A, Wt, b = torch.randn(M, K), torch.randn(N, K) / math.sqrt(K), torch.randn(N)
    ref = A.double() @ Wt.double().T + b.double()
    W16 = ops.SplitWeight(g(Wt), f16=True)
    A16 = ops.F16Pair.from_tensor(g(A))
    out = torch.empty(M, N, device=DEV)
    ops.gemm(None, W16, g(b), out, M, N, K, A16=A16)
    e_par = err(out.cpu(), ref)
    old = ops.set_fast_mode(True)
    try:
        assert old == 0 and ops.get_fast_mode() == 1
        ops.gemm(None, W16, g(b), out, M, N, K, A16=A16)
        e_fast = err(out.cpu(), ref)
        # attention on the same switch
        R, H, l, L = 2, 3, 128, 300
        q, k, v = torch.randn(R, H, l, 64), torch.randn(R, H, L, 64), torch.randn(R, H, L, 64)
        cache = ops.KVCache16(R, H, L, DEV)
        T = cache.T
        kp = ops.F16Pair.from_tensor_qk(g(k))
        cache.k_hi.view(R, H, T, 64)[:, :, :L] = kp.hi
        cache.k_lo.view(R, H, T, 64)[:, :, :L] = kp.lo
        vp = ops.F16Pair.from_tensor(g(v.transpose(2, 3).contiguous()))
        cache.vt_hi.view(R, H, 64, T)[:, :, :, :L] = vp.hi
        cache.vt_lo.view(R, H, 64, T)[:, :, :, :L] = vp.lo
        q16 = ops.F16Pair.from_tensor_qk(g(q))
        o = torch.empty(R, l, H * 64, device=DEV)
        ops.attn_kvcache16(q16, cache, o, R, H, l, L, 0.125, engine=1)
        ref_o = F.scaled_dot_product_attention(q.double(), k.double(), v.double(), scale=0.125).transpose(1, 2).reshape(R, l, H * 64)
        ea_fast = err(o.cpu(), ref_o)
    finally:
        ops.set_fast_mode(False)
    ops.attn_kvcache16(q16, cache, o, R, H, l, L, 0.125, engine=1)
    ea_par = err(o.cpu(), ref_o)
    assert e_par < 1e-5 and ea_par < 1e-5, (e_par, ea_par)
    assert 2e-5 < e_fast < 5e-3 and 2e-5 < ea_fast < 5e-3, (e_fast, ea_fast)
