"""GPU: the 2-CTA (cta_group::2) all-TMA GEMM against fp64 and against the 1-CTA engine, on the shapes that stress a pair
kernel: M / N tails, a single K-block (fewer than the pipeline depth), more tiles than pairs (accumulator reuse, uneven
work per pair), fewer tiles than pairs, every epilogue incl. the split output, run-to-run determinism."""
import math

import pytest
import torch
import torch.nn.functional as F

from controlvar_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def g(t):
    return t.to(DEV).contiguous()


def err(a, ref):
    return ((a.double() - ref).abs().max() / ref.abs().max()).item()


def split(t):
    s = ops.SplitWeight(t)           # elementwise TF32 hi/lo split; the same op the producers apply to activations
    return s.hi, s.lo


@pytest.fixture
def engine3():
    old = ops.set_gemm_engine(ops.ENGINE_TC_2CTA)
    yield
    ops.set_gemm_engine(old)


@pytest.mark.parametrize("M,N,K", [(256, 256, 32), (256, 256, 96), (300, 1536, 1536), (1000, 768, 3072), (512, 1920, 1920),
                                   (4096, 1536, 6144), (2304, 4096, 768), (37 * 256, 1536, 256), (65536, 1536, 1536)])
def test_tc2_gemm_vs_fp64_and_1cta(engine3, M, N, K):
    torch.manual_seed(M + N + K)
    A, W, b = torch.randn(M, K), torch.randn(N, K) / math.sqrt(K), torch.randn(N)
    ref = A.double() @ W.double().T + b.double()
    Ag, Wg, bg = g(A), ops.SplitWeight(g(W)), g(b)
    A_hi, A_lo = split(Ag)
    assert torch.equal(A_hi + A_lo, Ag)                      # the split is exact
    n0 = ops.launch_count()
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(A_hi, Wg, bg, out, M, N, K, A_lo=A_lo)
    torch.cuda.synchronize()
    assert ops.launch_count() - n0 == 1
    e2 = err(out.cpu(), ref)
    out_b = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(A_hi, Wg, bg, out_b, M, N, K, A_lo=A_lo)
    assert torch.equal(out, out_b), "2-CTA GEMM is not deterministic run to run"
    ops.set_gemm_engine(ops.ENGINE_TC_3XTF32)
    out1 = torch.empty(M, N, device=DEV)
    ops.gemm(Ag, Wg, bg, out1, M, N, K)
    ops.set_gemm_engine(ops.ENGINE_TC_2CTA)
    e1 = err(out1.cpu(), ref)
    d12 = (out - out1).abs().max().item() / ref.abs().max().item()
    print(f"\n[tc2-accuracy] M={M} N={N} K={K}: 2-CTA err {e2:.3e}  1-CTA err {e1:.3e}  |2CTA-1CTA| {d12:.2e}")
    assert not torch.isnan(out).any(), "some output elements were never written"
    assert e2 < 2e-5


def test_tc2_epilogues(engine3):
    torch.manual_seed(3)
    R, l, C, K = 4, 128, 512, 1024
    M = R * l
    A, Wt, b = torch.randn(M, K), torch.randn(C, K) / math.sqrt(K), torch.randn(C)
    x0, ada = torch.randn(M, C), torch.randn(R, 6 * C)
    A_hi, A_lo = split(g(A))
    Wg = ops.SplitWeight(g(Wt))
    ref_lin = A.double() @ Wt.double().T + b.double()
    x, ada_g = g(x0), g(ada)
    ops.gemm(A_hi, Wg, g(b), x, M, C, K, A_lo=A_lo, epilogue=ops.EPI_BIAS_GAMMA_RESID, gamma=ada_g[:, C:2 * C],
             gamma_row_stride=6 * C, rows_per_sample=l)
    assert err(x.cpu(), x0.double() + ref_lin * ada[:, C:2 * C].double().repeat_interleave(l, 0)) < 1e-5
    hi, lo = torch.empty(M, C, device=DEV), torch.empty(M, C, device=DEV)
    ops.gemm(A_hi, Wg, g(b), hi, M, C, K, A_lo=A_lo, epilogue=ops.EPI_BIAS_GELU, out_lo=lo)
    assert err((hi + lo).cpu(), F.gelu(ref_lin, approximate="tanh")) < 1e-5
    assert (hi.view(torch.int32) & 0x1FFF).abs().max().item() == 0, "hi must have its 13 low mantissa bits clear"
    assert (lo.abs() <= hi.abs() * 2.0 ** -10 + 1e-30).all()


def test_tc2_throughput(engine3):
    """Sanity only (bench.py is the benchmark): the pair kernel should not be slower than the 1-CTA engine."""
    M, N, K = 65536, 6144, 1536
    A = torch.randn(M, K, device=DEV)
    A_hi, A_lo = split(A)
    W, b = ops.SplitWeight(torch.randn(N, K, device=DEV) / 40), torch.randn(N, device=DEV)
    out = torch.empty(M, N, device=DEV)
    res = {}
    for eng in (ops.ENGINE_TC_2CTA, ops.ENGINE_TC_3XTF32):
        ops.set_gemm_engine(eng)
        fn = (lambda: ops.gemm(A_hi, W, b, out, M, N, K, A_lo=A_lo)) if eng == ops.ENGINE_TC_2CTA else \
             (lambda: ops.gemm(A, W, b, out, M, N, K))
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[eng] = 10 * 2.0 * M * N * K / (e0.elapsed_time(e1) * 1e-3) / 1e12
    ops.set_gemm_engine(ops.ENGINE_TC_2CTA)
    print(f"\n[tc2-throughput] M={M} N={N} K={K}: 2-CTA {res[3]:.1f} TFLOP/s   1-CTA {res[1]:.1f} TFLOP/s")
