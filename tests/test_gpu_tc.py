"""GPU: the tcgen05 3xTF32 engine against fp64 references, side by side with the SIMT fp32 engine (the error of the
tensor-core path must stay in the fp32 class, otherwise token parity with the reference is lost)."""
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


@pytest.fixture(params=[32, 16], ids=["bk32", "bk16"])
def tc(request):
    old_e = ops.set_gemm_engine(1)
    old_b = ops.set_tc_kblock(request.param)
    yield request.param
    ops.set_gemm_engine(old_e)
    ops.set_tc_kblock(old_b)


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (128, 128, 64), (256, 256, 256), (300, 1536, 1536), (1000, 768, 3072),
                                   (4096, 1536, 6144), (777, 1920, 1920), (520, 160, 1440), (390, 320, 2880),
                                   (200, 640, 288), (130, 4096, 768)])
def test_tc_gemm_accuracy(tc, M, N, K):
    torch.manual_seed(M + N + K)
    A, W, b = torch.randn(M, K), torch.randn(N, K) / math.sqrt(K), torch.randn(N)
    ref = A.double() @ W.double().T + b.double()
    out = torch.empty(M, N, device=DEV)
    ops.gemm(g(A), ops.SplitWeight(g(W)), g(b), out, M, N, K)
    torch.cuda.synchronize()
    e_tc = err(out.cpu(), ref)
    ops.set_gemm_engine(0)
    out2 = torch.empty(M, N, device=DEV)
    ops.gemm(g(A), ops.SplitWeight(g(W)), g(b), out2, M, N, K)
    e_simt = err(out2.cpu(), ref)
    ops.set_gemm_engine(1)
    print(f"\n[tc-accuracy] bk={tc} M={M} N={N} K={K}: 3xTF32 err {e_tc:.3e}   SIMT fp32 err {e_simt:.3e}")
    assert e_tc < 2e-5


def test_tc_gemm_epilogues(tc):
    torch.manual_seed(2)
    R, l, C, K = 4, 50, 512, 1024
    M = R * l
    A, Wt, b = torch.randn(M, K), torch.randn(C, K) / math.sqrt(K), torch.randn(C)
    x0, ada = torch.randn(M, C), torch.randn(R, 6 * C)
    ref = x0.double() + (A.double() @ Wt.double().T + b.double()) * ada[:, C:2 * C].double().repeat_interleave(l, 0)
    x, ada_g = g(x0), g(ada)
    ops.gemm(g(A), ops.SplitWeight(g(Wt)), g(b), x, M, C, K, epilogue=ops.EPI_BIAS_GAMMA_RESID, gamma=ada_g[:, C:2 * C],
             gamma_row_stride=6 * C, rows_per_sample=l)
    assert err(x.cpu(), ref) < 1e-5
    out = torch.empty(M, C, device=DEV)
    ops.gemm(g(A), ops.SplitWeight(g(Wt)), g(b), out, M, C, K, epilogue=ops.EPI_BIAS_GELU)
    assert err(out.cpu(), F.gelu(A.double() @ Wt.double().T + b.double(), approximate="tanh")) < 1e-5


@pytest.mark.parametrize("cin,cout,ks,up,H", [(32, 32, 3, False, 16), (160, 160, 3, False, 24), (320, 160, 1, False, 16),
                                              (160, 160, 3, True, 12), (640, 320, 3, False, 16), (32, 640, 3, False, 16)])
def test_tc_conv(tc, cin, cout, ks, up, H):
    torch.manual_seed(8)
    B, Wd = 2, H + 3
    x = torch.randn(B, cin, H, Wd) * 2 + 0.3
    w = torch.randn(cout, cin, ks, ks) / math.sqrt(cin * ks * ks)
    b = torch.randn(cout)
    gam, bet = torch.rand(cin) + 0.5, torch.randn(cin) * 0.1
    xin = F.silu(F.group_norm(x, 32, gam, bet, 1e-6))
    if up:
        xin = F.interpolate(xin, scale_factor=2, mode="nearest")
    ref = F.conv2d(xin.double(), w.double(), b.double(), padding=ks // 2)
    Ho, Wo = ref.shape[2], ref.shape[3]
    resid = torch.randn(B, cout, Ho, Wo)
    ref = ref + resid.double()
    x_nhwc = g(x.permute(0, 2, 3, 1))
    wp = ops.SplitWeight(ops.repack_conv_weight(g(w), torch.empty(cout, ks * ks * cin, device=DEV)))
    a, bb = torch.empty(B, cin, device=DEV), torch.empty(B, cin, device=DEV)
    scratch = torch.empty(2 * B * 32 * ops.gn_chunks(H * Wd), dtype=torch.float64, device=DEV)
    ops.gn_stats(x_nhwc, g(gam), g(bet), a, bb, scratch, B, H * Wd, cin)
    out = torch.empty(B, Ho, Wo, cout, device=DEV)
    ops.conv2d(x_nhwc, wp, g(b), out, B, H, Wd, cin, cout, ks, in_a=a, in_b=bb, in_silu=True,
               resid=g(resid.permute(0, 2, 3, 1)), upsample2x=up)
    assert err(out.cpu().permute(0, 3, 1, 2), ref) < 2e-5


def test_tc_gemm_throughput(tc):
    """Not a benchmark (bench.py is): a sanity check that the tensor-core engine is far above the SIMT engine."""
    M, N, K = 16384, 1536, 1536
    A, W, b = torch.randn(M, K, device=DEV), ops.SplitWeight(torch.randn(N, K, device=DEV) / 40), torch.randn(N, device=DEV)
    out = torch.empty(M, N, device=DEV)
    res = {}
    for eng in (1, 0):
        ops.set_gemm_engine(eng)
        for _ in range(3):
            ops.gemm(A, W, b, out, M, N, K)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(A, W, b, out, M, N, K)
        e1.record()
        torch.cuda.synchronize()
        res[eng] = 10 * 2.0 * M * N * K / (e0.elapsed_time(e1) * 1e-3) / 1e12
    ops.set_gemm_engine(1)
    print(f"\n[tc-throughput] bk={tc} M={M} N={N} K={K}: 3xTF32 {res[1]:.1f} TFLOP/s   SIMT {res[0]:.1f} TFLOP/s")
    assert res[1] > res[0]
