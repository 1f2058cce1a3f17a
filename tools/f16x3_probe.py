"""Accuracy and throughput of the FP16-pair engine (4) next to the 3xTF32 pair engine (3) on the d24 dense-layer shapes
(diagnostic; bench.py is the benchmark).  Errors are max |out - fp64| / max |fp64| on a 2048-row slice."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import ops  # noqa: E402

dev = "cuda"


def timed(fn, n=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(M, N, K, epi, tag):
    torch.manual_seed(1)
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / math.sqrt(K)
    b = torch.randn(N, device=dev)
    gamma = torch.randn(M // 512 + 1, N, device=dev)
    kw = dict(epilogue=epi)
    if epi == ops.EPI_BIAS_GAMMA_RESID:
        kw.update(gamma=gamma, gamma_row_stride=N, rows_per_sample=512)
    rows = min(M, 2048)
    ref = A[:rows].double() @ W.double().T + b.double()
    res = {}
    # engine 3
    ops.set_gemm_engine(3)
    Wt = ops.SplitWeight(W)
    At = ops.SplitWeight(A)
    out = torch.zeros(M, N, device=dev)
    ms3 = timed(lambda: ops.gemm(At.hi, Wt, b, out, M, N, K, A_lo=At.lo, **kw))
    o = torch.zeros(M, N, device=dev)
    ops.gemm(At.hi, Wt, b, o, M, N, K, A_lo=At.lo)
    e3 = ((o[:rows].double() - ref).abs().max() / ref.abs().max()).item()
    del Wt, At
    # engine 4
    ops.set_gemm_engine(4)
    W16 = ops.SplitWeight(W, f16=True)
    A16 = ops.F16Pair.from_tensor(A)
    ms4 = timed(lambda: ops.gemm(None, W16, b, out, M, N, K, A16=A16, **kw))
    ops.gemm(None, W16, b, o, M, N, K, A16=A16)
    e4 = ((o[:rows].double() - ref).abs().max() / ref.abs().max()).item()
    ops.set_gemm_engine(0)
    ops.gemm(A[:rows].contiguous(), W, b, o[:rows], rows, N, K)
    e0 = ((o[:rows].double() - ref).abs().max() / ref.abs().max()).item()
    fl = 2.0 * M * N * K / 1e9
    print(f"{tag:22s} M={M:6d} N={N:5d} K={K:5d}: 3xTF32 {ms3:7.3f} ms {fl / ms3:6.1f} TF/s err {e3:.2e} | "
          f"f16x3 {ms4:7.3f} ms {fl / ms4:6.1f} TF/s err {e4:.2e} | SIMT err {e0:.2e}", flush=True)


M = 65536
run(M, 6144, 1536, ops.EPI_BIAS_GELU, "fc1 gelu")
run(M, 1536, 6144, ops.EPI_BIAS_GAMMA_RESID, "fc2 gamma-resid")
run(M, 1536, 1536, ops.EPI_BIAS_GAMMA_RESID, "proj gamma-resid")
run(M, 4608, 1536, ops.EPI_BIAS, "qkv-shape bias")
run(M, 4096, 1536, ops.EPI_BIAS, "head")
run(16384, 6144, 1536, ops.EPI_BIAS_GELU, "fc1 scale6")
run(4096, 6144, 1536, ops.EPI_BIAS_GELU, "fc1 scale3")
run(256, 6144, 1536, ops.EPI_BIAS_GELU, "fc1 scale0")
