"""Throughput sweep of the tcgen05 GEMM engine over the shapes of the d24 workload (diagnostic, not the benchmark)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import ops  # noqa: E402

ops.set_gemm_engine(1)
ops.set_tc_kblock(int(os.environ.get("CVAR_TC_BK", "32")))
dev = "cuda"


def run(M, N, K, epi, tag):
    A = torch.randn(M, K, device=dev)
    W = ops.SplitWeight(torch.randn(N, K, device=dev) / 40)
    b = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev)
    gamma = torch.randn(M // 512 + 1, N, device=dev)
    kw = dict(epilogue=epi)
    if epi == ops.EPI_BIAS_GAMMA_RESID:
        kw.update(gamma=gamma, gamma_row_stride=N, rows_per_sample=512)
    fn = lambda: ops.gemm(A, W, b, out, M, N, K, **kw)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{tag:28s} M={M:6d} N={N:5d} K={K:5d}: {ms:8.3f} ms {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)


M = 65536
run(M, 6144, 1536, ops.EPI_BIAS_GELU, "fc1 gelu")
run(M, 6144, 1536, ops.EPI_BIAS, "fc1-shape bias only")
run(M, 1536, 6144, ops.EPI_BIAS_GAMMA_RESID, "fc2 gamma-resid")
run(M, 1536, 1536, ops.EPI_BIAS_GAMMA_RESID, "proj gamma-resid")
run(M, 4608, 1536, ops.EPI_BIAS, "qkv-shape bias")
run(M, 4096, 1536, ops.EPI_BIAS, "head")
run(16384, 6144, 1536, ops.EPI_BIAS_GELU, "fc1 scale6")
run(4096, 6144, 1536, ops.EPI_BIAS_GELU, "fc1 scale3")
run(1024, 6144, 1536, ops.EPI_BIAS_GELU, "fc1 scale1")
run(256, 6144, 1536, ops.EPI_BIAS_GELU, "fc1 scale0")
