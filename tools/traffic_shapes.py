"""One launch of each dense-layer shape of d24 at the last scale (fc1, fc2, proj, head, QKV), for an ncu DRAM-traffic capture:
   CVAR_GROUP_M=<g> ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
       -k regex:tc_gemm2_kernel --csv --log-file out.csv python tools/traffic_shapes.py
Prints the algorithmic bytes of every launch (operands read once + result written once) in launch order.  Diagnostic."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import ops  # noqa: E402

dev = "cuda"
ops.set_gemm_engine(4)
torch.manual_seed(0)
R, l, H = 128, 512, 24
M, C = R * l, H * 64


def dense(tag, N, K, epi):
    A16 = ops.F16Pair.from_tensor(torch.randn(M, K, device=dev))
    W = ops.SplitWeight(torch.randn(N, K, device=dev) / math.sqrt(K), f16=True)
    b = torch.randn(N, device=dev)
    kw, out, out16 = dict(epilogue=epi), None, None
    if epi == ops.EPI_BIAS_GELU:
        out16 = ops.F16Pair.empty((M, N), dev)
        alg = 4 * (M * K + N * K + M * N)
    elif epi == ops.EPI_BIAS_GAMMA_RESID:
        out = torch.randn(M, N, device=dev)
        kw.update(gamma=torch.randn(R, N, device=dev), gamma_row_stride=N, rows_per_sample=l)
        alg = 4 * (M * K + N * K + 2 * M * N)
    else:
        out = torch.empty(M, N, device=dev)
        alg = 4 * (M * K + N * K + M * N)
    torch.cuda.synchronize()
    ops.gemm(None, W, b, out, M, N, K, A16=A16, out16=out16, **kw)
    torch.cuda.synchronize()
    print(f"{tag:5s} M={M} N={N} K={K}: algorithmic {alg} B", flush=True)


dense("fc1", 4 * C, C, ops.EPI_BIAS_GELU)
dense("fc2", C, 4 * C, ops.EPI_BIAS_GAMMA_RESID)
dense("proj", C, C, ops.EPI_BIAS_GAMMA_RESID)
dense("head", 4096, C, ops.EPI_BIAS)

T = 1360
A16 = ops.F16Pair.from_tensor(torch.randn(M, C, device=dev))
Wq = ops.SplitWeight(torch.randn(3 * C, C, device=dev) / math.sqrt(C), f16=True)
qb, kb, vb = (torch.randn(C, device=dev) for _ in range(3))
q16 = ops.F16Pair.empty((R, H, l, 64), dev)
kv = ops.KVCache16(R, H, T, dev)
torch.cuda.synchronize()
ops.qkv_project16(A16, Wq, qb, kb, vb, q16, kv, R, l, T - l, H, False, None)
torch.cuda.synchronize()
print(f"qkv   M={M} N={3 * C} K={C}: algorithmic {4 * (M * C + 3 * C * C + M * 3 * C)} B", flush=True)
