"""A/B of the 2-CTA f16x3 GEMM epilogue (cvar_set_epilogue_overlap 0 / 1) on the d24 dense-layer shapes: time, TFLOP/s,
bit-identity of the two outputs, and the per-tile clock64 trace of CTA 0 (cvar_debug_set_trace) - where a tile's cycles go.
Diagnostic; bench.py is the benchmark."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import ops, _lib  # noqa: E402

dev = "cuda"


def timed(fn, n=10):
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


def trace(fn, nkb):
    tr = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
    _lib.load().cvar_debug_set_trace(tr.data_ptr())
    fn()
    torch.cuda.synchronize()
    _lib.load().cvar_debug_set_trace(None)
    t = tr.cpu().view(64, 8)
    rows = [i for i in range(2, 40) if t[i, 0] > 0 and t[i + 1, 0] > 0]
    if not rows:
        return "no trace"
    tile = sum((t[i + 1, 0] - t[i, 0]).item() for i in rows) / len(rows)
    mma = sum((t[i, 1] - t[i, 0]).item() for i in rows) / len(rows)
    drain = sum((t[i, 3] - t[i, 2]).item() for i in rows) / len(rows)
    store = sum((t[i, 4] - t[i, 3]).item() for i in rows) / len(rows)
    wait_tm = sum((t[i + 1, 0] - t[i, 1]).item() for i in rows) / len(rows)
    tma = sum((t[i, 6] - t[i, 5]).item() for i in rows) / len(rows)
    return (f"tile period {tile:8.0f} cyc | MMA issue span {mma:8.0f} | last commit -> next tile start {wait_tm:7.0f} | "
            f"epilogue: tmem held {drain:7.0f}, then stores {store:7.0f} | TMA first->last stage issue {tma:8.0f} "
            f"(ideal MMA {nkb * 1536} cyc)")


def run(M, N, K, epi, tag):
    torch.manual_seed(1)
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / math.sqrt(K)
    b = torch.randn(N, device=dev)
    gamma = torch.randn(M // 512 + 1, N, device=dev)
    kw = dict(epilogue=epi)
    if epi == ops.EPI_BIAS_GAMMA_RESID:
        kw.update(gamma=gamma, gamma_row_stride=N, rows_per_sample=512)
    ops.set_gemm_engine(4)
    W16 = ops.SplitWeight(W, f16=True)
    A16 = ops.F16Pair.from_tensor(A)
    out16 = ops.F16Pair.empty((M, N), dev) if epi in (ops.EPI_BIAS_GELU,) else None
    base = torch.randn(M, N, device=dev)
    res = {}
    for mode in (0, 1):
        ops.set_epilogue_overlap(mode)
        out = base.clone()
        call = lambda o=out: ops.gemm(None, W16, b, None if out16 is not None else o, M, N, K, A16=A16, out16=out16, **kw)
        call()
        torch.cuda.synchronize()
        res[mode] = (out16.hi.clone(), out16.lo.clone()) if out16 is not None else (out.clone(),)
        ms = timed(call)
        tr = trace(call, K // 64)
        print(f"{tag:18s} M={M:6d} N={N:5d} K={K:5d} overlap={mode}: {ms:7.3f} ms {2.0 * M * N * K / 1e9 / ms:6.1f} TF/s\n"
              f"    {tr}", flush=True)
    same = all(torch.equal(a, c) for a, c in zip(res[0], res[1]))
    print(f"    outputs bit-identical: {same}", flush=True)
    ops.set_epilogue_overlap(1)


if __name__ == "__main__":
    M = 65536
    run(M, 6144, 1536, ops.EPI_BIAS_GELU, "fc1 gelu")
    run(M, 1536, 6144, ops.EPI_BIAS_GAMMA_RESID, "fc2 gamma-resid")
    run(M, 1536, 1536, ops.EPI_BIAS_GAMMA_RESID, "proj gamma-resid")
    run(M, 4096, 1536, ops.EPI_BIAS, "head")
    run(16384, 6144, 1536, ops.EPI_BIAS_GELU, "fc1 scale6")
    run(4096, 6144, 1536, ops.EPI_BIAS_GELU, "fc1 scale3")
