"""Runs the three hot kernels once each at the sizes of the d24 / batch-64 workload (last scale), for ncu captures:
   ncu --set full --clock-control none --import-source on -k regex:<name> -o gpurun_out/<name> python tools/prof_kernels.py <which>
which in {gemm, conv, attn, all}.  Never a benchmark: numbers printed under a profiler are not bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
engine = int(os.environ.get("CVAR_GEMM_ENGINE", "4"))
ops.set_gemm_engine(engine)
ops.set_tc_kblock(int(os.environ.get("CVAR_TC_BK", "32")))
dev = "cuda"
torch.manual_seed(0)


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if which in ("gemm", "all"):
    M, N, K = 65536, 6144, 1536            # fc1 of d24 at the last scale (R*l = 128*512 rows)
    A = torch.randn(M, K, device=dev)
    b = torch.randn(N, device=dev)
    if engine == 4:      # FP16 pairs in, FP16 pair out: exactly what the sampler's fc1 call does
        W = ops.SplitWeight(torch.randn(N, K, device=dev) / 40, f16=True)
        A16, out = ops.F16Pair.from_tensor(A), ops.F16Pair.empty((M, N), dev)
        ms = timed(lambda: ops.gemm(None, W, b, None, M, N, K, epilogue=ops.EPI_BIAS_GELU, A16=A16, out16=out))
    else:
        W = ops.SplitWeight(torch.randn(N, K, device=dev) / 40)
        out = torch.empty(M, N, device=dev)
        ms = timed(lambda: ops.gemm(A, W, b, out, M, N, K, epilogue=ops.EPI_BIAS_GELU))
    print(f"gemm fc1 M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
    del A, out

if which in ("conv", "all"):
    B, H, C = 8, 256, 160                  # decoder up.0 ResnetBlock conv at 256x256
    x = torch.randn(B, H, H, C, device=dev)
    w = ops.SplitWeight(torch.randn(C, 9 * C, device=dev) / 38)
    b = torch.randn(C, device=dev)
    out = torch.empty(B, H, H, C, device=dev)
    if engine == 4:
        x16, w16 = ops.F16Pair.from_tensor(x), ops.F16Pair.from_tensor(w.w)
        ms = timed(lambda: ops.conv2d(None, w, b, out, B, H, H, C, C, 3, x16=x16, w16=w16))
    else:
        ms = timed(lambda: ops.conv2d(x, w, b, out, B, H, H, C, C, 3))
    print(f"conv3x3 B={B} {H}x{H} {C}->{C}: {ms:.3f} ms  {2.0 * B * H * H * C * 9 * C / ms / 1e9:.1f} TFLOP/s")
    del x, out

if which in ("attn", "all"):
    R, Hh, l, L, T = 128, 24, 512, 1360, 1360
    q = torch.randn(R, Hh, l, 64, device=dev)
    kv = ops.KVCache(R, Hh, T, dev)
    for t in (kv.k_hi, kv.vt_hi):
        t.normal_()
        t.copy_(t.view(torch.int32).bitwise_and_(-8192).view(torch.float32))
    for t in (kv.k_lo, kv.vt_lo):
        t.normal_().mul_(2.0 ** -12)
    out = torch.empty(R, l, Hh * 64, device=dev)
    fl = 4.0 * l * L * 64 * R * Hh
    by = (2.0 * l + 2.0 * L) * 64 * 4 * R * Hh
    for eng in (1, 0):
        ms = timed(lambda: ops.attn_kvcache(q, kv, out, R, Hh, l, L, 1 / 32, engine=eng))
        print(f"attn engine={eng} R={R} H={Hh} l={l} L={L}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s algorithmic")

if which in ("attn16", "all"):
    R, Hh, l, L, T = 128, 24, 512, 1360, 1360
    kv = ops.KVCache16(R, Hh, T, dev)
    kv.k_hi.normal_().mul_(16), kv.k_lo.normal_().mul_(2.0 ** -8)          # qk pairs: 16 x = hi + lo
    kv.vt_hi.normal_(), kv.vt_lo.normal_()
    q16 = ops.F16Pair.empty((R, Hh, l, 64), dev)
    q16.hi.normal_().mul_(16), q16.lo.normal_().mul_(2.0 ** -8)
    o16 = ops.F16Pair.empty((R, l, Hh * 64), dev)
    fl = 4.0 * l * L * 64 * R * Hh
    by = (2.0 * l + 2.0 * L) * 64 * 4 * R * Hh
    for eng in (1, 0):
        ms = timed(lambda: ops.attn_kvcache16(q16, kv, None, R, Hh, l, L, 1 / 32, engine=eng, out16=o16))
        print(f"attn16 engine={eng} R={R} H={Hh} l={l} L={L}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s algorithmic")
    for (ls, Ls) in ((338, 848), (200, 510), (128, 310), (72, 182), (50, 110), (32, 60)):
        q16s = ops.F16Pair.empty((R, Hh, ls, 64), dev)
        q16s.hi.normal_().mul_(16), q16s.lo.normal_().mul_(2.0 ** -8)
        o16s = ops.F16Pair.empty((R, ls, Hh * 64), dev)
        ms = timed(lambda: ops.attn_kvcache16(q16s, kv, None, R, Hh, ls, Ls, 1 / 32, engine=1, out16=o16s))
        ms0 = timed(lambda: ops.attn_kvcache16(q16s, kv, None, R, Hh, ls, Ls, 1 / 32, engine=0, out16=o16s)) if ls <= 72 else float("nan")
        print(f"attn16 engine=1 l={ls} L={Ls}: {ms:.3f} ms  {4.0 * ls * Ls * 64 * R * Hh / ms / 1e9:.1f} TFLOP/s   (SIMT: {ms0:.3f} ms)")
