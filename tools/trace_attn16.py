"""Phase trace + timing of the FP16-pair tensor-core attention kernel (attn16_tc_kernel) at the d24 / batch-64 shapes.
CTA (0,0,0), KV tiles j < 22, SM cycles (cvar_debug_set_attn_trace).  CVAR_ATTN_ONE_PASS=0|1 selects the softmax variant.
Diagnostic; bench.py is the benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import ops, _lib  # noqa: E402

dev = "cuda"
R, H, T = 128, 24, 1360
kv = ops.KVCache16(R, H, T, dev)
kv.k_hi.normal_().mul_(16), kv.k_lo.normal_().mul_(2.0 ** -8)          # qk pairs: 16 x = hi + lo
kv.vt_hi.normal_(), kv.vt_lo.normal_()


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"CVAR_ATTN_ONE_PASS={os.environ.get('CVAR_ATTN_ONE_PASS', '1')} fast_mode={ops.get_fast_mode()}")
tot = 0.0
for (l, L) in ((512, 1360), (338, 848), (200, 510), (128, 310), (72, 182), (50, 110), (32, 60)):
    q16 = ops.F16Pair.empty((R, H, l, 64), dev)
    q16.hi.normal_().mul_(16), q16.lo.normal_().mul_(2.0 ** -8)
    o16 = ops.F16Pair.empty((R, l, H * 64), dev)
    ms = timed(lambda: ops.attn_kvcache16(q16, kv, None, R, H, l, L, 1 / 32, engine=1, out16=o16))
    fl = 4.0 * l * L * 64 * R * H
    by = (2.0 * l + 2.0 * L) * 64 * 4 * R * H
    tot += ms
    print(f"attn16 l={l:3d} L={L:4d}: {ms:.3f} ms  {fl / ms / 1e9:6.1f} TFLOP/s  {by / ms / 1e6:5.0f} GB/s algorithmic")
    if l == 512:
        tr = torch.zeros(3 * 32 * 8, dtype=torch.int64, device=dev)
        _lib.load().cvar_debug_set_attn_trace(tr.data_ptr())
        ops.attn_kvcache16(q16, kv, None, R, H, l, L, 1 / 32, engine=1, out16=o16)
        torch.cuda.synchronize()
        _lib.load().cvar_debug_set_attn_trace(None)
        t = tr.cpu().view(3, 32, 8)
        t0 = t[t > 0].min().item()
        t = torch.where(t > 0, t - t0, t)
        n = (L + 63) // 64
        print(" j | softmax: s_full seen  max known  P stored  p_ready | MMA: p_ready seen  PV issued  k_full(j+2)  S(j+2) issued | TMA: K(j) issue  V(j) issue")
        for j in range(6, 14):
            s, m, a = t[0, j], t[1, j], t[2, j]
            print(f"{j:2d} | {s[0]:9d} {s[1]:10d} {s[2]:9d} {s[3]:8d} | {m[0]:9d} {m[1]:10d} {m[2]:11d} {m[3]:13d} | {a[0]:9d} {a[1]:10d}")
        S, M = t[0, 4:n - 2].float(), t[1, 4:n - 2].float()
        per = (t[0, 5:n - 2, 0] - t[0, 4:n - 3, 0]).float().mean()
        print(f"period per 64-key tile {per:.0f} cyc (MMA work per CTA and tile: 768; two CTAs share the SM)")
        print(f"softmax: wait s_full {(t[0, 5:n - 2, 0] - t[0, 4:n - 3, 3]).float().mean():.0f} | S load + max {(S[:, 1] - S[:, 0]).mean():.0f} | "
              f"exp + P store {(S[:, 2] - S[:, 1]).mean():.0f} | drain / wait st / signal {(S[:, 3] - S[:, 2]).mean():.0f}")
        print(f"MMA: p_ready(j) signalled -> seen {(M[:, 0] - S[:, 3]).mean():.0f} | PV issue {(M[:, 1] - M[:, 0]).mean():.0f} | "
              f"wait k_full(j+2) {(M[:, 2] - M[:, 1]).mean():.0f} | S issue {(M[:, 3] - M[:, 2]).mean():.0f} | "
              f"S(j+2) issued -> s_full(j+2) seen by softmax {(t[0, 6:n, 0] - t[1, 4:n - 2, 3]).float().mean():.0f}")
print(f"sum over the seven tensor-core scales: {tot:.3f} ms per layer")
