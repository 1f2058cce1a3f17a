"""Phase trace of the tensor-core attention kernel (CTA 0, first 22 KV tiles, SM cycles)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import ops, _lib  # noqa: E402

R, H, l, L = 128, 24, 512, 1360
q = torch.randn(R, H, l, 64, device="cuda")
kv = ops.KVCache(R, H, L, "cuda")
for t in (kv.k_hi, kv.vt_hi):
    t.normal_()
    t.copy_(t.view(torch.int32).bitwise_and_(-8192).view(torch.float32))
for t in (kv.k_lo, kv.vt_lo):
    t.normal_().mul_(2.0 ** -12)
out = torch.empty(R, l, H * 64, device="cuda")
ops.attn_kvcache(q, kv, out, R, H, l, L, 1 / 32, engine=1)
torch.cuda.synchronize()
tr = torch.zeros(2 * 32 * 8, dtype=torch.int64, device="cuda")
_lib.load().cvar_debug_set_attn_trace(tr.data_ptr())
ops.attn_kvcache(q, kv, out, R, H, l, L, 1 / 32, engine=1)
torch.cuda.synchronize()
_lib.load().cvar_debug_set_attn_trace(None)
t = tr.cpu().view(2, 32, 8)
t0 = t[t > 0].min().item()
t = t - t0
n = (L + 63) // 64
print(" j | softmax: s_full  S-loaded  p-done  o_full(j-1)  O-upd  P-stored | MMA: S(j+1) issued  p_ready seen  PV issued")
for j in range(4, 14):
    s, m = t[0, j], t[1, j]
    print(f"{j:2d} | {s[0]:9d} {s[1]:9d} {s[2]:8d} {s[3]:10d} {s[4]:8d} {s[5]:9d} | {m[0]:12d} {m[1]:14d} {m[2]:11d}")
sl = slice(3, n - 1)
S = t[0, sl].float()
M = t[1, sl].float()
per = (t[0, 4:n - 1, 0] - t[0, 3:n - 2, 0]).float().mean()
print(f"\nperiod per 64-key tile: {per:.0f} cycles   (MMA work per tile at 100 %: 1536)")
print(f"softmax thread, mean cycles per tile:  wait s_full {(t[0, 4:n-1, 0] - t[0, 3:n-2, 5]).float().mean():.0f} | "
      f"TMEM ld S {(S[:, 1] - S[:, 0]).mean():.0f} | mask+max+exp+sum {(S[:, 2] - S[:, 1]).mean():.0f} | "
      f"wait o_full(j-1) {(S[:, 3] - S[:, 2]).mean():.0f} | TMEM ld O + update {(S[:, 4] - S[:, 3]).mean():.0f} | "
      f"split + TMEM st P + signal {(S[:, 5] - S[:, 4]).mean():.0f}")
print(f"MMA thread, mean cycles per tile:  wait p_ready after issuing S(j+1) {(M[:, 1] - M[:, 0]).mean():.0f} | "
      f"issue PV {(M[:, 2] - M[:, 1]).mean():.0f}")
