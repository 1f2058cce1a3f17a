"""Pipeline trace of the tcgen05 GEMM engine (CTA 0, first K-blocks): who waits for whom, in SM cycles."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import ops, _lib  # noqa: E402

ops.set_gemm_engine(1)
ops.set_tc_kblock(int(os.environ.get("CVAR_TC_BK", "32")))
M, N, K = 65536, 6144, 1536
A = torch.randn(M, K, device="cuda")
W = ops.SplitWeight(torch.randn(N, K, device="cuda") / 40)
b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda")
ops.gemm(A, W, b, out, M, N, K)
torch.cuda.synchronize()
tr = torch.zeros(4 * 64 * 2, dtype=torch.int64, device="cuda")
_lib.load().cvar_debug_set_trace(tr.data_ptr())
ops.gemm(A, W, b, out, M, N, K)
torch.cuda.synchronize()
_lib.load().cvar_debug_set_trace(None)
t = tr.cpu().view(4, 64, 2)
t0 = t[t > 0].min().item()
t = t - t0
print("kb |  TMA issue | A empty seen  A done | MMA: B full seen  A ready seen  commit issued")
for kb in range(24, 40):
    print(f"{kb:2d} | {t[0, kb, 0]:10d} | {t[2, kb, 0]:12d} {t[2, kb, 1]:7d} | {t[3, kb, 0]:16d} {t[1, kb, 0]:13d} {t[3, kb, 1]:14d}")
d = t[1, 20:46, 0][1:] - t[1, 20:46, 0][:-1]
print("MMA operands-ready period per k-block (cycles): mean %.0f min %d max %d" % (d.float().mean(), d.min(), d.max()))
print("TMA issue -> B full seen by MMA: mean %.0f" % (t[3, 20:46, 0] - t[0, 20:46, 0]).float().mean())
print("A empty -> A done:     mean %.0f" % (t[2, 20:46, 1] - t[2, 20:46, 0]).float().mean())
print("MMA ready -> commit issued: mean %.0f" % (t[3, 20:46, 1] - t[1, 20:46, 0]).float().mean())
