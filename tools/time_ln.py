"""Times cvar_ln_modulate at the shapes of the d24 / batch-64 step (FP16-pair output) and reports GB/s (x read + pair write)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import ops  # noqa: E402

dev = "cuda"
R, C = 128, 1536
ada = torch.randn(R, 6 * C, device=dev)
tot_ms, tot_b = 0.0, 0.0
for l in (2, 8, 18, 32, 50, 72, 128, 200, 338, 512):
    M = R * l
    x = torch.randn(M, C, device=dev)
    out = ops.F16Pair.empty((M, C), dev)
    f = lambda: ops.ln_modulate(x, ada[:, 2 * C:3 * C], ada[:, 4 * C:5 * C], 6 * C, None, M, C, l, 1e-6, out16=out)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    by = M * C * 8.0
    tot_ms += ms * 49
    tot_b += by * 49
    print(f"l={l:4d} M={M:6d}: {ms * 1e3:8.1f} us  {by / ms / 1e6:7.0f} GB/s")
print(f"one sampling call (49 LN per scale): {tot_ms:.1f} ms, {tot_b / tot_ms / 1e6:.0f} GB/s average")
