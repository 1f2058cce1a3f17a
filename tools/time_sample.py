"""Times cvar_cfg_sample at the last scale of the d24 / batch-64 step (32 768 rows of 4096 logits, two guidance groups) with the
phases switched off from the outside (top_k = 0: no radix select; top_p = 0: no compaction / sort / running sum).  Diagnostic."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
B, l, V = 64, 512, 4096
logits = torch.randn(2 * B * l, V, device=dev) * 2.5
noise = torch.empty(B * l, V, device=dev).exponential_()
idx = torch.empty(B, l, dtype=torch.int64, device=dev)
ref = None
for top_k, top_p in ((900, 0.96), (0, 0.96), (900, 0.0), (0, 0.0)):
    f = lambda: ops.cfg_sample(logits, noise, idx, B, l, V, 1.5 * 9 / 9, top_k, top_p)
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"top_k={top_k:4d} top_p={top_p:4.2f}: {ms:7.3f} ms  ({(3 * 4.0 * B * l * V) / ms / 1e6:6.0f} GB/s algorithmic)  "
          f"checksum {int(idx.sum().item())}")
