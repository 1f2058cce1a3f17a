"""Times Decoder.conv_out (160 -> 3 channels, 3x3, image output) at the bench shape: 128 maps of 256 x 256.
CVAR_CONV3_ROWS4=0 selects the 2 x 128 tile kernel (round 2, first step) instead of the 4-rows-per-thread one.  Diagnostic."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
B, H, cin = 128, 256, 160
x = torch.randn(B, H, H, cin, device=dev)
w = ops.repack_conv_weight(torch.randn(3, cin, 3, 3, device=dev) / math.sqrt(cin * 9), torch.empty(3, 9 * cin, device=dev))
b = torch.randn(3, device=dev) * 0.1
img = torch.zeros(B // 2, 3, 2 * H, H, device=dev)
f = lambda: ops.conv2d(x, w, b, img, B, H, H, cin, 3, 3, out_mode=1, out_rows_total=2 * H, row_offset=0, out_samples=B // 2)
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
print(f"conv_out {B} x {H} x {H} x {cin} -> 3 (CVAR_CONV3_ROWS4={os.environ.get('CVAR_CONV3_ROWS4', '1')}): {ms:.3f} ms  "
      f"{2.0 * B * H * H * cin * 27 / ms / 1e9:.1f} TFLOP/s fp32  {B * H * H * cin * 4 / ms / 1e6:.0f} GB/s of input  checksum {img.double().sum().item():.6f}")
