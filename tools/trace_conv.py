"""Tile trace of the 2-CTA f16x3 convolution kernel (CTA 0, cvar_debug_set_trace) on the decoder's two dominant layers."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import ops, _lib  # noqa: E402

dev = "cuda"
ops.set_gemm_engine(4)


def run(B, H, C, Cout, with_resid=False):
    torch.manual_seed(0)
    x = torch.randn(B, H, H, C, device=dev)
    wp = torch.randn(Cout, 9 * C, device=dev) / 38
    b = torch.randn(Cout, device=dev)
    out = torch.empty(B, H, H, Cout, device=dev)
    x16, w16 = ops.F16Pair.from_tensor(x), ops.F16Pair.from_tensor(wp)
    resid = torch.randn(B, H, H, Cout, device=dev) if with_resid else None
    call = lambda: ops.conv2d(None, wp, b, out, B, H, H, C, Cout, 3, x16=x16, w16=w16, resid=resid)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    tr = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
    _lib.load().cvar_debug_set_trace(tr.data_ptr())
    call()
    torch.cuda.synchronize()
    _lib.load().cvar_debug_set_trace(None)
    t = tr.cpu().view(64, 8)
    rows = [i for i in range(2, 30) if t[i, 0] > 0 and t[i + 1, 0] > 0]
    n = len(rows)
    tile = sum((t[i + 1, 0] - t[i, 0]).item() for i in rows) / n
    mma = sum((t[i, 1] - t[i, 0]).item() for i in rows) / n
    held = sum((t[i, 3] - t[i, 2]).item() for i in rows) / n
    store = sum((t[i, 4] - t[i, 3]).item() for i in rows) / n
    bn = 160 if Cout % 160 == 0 else Cout
    ideal = 9 * (C // 32) * 2 * 3 * (bn // 2)
    print(f"conv3x3 B={B} {H}x{H} {C}->{Cout}{' + residual' if with_resid else ''}: {ms:.3f} ms {2.0 * B * H * H * Cout * 9 * C / ms / 1e9:.1f} TFLOP/s | tile period {tile:.0f} "
          f"cyc, MMA issue span {mma:.0f} (ideal MMA {ideal}), hand-over {tile - mma:.0f}, tmem held {held:.0f}, stores after "
          f"release {store:.0f}", flush=True)


run(8, 256, 160, 160)
run(8, 256, 160, 160, with_resid=True)
run(32, 256, 160, 160)
run(32, 256, 160, 160, with_resid=True)
run(8, 128, 320, 320)
run(16, 64, 320, 320)
