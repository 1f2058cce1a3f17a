"""Decoder alone at the bench shape (2 x 64 maps of 16x16x32 -> 64 images of 512x256): whole-pass time with CUDA events,
with the GroupNorm statistics taken from the conv epilogues (default) and by the separate read pass.  Diagnostic."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import VQVAE, ops, weights as W  # noqa: E402
from controlvar_b200.config import PathConfig  # noqa: E402

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = PathConfig(depth=2)
vae = VQVAE(ch=160).to(dev)
vae.load_state_dict(W.synthetic_vae_state_dict(cfg, 0, device=dev))
f_hat = torch.randn(B, 32, 32, 16, device=dev) * 0.5


def timed(n=3):
    for _ in range(2):
        vae._fhat_halves_to_img(f_hat, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        img = vae._fhat_halves_to_img(f_hat, B)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, img


res = {}
for fuse in (True, False, True):
    vae.fuse_gn_stats = fuse
    ms, img = timed()
    res[fuse] = img
    print(f"decoder, {2 * B} maps, GroupNorm statistics from the conv epilogue = {fuse}: {ms:.2f} ms per pass "
          f"({2 * B * 393.04 / ms:.1f} TFLOP/s of convolution)")
print(f"max |pixel| difference fused vs read pass: {(res[True] - res[False]).abs().max().item():.3e}")
