"""Decoder: pixel error vs the CPU oracle and time, as a function of which layers run on tensor cores
(VQVAE.tc_min_hw: convolutions with an output side below it use the SIMT fp32 engine)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import VQVAE, ops, weights as W  # noqa: E402
from controlvar_b200.config import PathConfig  # noqa: E402
from oracle import controlvar_oracle as O  # noqa: E402

DEV = "cuda"
cfg = PathConfig(depth=4)
vsd = W.synthetic_vae_state_dict(cfg, 0)
vae = VQVAE(ch=160).to(DEV)
vae.load_state_dict(vsd)
torch.set_num_threads(os.cpu_count())
sd = W.synthetic_var_state_dict(cfg, 0)
fs = []
for seed in (1, 2):
    o = O.autoregressive_infer_cfg(sd, vsd, cfg.patch_nums, 4, 1, torch.tensor([7 * seed]), torch.tensor([seed]), 1.5, 900,
                                   0.96, O.cpu_generator_noise(seed), decode=False)
    fs += [o["f_hat"][:, :, :16].contiguous(), o["f_hat"][:, :, 16:].contiguous()]
refs = [O.fhat_to_img(f.clone(), vsd) for f in fs]
big = torch.cat(fs * 16, 0).to(DEV)          # 64 images for timing
ops.set_gemm_engine(3)
print("tc_min_hw | layers on tensor cores            | max pixel err (4 realistic f_hat) | mean err | decode 64 imgs")
KS = {}   # row label -> ksplit_min_k
ROWS = ((0, "all", 3), (32, "output side >= 32", 3), (64, ">= 64", 3), (128, ">= 128", 3), (256, ">= 256 only", 3),
        (9999, "no conv (attn-block GEMMs still TC)", 3), (9999, "nothing: global SIMT engine", 0),
        (0, "f16x3: all", 4), (32, "f16x3: output side >= 32", 4), (64, "f16x3: >= 64", 4), (128, "f16x3: >= 128", 4),
        (0, "f16x3: all, K-split K>=5760", 5), (0, "f16x3: all, K-split K>=2880", 6),
        (32, "f16x3: >= 32, K-split K>=5760", 5), (32, "f16x3: >= 32, K-split K>=2880", 6))
if len(sys.argv) > 1:
    ROWS = tuple(r for r in ROWS if str(r[2]) in sys.argv[1:])
for hw, what, eng in ROWS:
    vae.tc_min_hw = hw
    vae.ksplit_min_k = {5: 5760, 6: 2880}.get(eng, 0)
    ops.set_gemm_engine(4 if eng in (5, 6) else eng)
    tag = {3: "", 4: "'", 0: "*", 5: "k", 6: "K"}[eng]
    mx, mean = 0.0, 0.0
    for f, r in zip(fs, refs):
        d = (vae.fhat_to_img(f.to(DEV)).cpu() - r).abs()
        mx, mean = max(mx, d.max().item()), mean + d.mean().item() / len(fs)
    for _ in range(2):
        vae.fhat_to_img(big)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vae.fhat_to_img(big)
    e1.record()
    torch.cuda.synchronize()
    print(f"{str(hw) + tag:>9s} | {what:34s} | {mx:.3e}                         | {mean:.2e} | {e0.elapsed_time(e1):8.1f} ms")
