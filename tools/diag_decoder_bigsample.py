"""Worst decoder pixel error vs the CPU oracle over many realistic f_hat (default engine policy)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import VQVAE, ops, weights as W  # noqa: E402
from controlvar_b200.config import PathConfig  # noqa: E402
from oracle import controlvar_oracle as O  # noqa: E402

n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = PathConfig(depth=4)
vsd, sd = W.synthetic_vae_state_dict(cfg, 0), W.synthetic_var_state_dict(cfg, 0)
vae = VQVAE(ch=160).to("cuda")
vae.load_state_dict(vsd)
torch.set_num_threads(os.cpu_count())
if "CVAR_KSPLIT_MIN_K" in os.environ:
    vae.ksplit_min_k = int(os.environ["CVAR_KSPLIT_MIN_K"])
if "CVAR_TC_MIN_HW" in os.environ:
    vae.tc_min_hw = int(os.environ["CVAR_TC_MIN_HW"])
print(f"engine {ops.get_gemm_engine()}, tc_min_hw {vae._min_hw()}")
worst, means, n = 0.0, [], 0
for seed in range(n_calls):
    B = 2
    o = O.autoregressive_infer_cfg(sd, vsd, cfg.patch_nums, 4, B, torch.tensor([13 * seed % 1000, 7 * seed % 1000]),
                                   torch.tensor([seed % 4, (seed + 1) % 4]), 1.5, 900, 0.96, O.cpu_generator_noise(100 + seed),
                                   decode=False)
    for half in (o["f_hat"][:, :, :16].contiguous(), o["f_hat"][:, :, 16:].contiguous()):
        ref = O.fhat_to_img(half.clone(), vsd)
        d = (vae.fhat_to_img(half.to("cuda")).cpu() - ref).abs()
        per_img = d.flatten(1).max(1)[0]
        worst = max(worst, per_img.max().item())
        means.append(d.mean().item())
        n += B
    print(f"  after {n:3d} images: worst pixel {worst:.3e}  mean {sum(means) / len(means):.2e}", flush=True)
