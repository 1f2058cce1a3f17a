"""Decoder pixel error vs the CPU oracle, per GEMM engine: (a) random f_hat like the unit test, (b) a realistic f_hat
produced by the sampler itself (d4 golden config)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlvar_b200 import VQVAE, build_control_var, ops, weights as W  # noqa: E402
from controlvar_b200.config import PathConfig  # noqa: E402
from oracle import controlvar_oracle as O  # noqa: E402

DEV = "cuda"
cfg = PathConfig(depth=4)
vsd = W.synthetic_vae_state_dict(cfg, 0)
vae = VQVAE(ch=160).to(DEV)
vae.load_state_dict(vsd)
torch.set_num_threads(os.cpu_count())

cases = {}
torch.manual_seed(10)
cases["random N(0,1.5^2) (unit test input)"] = torch.randn(1, 32, 16, 16) * 1.5
torch.manual_seed(11)
cases["random N(0,1)"] = torch.randn(1, 32, 16, 16)
# realistic: f_hat from the oracle sampler (d4, full pyramid)
sd = W.synthetic_var_state_dict(cfg, 0)
o = O.autoregressive_infer_cfg(sd, vsd, cfg.patch_nums, 4, 1, torch.tensor([7]), torch.tensor([1]), 1.5, 900, 0.96,
                               O.cpu_generator_noise(1), decode=False)
cases["sampler f_hat, control half"] = o["f_hat"][:, :, :16].contiguous()
cases["sampler f_hat, image half"] = o["f_hat"][:, :, 16:].contiguous()
for name, f in cases.items():
    ref = O.fhat_to_img(f.clone(), vsd)
    line = f"{name:38s} |f_hat| max {f.abs().max():5.2f}  frac of pixels clamped {((ref.abs() >= 1).float().mean()):.3f} :"
    for eng in (0, 1):
        ops.set_gemm_engine(eng)
        got = vae.fhat_to_img(f.to(DEV)).cpu()
        d = (got - ref).abs()
        line += f"  engine {eng}: max {d.max():.3e} mean {d.mean():.2e} p99.9 {d.flatten().kthvalue(int(0.999 * d.numel()))[0]:.2e}"
    print(line)
