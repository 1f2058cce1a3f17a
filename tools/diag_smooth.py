"""more_smooth goldens: per-scale token mismatches, f_hat and pixel error of the GPU path per engine.  Diagnostic."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import golden_names, load_golden  # noqa: E402
from test_gpu_sampler import build  # noqa: E402
from controlvar_b200 import ops  # noqa: E402

for name in [n for n in golden_names() if n.startswith("smooth")]:
    gold = load_golden(name)
    m, cfg = gold["meta"], gold["cfg"]
    for eng in (0, 4):
        ops.set_gemm_engine(eng)
        vae, var, _, vsd = build(cfg, m["weight_seed"])
        img = var.autoregressive_infer_cfg(m["B"], torch.tensor(m["labels"]), g_seed=m["seed"], cfg=m["cfg"], top_k=m["top_k"],
                                           top_p=m["top_p"], cond_type=torch.tensor(m["cond"]), more_smooth=True)
        mism = [int((a != b.cpu()).sum()) for a, b in zip(gold["idx"], var.last_idx)]
        sub = m["img_sub"]
        print(f"{name} engine {eng}: token mismatches per scale {mism} of {[a.numel() for a in gold['idx']]}; "
              f"f_hat err {(var.last_f_hat.cpu() - gold['f_hat']).abs().max().item():.3e} (absmax {gold['f_hat'].abs().max().item():.2f}); "
              f"pixel err {(img[:, :, ::sub, ::sub].cpu() - gold['img_sub']).abs().max().item():.3e}", flush=True)
