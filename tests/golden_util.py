"""Helpers shared by the golden-vector tests (tests/golden/*.npz are written by oracle/make_golden.py)."""
import glob
import json
import os

import numpy as np
import torch

from controlvar_b200.config import PathConfig

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _kind(name):
    for k in ("enc", "cond", "fwd"):
        if name.startswith(k + "_"):
            return k
    return "sample"


def golden_names(kind="sample"):
    """kind: 'sample' (autoregressive_infer_cfg), 'cond' (conditional_infer_cfg), 'enc' (img_to_idxBl), 'fwd' (forward)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if _kind(n) == kind]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    cfg = PathConfig(depth=meta["depth"], patch_nums=tuple(meta["patch_nums"]), embed_dim=meta.get("embed_dim", 0),
                     heads=meta.get("heads", 0))
    if meta.get("kind") == "fwd":
        return dict(meta=meta, cfg=cfg, logits_sub=torch.from_numpy(z["logits_sub"]))
    idx = [torch.from_numpy(z[f"idx_{si}"].astype(np.int64)) for si in range(len(cfg.patch_nums))]
    if meta.get("kind") == "enc":
        out = dict(meta=meta, cfg=cfg, idx=idx, f=torch.from_numpy(z["f"]))
        if "recon_sub" in z:
            out.update(img_from_tokens_sub=torch.from_numpy(z["img_from_tokens_sub"]), recon_sub=torch.from_numpy(z["recon_sub"]),
                       h=[torch.from_numpy(z[f"h_{si}"]) for si in range(len(cfg.patch_nums) - 1)])
        return out
    if meta.get("kind") == "cond":
        forced = [torch.from_numpy(z[f"forced_{si}"].astype(np.int64)) for si in range(len(cfg.patch_nums))]
        return dict(meta=meta, cfg=cfg, idx=idx, forced=forced, f_hat=torch.from_numpy(z["f_hat"]),
                    img_sub=torch.from_numpy(z["img_sub"]))
    return dict(meta=meta, cfg=cfg, idx=idx, f_hat=torch.from_numpy(z["f_hat"]),
                img_sub=torch.from_numpy(z["img_sub"]), img_mean=float(z["img_mean"][0]),
                logits_row=torch.from_numpy(z["logits_cfg_last_row0"]))
