"""The evaluation driver around the sampling path: rank-sharded per-class sampling with
PNG dump, pixel-conditioned sampling, and the Gibbs refinement loop - host code mirroring
/root/reference/train_control_var_hpu.py:300-408 (``pix_cond_inference``, ``cls_cond_inference``, ``validate``) with the
argparse fields spelled out.  Everything numeric runs through ControlVAR / VQVAE of this package.

Differences from the reference, on purpose:
  * :398 indexes ``images[b, 256]`` (a single pixel row) where the pixel-conditioned branch (:354) saves ``images[b, 256:]``
    (the image half under the control map); both branches save the image half here.
  * no wandb branch (``save_val=False`` returns the uint8 arrays to the caller instead), no tqdm.
  * Gibbs refinement (:380-393).  As written, the reference sets ``args.c_mask = True`` for the first half-step and then
    ``args.c_img = True`` for the second WITHOUT clearing ``c_mask``; its ``pix_cond_inference`` tests ``c_mask`` first
    (:314), so the second half-step tokenises the masks again and hands ``c_img=True`` (a bool, not a token list) to
    ``conditional_infer_cfg``, which then fails at ``c_img[si]`` (control_var.py:317-321) - the loop cannot complete a
    round in the reference.  Here the second half-step is what the loop evidently intends: ``(c_mask=None, c_img=True)``,
    i.e. the IMAGE tokens are forced and the control map is re-sampled.  Anyone comparing Gibbs output with a patched
    reference must patch it the same way.  tests/test_validate_cpu.py pins this behaviour, not reference parity.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

COND_TYPES = {"mask": 0, "canny": 1, "depth": 2, "normal": 3, "none": 4}      # train_control_var_hpu.py:302


def to_uint8_hwc(images_B3HW: torch.Tensor) -> np.ndarray:
    """images.permute(0, 2, 3, 1).mul_(255).cpu().numpy().astype(np.uint8) - :351, :396 (truncation, not rounding)."""
    return images_B3HW.permute(0, 2, 3, 1).mul(255).cpu().numpy().astype(np.uint8)


def class_slice(rank: int, gpus: int, num_classes: int = 1000) -> List[int]:
    """Classes a rank samples (:365-367): equal slices, the last rank takes the remainder."""
    per = num_classes // gpus
    return list(range(per * rank, per * (rank + 1))) if rank != gpus - 1 else list(range(per * rank, num_classes))


def batch_plan(per_class: int, batch_size: int) -> List[int]:
    """Batch sizes of one class (:372-375): per_class // batch_size full batches and the remainder (empty ones skipped);
    the position in the list is the reference's loop index i (it enters the seed and the file names)."""
    assert per_class > batch_size                                       # :371
    n = per_class // batch_size
    return [batch_size if i != n else per_class - i * batch_size for i in range(n + 1)]


def _cond_tensor(cond_type, B: int, device) -> torch.Tensor:
    if isinstance(cond_type, str):
        return torch.full((B,), COND_TYPES[cond_type], dtype=torch.long, device=device)
    return cond_type.to(device=device, dtype=torch.long)


@torch.no_grad()
def cls_cond_inference(cls: int, B: int, var, cond_type, guidance_scale: Sequence[float], top_k: int, top_p: float,
                       seed: int) -> torch.Tensor:
    """:327-336 - class + condition-type conditioned sampling of B images (control map on top, image below)."""
    dev = var.device
    conditions = torch.full((B,), cls, dtype=torch.long, device=dev)
    return var.autoregressive_infer_cfg(B=B, label_B=conditions, cond_type=_cond_tensor(cond_type, B, dev),
                                        cfg=guidance_scale[0], top_k=top_k, top_p=top_p, g_seed=seed)


@torch.no_grad()
def pix_cond_inference(images: torch.Tensor, masks: torch.Tensor, conditions, cond_type, B: int, var, vqvae,
                       c_mask, c_img, guidance_scale: Sequence[float], top_k: int, top_p: float, seed: int,
                       v_patch_nums: Sequence[int]) -> torch.Tensor:
    """:300-325 - pixel-level control: the condition map (c_mask) or the image (c_img) is tokenised by the VQVAE and its
    tokens are teacher-forced into conditional_infer_cfg.  images / masks in [-1, 1]."""
    dev = var.device
    if isinstance(conditions, int):
        conditions = torch.full((B,), conditions, dtype=torch.long, device=dev)
    else:
        conditions = conditions.to(dev)
    ct = _cond_tensor(cond_type, B, dev)
    # :314-320: `if c_mask: ... elif c_img: ... else: both None`.  As in the reference, when c_mask is set a truthy c_img
    # flag is passed on UNCHANGED (it is not a token list then, so the call below would fail): callers set one of the two.
    if c_mask:
        c_mask = vqvae.img_to_idxBl(masks.to(dev), v_patch_nums=v_patch_nums)
        c_img = c_img if isinstance(c_img, (list, tuple)) else None
    elif c_img:
        c_mask, c_img = None, vqvae.img_to_idxBl(images.to(dev), v_patch_nums=v_patch_nums)
    else:
        c_mask, c_img = None, None
    return var.conditional_infer_cfg(B=B, label_B=conditions, cfg=tuple(guidance_scale), top_k=top_k, top_p=top_p,
                                     g_seed=seed, c_mask=c_mask, c_img=c_img, cond_type=ct)


def _save_pngs(arr_BHWC: np.ndarray, paths: Sequence[str]) -> None:
    from PIL import Image
    for a, p in zip(arr_BHWC, paths):
        Image.fromarray(a).save(p)


@torch.no_grad()
def validate_classes(var, vqvae, project_dir: str, *, rank: int = 0, gpus: int = 1, batch_size: int = 25,
                     guidance_scale: Sequence[float] = (6, 6, 6), top_k: int = 900, top_p: float = 0.96, seed: int = 42,
                     gibbs: int = 0, cond_type="depth", per_class: int = 50, classes: Optional[Iterable[int]] = None,
                     save_val: bool = True, num_classes: int = 1000) -> Dict[int, List[np.ndarray]]:
    """The class-conditional branch of validate() (:363-408): every class of this rank's slice is sampled per_class times
    in batches, optionally refined by `gibbs` rounds of (control map -> image, image -> control map) pixel-conditioned
    resampling, converted to uint8 and written to <project_dir>/cfg_<g0>/<cls>/<n>.png (image half).
    Returns {cls: [uint8 (B, side, side, 3) per batch]} (also when save_val is False, instead of the wandb branch)."""
    classes = class_slice(rank, gpus, num_classes) if classes is None else list(classes)
    pn = var.patch_nums
    out: Dict[int, List[np.ndarray]] = {}
    for cls in classes:
        cls_dir = os.path.join(project_dir, f"cfg_{guidance_scale[0]}", f"{cls}")
        if save_val:
            os.makedirs(cls_dir, exist_ok=True)
        for i, B in enumerate(batch_plan(per_class, batch_size)):
            if B == 0:
                continue
            seed = seed + i * (cls + 1)                                 # :377 (cumulative, as written)
            images = cls_cond_inference(cls, B, var, cond_type, guidance_scale, top_k, top_p, seed)
            side = images.shape[-1]
            for _ in range(gibbs):                                      # :380-393
                masks, imgs = images[:, :, :side, :], images[:, :, side:, :]
                masks, imgs = (masks - 0.5) / 0.5, (imgs - 0.5) / 0.5
                images = pix_cond_inference(imgs, masks, cls, cond_type, B, var, vqvae, True, None, guidance_scale, top_k,
                                            top_p, seed, pn)
                masks, imgs = images[:, :, :side, :], images[:, :, side:, :]
                masks, imgs = (masks - 0.5) / 0.5, (imgs - 0.5) / 0.5
                images = pix_cond_inference(imgs, masks, cls, cond_type, B, var, vqvae, None, True, guidance_scale, top_k,
                                            top_p, seed, pn)
            arr = to_uint8_hwc(images)[:, side:]                        # the image half (see the module docstring)
            out.setdefault(cls, []).append(arr)
            if save_val:
                _save_pngs(arr, [os.path.join(cls_dir, f"{i * batch_size + b}.png") for b in range(B)])
    return out


@torch.no_grad()
def validate_pixel_conditioned(var, vqvae, dataloader, project_dir: str, *, val_cond: str = "depth", rank: int = 0,
                               guidance_scale: Sequence[float] = (6, 6, 6), top_k: int = 900, top_p: float = 0.96,
                               seed: int = 42, c_mask=None, c_img=None, save_val: bool = True) -> List[np.ndarray]:
    """The pixel-conditioned branch of validate() (:343-361): batches of {'image', 'mask', 'cls', 'type'} from a loader."""
    g = guidance_scale
    save_path = os.path.join(project_dir, f"cfg_{g[0]}_{g[1]}_{g[2]}_{val_cond}", f"{rank}")
    if save_val:
        os.makedirs(save_path, exist_ok=True)
    out = []
    for batch_idx, batch in enumerate(dataloader):
        images, masks, conditions, cond_type = batch["image"], batch["mask"], batch["cls"], batch["type"]
        B = masks.shape[0]
        res = pix_cond_inference(images, masks, conditions, cond_type, B, var, vqvae, c_mask, c_img, g, top_k, top_p, seed,
                                 var.patch_nums)
        side = res.shape[-1]
        arr = to_uint8_hwc(res)[:, side:]                               # images[b, 256:]  (:354)
        out.append(arr)
        if save_val:
            _save_pngs(arr, [os.path.join(save_path, f"{batch_idx * B + b}.png") for b in range(B)])
    return out
