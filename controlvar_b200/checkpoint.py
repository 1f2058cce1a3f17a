"""Checkpoint compatibility with the reference - pure host code.

  unwrap_state_dict   the ``{'model_state_dict': ...}`` wrapper and the DDP ``module.`` prefix that the reference's
                      save_checkpoint / load_var_weight / resume handle (train_control_var_hpu.py:411-428, 472-479, 430-435)
  load_checkpoint     ``resume``'s strict load of a ControlVAR checkpoint, or of the released VQVAE weights
                      (vae_ch160v4096z32.pth, a bare state_dict; README.md:19-24)
  load_var_weight     the surgery that initialises a ControlVAR (mask_type='interleave_append') from a plain VAR
                      checkpoint (train_control_var_hpu.py:472-534): positional table doubled to the joint
                      control+image token layout, buffers that the constructor rebuilds dropped, strict=False.

Branches that need options this package does not implement (``separator``: extra special-token rows in pos_1LC and
in the head) raise NotImplementedError, like the constructor does.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Mapping, Optional, Sequence, Union

import torch
import torch.nn as nn

StateDict = Mapping[str, torch.Tensor]


def unwrap_state_dict(obj: Union[str, Mapping]) -> "OrderedDict[str, torch.Tensor]":
    """path or loaded object -> flat state_dict without the 'model_state_dict' wrapper and 'module.' prefixes."""
    if isinstance(obj, (str, bytes)) or hasattr(obj, "__fspath__"):
        obj = torch.load(obj, map_location=torch.device("cpu"))
    if "model_state_dict" in obj.keys():                       # train_control_var_hpu.py:474-475
        obj = obj["model_state_dict"]
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in obj.items():
        out[k.replace("module.", "")] = v                       # train_control_var_hpu.py:477-478
    return out


def load_checkpoint(module: nn.Module, ckpt: Union[str, Mapping], strict: bool = True):
    """resume() / released-weights load: strict by default (train_control_var_hpu.py:430-435)."""
    return module.load_state_dict(unwrap_state_dict(ckpt), strict=strict)


def convert_var_state_dict(sd: StateDict, v_patch_nums: Sequence[int], embed_dim: int, *, interpos: bool = False,
                           separator: bool = False, generator: Optional[torch.Generator] = None
                           ) -> "OrderedDict[str, torch.Tensor]":
    """The dictionary surgery of load_var_weight for mask_type='interleave_append' (train_control_var_hpu.py:482-533).

    A VAR checkpoint has one token per position (pos_1LC of length sum(pn^2) = 680); ControlVAR interleaves a control
    and an image token map per scale (length 1360).  Default (interpos=False): the whole table is concatenated with
    itself (:519) - NOTE this is [all scales | all scales], not per-scale interleaving; it is what the reference does.
    interpos=True (:491-505): per scale, the scale's slice is written twice back to back.
    """
    if separator:
        raise NotImplementedError("separator=True (special tokens in pos_1LC / head) is not implemented")
    out = OrderedDict(sd)
    for key in ("lvl_1L", "pos_start", "attn_bias_for_masking"):     # rebuilt by the constructor (:484-485)
        del out[key]                                                 # KeyError if absent, like the reference
    pos = out["pos_1LC"]
    if interpos:
        init_std = math.sqrt(1 / embed_dim / 3)
        parts, L = [], 0
        for pn in v_patch_nums:
            pe = torch.empty(pn * pn * 2, embed_dim)
            nn.init.trunc_normal_(pe, mean=0, std=init_std, generator=generator)    # fully overwritten below
            pe[:pn * pn] = pos[:, L:L + pn * pn]
            pe[pn * pn:pn * pn * 2] = pos[:, L:L + pn * pn]
            parts.append(pe)
            L += pn * pn
        out["pos_1LC"] = torch.cat(parts, dim=0).unsqueeze(0)
    else:
        out["pos_1LC"] = torch.concat([pos, pos], dim=1)
    return out


def load_var_weight(var: nn.Module, ckpt: Union[str, Mapping], *, mask_type: str = "interleave_append",
                    interpos: bool = False, separator: bool = False, v_patch_nums: Optional[Sequence[int]] = None,
                    embed_dim: Optional[int] = None):
    """Drop-in for load_var_weight(var, args) with the argparse fields spelled out (train_control_var_hpu.py:472-534).
    Returns load_state_dict's (missing_keys, unexpected_keys); the load is strict=False as in the reference: pos_start,
    cond_embed and the other ControlVAR-only tensors keep their constructor values."""
    sd = unwrap_state_dict(ckpt)
    if mask_type == "interleave_append":
        sd = convert_var_state_dict(sd, v_patch_nums if v_patch_nums is not None else var.patch_nums,
                                    embed_dim if embed_dim is not None else var.C, interpos=interpos,
                                    separator=separator)
    return var.load_state_dict(sd, strict=False)
