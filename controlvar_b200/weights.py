"""state_dict key layout of the reference and a portable synthetic-weight generator.

The on-disk contract is the reference's ``state_dict`` key set:
  ControlVAR: /root/reference/models/control_var.py:70-213, models/basic_var.py:73-77,190-198
  VQVAE:      /root/reference/models/vqvae.py:39-48, models/vae_modules.py:163-208, models/quant.py:24-37

No released checkpoint is reachable offline, so parity tests and the benchmark use *synthetic* weights.
They are produced by a counter-based integer hash (no transcendental functions, no library RNG), so the very
same bits come out in the build container (where the goldens are made by loading them into the unmodified
reference) and on the GPU box (where the CUDA path consumes them).  Scales follow PyTorch's default
initialisers, which is what the reference itself runs with when no checkpoint is loaded
(``special_init``/``init_weights`` are never called).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config import PathConfig

_M64 = (1 << 64) - 1


def _s64(x: int) -> int:
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


_C1 = _s64(0x9E3779B97F4A7C15)
_C2 = _s64(0xBF58476D1CE4E5B9)
_C3 = _s64(0x94D049BB133111EB)


def hash_uniform(n: int, seed: int, stream: int, device=None) -> torch.Tensor:
    """n float32 values in [-1, 1), bit-identical on every host and device (int64 wrap-around arithmetic only)."""
    i = torch.arange(n, dtype=torch.int64, device=device)
    z = i * _C1 + _s64((seed * 0x632BE59BD9B4E019 + stream * 0xD1342543DE82EF95) & _M64)
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _C2     # logical shifts (int64 '>>' is arithmetic)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _C3
    z = z ^ ((z >> 31) & ((1 << 33) - 1))
    k = (z >> 40) & 0xFFFFFF                       # 24 bits
    u = k.to(torch.float32) * (1.0 / 16777216.0)    # exact
    return u * 2.0 - 1.0                            # exact


_DEVICE = None   # set by synthetic_*_state_dict(device=...) for the duration of one call


def _fill(shape, bound, seed, key, center=0.0):
    n = 1
    for s in shape:
        n *= s
    t = hash_uniform(n, seed, zlib.crc32(key.encode()), _DEVICE) * float(bound)
    if center != 0.0:
        t = t + float(center)
    return t.reshape(shape).contiguous()


# ---------------------------------------------------------------------------------------------- key specs
def var_key_shapes(cfg: PathConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """ControlVAR.state_dict() keys in registration order (released branch: multi_cond, mask_factor 2)."""
    C, L, H = cfg.C, cfg.L, cfg.num_heads
    ks: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    ks["pos_start"] = (1, cfg.first_l, C)
    ks["pos_1LC"] = (1, L, C)
    ks["lvl_1L"] = (1, L)
    ks["attn_bias_for_masking"] = (1, 1, L, L)
    ks["word_embed.weight"] = (C, cfg.Cvae)
    ks["word_embed.bias"] = (C,)
    ks["class_emb.weight"] = (cfg.num_classes + 1, C)
    ks["lvl_embed.weight"] = (len(cfg.patch_nums), C)
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        if cfg.cos_attn:
            ks[p + "attn.scale_mul_1H11"] = (1, H, 1, 1)
        ks[p + "attn.q_bias"] = (C,)
        ks[p + "attn.v_bias"] = (C,)
        ks[p + "attn.zero_k_bias"] = (C,)
        ks[p + "attn.mat_qkv.weight"] = (3 * C, C)
        ks[p + "attn.proj.weight"] = (C, C)
        ks[p + "attn.proj.bias"] = (C,)
        ks[p + "ffn.fc1.weight"] = (cfg.hidden, C)
        ks[p + "ffn.fc1.bias"] = (cfg.hidden,)
        ks[p + "ffn.fc2.weight"] = (C, cfg.hidden)
        ks[p + "ffn.fc2.bias"] = (C,)
        ks[p + "ada_lin.1.weight"] = (6 * C, C)
        ks[p + "ada_lin.1.bias"] = (6 * C,)
    ks["head_nm.ada_lin.1.weight"] = (2 * C, C)
    ks["head_nm.ada_lin.1.bias"] = (2 * C,)
    ks["head.weight"] = (cfg.vocab_size, C)
    ks["head.bias"] = (cfg.vocab_size,)
    if cfg.multi_cond:
        ks["cond_embed.weight"] = (5, C)
    return ks


def _resblock(ks, p, cin, cout):
    ks[p + "norm1.weight"] = (cin,)
    ks[p + "norm1.bias"] = (cin,)
    ks[p + "conv1.weight"] = (cout, cin, 3, 3)
    ks[p + "conv1.bias"] = (cout,)
    ks[p + "norm2.weight"] = (cout,)
    ks[p + "norm2.bias"] = (cout,)
    ks[p + "conv2.weight"] = (cout, cout, 3, 3)
    ks[p + "conv2.bias"] = (cout,)
    if cin != cout:
        ks[p + "nin_shortcut.weight"] = (cout, cin, 1, 1)
        ks[p + "nin_shortcut.bias"] = (cout,)


def _attnblock(ks, p, c):
    ks[p + "norm.weight"] = (c,)
    ks[p + "norm.bias"] = (c,)
    ks[p + "qkv.weight"] = (3 * c, c, 1, 1)
    ks[p + "qkv.bias"] = (3 * c,)
    ks[p + "proj_out.weight"] = (c, c, 1, 1)
    ks[p + "proj_out.bias"] = (c,)


def decoder_plan(cfg: PathConfig):
    """Static walk of Decoder.forward (vae_modules.py:210-226) as a list of (op, prefix, cin, cout) records."""
    ch, mult, nrb = cfg.vae_ch, cfg.vae_ch_mult, cfg.vae_num_res_blocks
    nres = len(mult)
    block_in = ch * mult[-1]
    plan = [("conv3", "decoder.conv_in", cfg.Cvae, block_in)]
    plan.append(("res", "decoder.mid.block_1.", block_in, block_in))
    plan.append(("attn", "decoder.mid.attn_1.", block_in, block_in))
    plan.append(("res", "decoder.mid.block_2.", block_in, block_in))
    for lvl in reversed(range(nres)):
        block_out = ch * mult[lvl]
        for ib in range(nrb + 1):
            plan.append(("res", f"decoder.up.{lvl}.block.{ib}.", block_in, block_out))
            block_in = block_out
            if lvl == nres - 1:
                plan.append(("attn", f"decoder.up.{lvl}.attn.{ib}.", block_in, block_in))
        if lvl != 0:
            plan.append(("up", f"decoder.up.{lvl}.upsample.conv", block_in, block_in))
    plan.append(("out", "decoder.", block_in, 3))
    return plan


def encoder_plan(cfg: PathConfig):
    """Static walk of Encoder.forward (vae_modules.py:145-160) as (op, prefix, cin, cout) records; 'down' is
    Downsample2x (vae_modules.py:31-37), 'out' is norm_out + SiLU + conv_out."""
    ch, mult, nrb = cfg.vae_ch, cfg.vae_ch_mult, cfg.vae_num_res_blocks
    nres = len(mult)
    in_mult = (1,) + tuple(mult)
    plan = [("conv_in", "encoder.conv_in", 3, ch)]
    block_in = ch
    for lvl in range(nres):
        block_in = ch * in_mult[lvl]
        block_out = ch * mult[lvl]
        for ib in range(nrb):
            plan.append(("res", f"encoder.down.{lvl}.block.{ib}.", block_in, block_out))
            block_in = block_out
            if lvl == nres - 1:
                plan.append(("attn", f"encoder.down.{lvl}.attn.{ib}.", block_in, block_in))
        if lvl != nres - 1:
            plan.append(("down", f"encoder.down.{lvl}.downsample.conv", block_in, block_in))
    plan.append(("res", "encoder.mid.block_1.", block_in, block_in))
    plan.append(("attn", "encoder.mid.attn_1.", block_in, block_in))
    plan.append(("res", "encoder.mid.block_2.", block_in, block_in))
    plan.append(("out", "encoder.", block_in, cfg.Cvae))
    return plan


def vae_key_shapes(cfg: PathConfig, with_encoder: bool = True) -> "OrderedDict[str, Tuple[int, ...]]":
    ch, mult, nrb, z = cfg.vae_ch, cfg.vae_ch_mult, cfg.vae_num_res_blocks, cfg.Cvae
    nres = len(mult)
    ks: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    if with_encoder:
        ks["encoder.conv_in.weight"] = (ch, 3, 3, 3)
        ks["encoder.conv_in.bias"] = (ch,)
        in_mult = (1,) + tuple(mult)
        block_in = ch
        for lvl in range(nres):
            block_in = ch * in_mult[lvl]
            block_out = ch * mult[lvl]
            for ib in range(nrb):
                _resblock(ks, f"encoder.down.{lvl}.block.{ib}.", block_in, block_out)
                block_in = block_out
            if lvl == nres - 1:
                for ib in range(nrb):
                    _attnblock(ks, f"encoder.down.{lvl}.attn.{ib}.", block_in)
            if lvl != nres - 1:
                ks[f"encoder.down.{lvl}.downsample.conv.weight"] = (block_in, block_in, 3, 3)
                ks[f"encoder.down.{lvl}.downsample.conv.bias"] = (block_in,)
        _resblock(ks, "encoder.mid.block_1.", block_in, block_in)
        _attnblock(ks, "encoder.mid.attn_1.", block_in)
        _resblock(ks, "encoder.mid.block_2.", block_in, block_in)
        ks["encoder.norm_out.weight"] = (block_in,)
        ks["encoder.norm_out.bias"] = (block_in,)
        ks["encoder.conv_out.weight"] = (z, block_in, 3, 3)
        ks["encoder.conv_out.bias"] = (z,)
    # decoder (registration order of Decoder.__init__: conv_in, mid, up[0..], norm_out, conv_out)
    block_in = ch * mult[-1]
    ks["decoder.conv_in.weight"] = (block_in, z, 3, 3)
    ks["decoder.conv_in.bias"] = (block_in,)
    _resblock(ks, "decoder.mid.block_1.", block_in, block_in)
    _attnblock(ks, "decoder.mid.attn_1.", block_in)
    _resblock(ks, "decoder.mid.block_2.", block_in, block_in)
    per_level = {}
    for lvl in reversed(range(nres)):
        sub: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
        block_out = ch * mult[lvl]
        cin = block_in
        for ib in range(nrb + 1):
            _resblock(sub, f"decoder.up.{lvl}.block.{ib}.", cin, block_out)
            cin = block_out
        if lvl == nres - 1:
            for ib in range(nrb + 1):
                _attnblock(sub, f"decoder.up.{lvl}.attn.{ib}.", block_out)
        block_in = block_out
        if lvl != 0:
            sub[f"decoder.up.{lvl}.upsample.conv.weight"] = (block_in, block_in, 3, 3)
            sub[f"decoder.up.{lvl}.upsample.conv.bias"] = (block_in,)
        per_level[lvl] = sub
    for lvl in range(nres):
        ks.update(per_level[lvl])
    ks["decoder.norm_out.weight"] = (block_in,)
    ks["decoder.norm_out.bias"] = (block_in,)
    ks["decoder.conv_out.weight"] = (3, block_in, 3, 3)
    ks["decoder.conv_out.bias"] = (3,)
    ks["quantize.ema_vocab_hit_SV"] = (len(cfg.patch_nums), cfg.vocab_size)
    for k in range(cfg.share_quant_resi):
        ks[f"quantize.quant_resi.qresi_ls.{k}.weight"] = (z, z, 3, 3)
        ks[f"quantize.quant_resi.qresi_ls.{k}.bias"] = (z,)
    ks["quantize.embedding.weight"] = (cfg.vocab_size, z)
    ks["quant_conv.weight"] = (z, z, 3, 3)
    ks["quant_conv.bias"] = (z,)
    ks["post_quant_conv.weight"] = (z, z, 3, 3)
    ks["post_quant_conv.bias"] = (z,)
    return ks


# ------------------------------------------------------------------------------------- derived buffers
def lvl_1L(cfg: PathConfig) -> torch.Tensor:
    """control_var.py:158-166: level id of every position of the token pyramid."""
    return torch.cat([torch.full((n,), i, dtype=torch.int64) for i, n in enumerate(cfg.scale_lens)]).view(1, cfg.L)


def attn_bias_for_masking(cfg: PathConfig) -> torch.Tensor:
    """control_var.py:168: block-causal additive bias (0 where query level >= key level, -inf elsewhere)."""
    d = lvl_1L(cfg).view(1, cfg.L, 1)
    dT = d.transpose(1, 2)
    return torch.where(d >= dT, 0.0, -torch.inf).reshape(1, 1, cfg.L, cfg.L).contiguous()


# -------------------------------------------------------------------------------- synthetic generators
def synthetic_var_state_dict(cfg: PathConfig, seed: int = 0, device=None) -> Dict[str, torch.Tensor]:
    global _DEVICE
    _DEVICE = device
    try:
        return _synthetic_var_state_dict(cfg, seed)
    finally:
        _DEVICE = None


def _synthetic_var_state_dict(cfg: PathConfig, seed: int) -> Dict[str, torch.Tensor]:
    C = cfg.C
    sd: Dict[str, torch.Tensor] = OrderedDict()
    emb_bound = 1.0 / math.sqrt(C)      # uniform with std sqrt(1/C/3), cf. control_var.py:77
    for key, shape in var_key_shapes(cfg).items():
        if key == "lvl_1L":
            sd[key] = lvl_1L(cfg).to(_DEVICE) if _DEVICE else lvl_1L(cfg)
        elif key == "attn_bias_for_masking":
            sd[key] = attn_bias_for_masking(cfg).to(_DEVICE) if _DEVICE else attn_bias_for_masking(cfg)
        elif key in ("pos_start", "pos_1LC", "class_emb.weight", "lvl_embed.weight", "cond_embed.weight"):
            sd[key] = _fill(shape, emb_bound, seed, key)
        elif key.endswith("zero_k_bias"):
            sd[key] = torch.zeros(shape, device=_DEVICE)
        elif key.endswith("q_bias") or key.endswith("v_bias"):
            sd[key] = _fill(shape, 0.1, seed, key)          # zeros at init; trained checkpoints carry values
        elif key.endswith("scale_mul_1H11"):
            sd[key] = _fill(shape, 0.25, seed, key, center=math.log(4.0))
        elif key.endswith(".weight"):
            sd[key] = _fill(shape, 1.0 / math.sqrt(shape[1]), seed, key)        # nn.Linear default
        elif key.endswith(".bias"):
            wshape = var_key_shapes(cfg)[key[:-4] + "weight"]
            sd[key] = _fill(shape, 1.0 / math.sqrt(wshape[1]), seed, key)
        else:
            raise KeyError(key)
    return sd


def synthetic_vae_state_dict(cfg: PathConfig, seed: int = 0, with_encoder: bool = True, device=None) -> Dict[str, torch.Tensor]:
    global _DEVICE
    _DEVICE = device
    try:
        return _synthetic_vae_state_dict(cfg, seed, with_encoder)
    finally:
        _DEVICE = None


def _synthetic_vae_state_dict(cfg: PathConfig, seed: int, with_encoder: bool) -> Dict[str, torch.Tensor]:
    shapes = vae_key_shapes(cfg, with_encoder)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for key, shape in shapes.items():
        if key == "quantize.ema_vocab_hit_SV":
            sd[key] = torch.zeros(shape, device=_DEVICE)
        elif key == "quantize.embedding.weight":
            sd[key] = _fill(shape, math.sqrt(3.0), seed, key)                   # unit variance like N(0,1)
        elif ".norm" in key or "norm_out" in key:
            sd[key] = _fill(shape, 0.1, seed, key, center=1.0 if key.endswith("weight") else 0.0)
        elif key.endswith(".weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            sd[key] = _fill(shape, 1.0 / math.sqrt(fan_in), seed, key)          # nn.Conv2d default
        elif key.endswith(".bias"):
            w = shapes[key[:-4] + "weight"]
            sd[key] = _fill(shape, 1.0 / math.sqrt(w[1] * w[2] * w[3]), seed, key)
        else:
            raise KeyError(key)
    return sd


def synthetic_image(B: int, side: int = 256, seed: int = 0, device=None) -> torch.Tensor:
    """(B, 3, side, side) fp32 image in [-1, 1] with structure at two scales (8x8 blocks + pixel noise), from the same
    integer hash as the weights: identical bits in the build container (golden generation) and on the GPU box."""
    assert side % 8 == 0
    s8 = side // 8
    coarse = hash_uniform(B * 3 * s8 * s8, seed, 0x1ACE, device).reshape(B, 3, s8, s8)
    coarse = coarse.repeat_interleave(8, dim=2).repeat_interleave(8, dim=3)
    fine = hash_uniform(B * 3 * side * side, seed, 0xF19E, device).reshape(B, 3, side, side)
    return (coarse * 0.75 + fine * 0.25).contiguous()


def synthetic_teacher_input(cfg: PathConfig, B: int, seed: int = 0, device=None) -> torch.Tensor:
    """(B, L - first_l, Cvae) teacher-forcing input of ControlVAR.forward (what VectorQuantizer2.idxBl_to_var_input
    produces from ground-truth tokens), here hash-generated with the spread of a codebook entry."""
    n = B * (cfg.L - cfg.first_l) * cfg.Cvae
    return (hash_uniform(n, seed, 0x7EAC, device) * math.sqrt(3.0)).reshape(B, cfg.L - cfg.first_l, cfg.Cvae).contiguous()
