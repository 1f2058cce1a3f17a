"""ORACLE - TEST INFRASTRUCTURE ONLY.  Never imported by the product package ``controlvar_b200``.

CPU (fp32, PyTorch/ATen) restatement of the reference algorithm for the next-scale sampling hot path of
lxa9867/ControlVAR: ``ControlVAR.autoregressive_infer_cfg`` (released branch: multi_cond=True, mask_factor=2)
and ``VQVAE.fhat_to_img``, plus the pixel-conditioned caller of the same path:
``ControlVAR.conditional_infer_cfg`` and ``VQVAE.img_to_idxBl`` (encoder + multi-scale residual quantiser).  It is written functionally over a ``state_dict`` (no nn.Module, no reference
import) so that it can travel to the GPU box, where /root/reference does not exist.

Parity pinning: the reference ships NO tests / golden vectors for this path, so the
oracle is pinned against outputs of the *unmodified reference itself*, run in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/`` (token indices of every scale bit-exact,
images bit-exact on CPU).  ``tests/test_oracle_golden.py`` re-checks that on every run.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Every function cites the reference file:line it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------ transformer
def ada_lin(cond_BD: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """nn.Sequential(SiLU, Linear) - models/basic_var.py:197-198, control_var.py:697."""
    return F.linear(F.silu(cond_BD), w, b)


def ln_modulate(x: Tensor, scale: Tensor, shift: Tensor, eps: float = 1e-6) -> Tensor:
    """ln_wo_grad(x).mul(scale.add(1)).add_(shift) - models/basic_var.py:208-209, control_var.py:701."""
    C = x.shape[-1]
    return F.layer_norm(x, (C,), None, None, eps).mul(scale.add(1)).add_(shift)


def self_attention(x: Tensor, sd: Dict[str, Tensor], p: str, H: int, cache: Optional[dict], cos_attn: bool,
                   scale: float, attn_bias: Optional[Tensor] = None) -> Tensor:
    """SelfAttention.forward, fp32 SDPA branch with KV cache - models/basic_var.py:89-119."""
    B, L, C = x.shape
    hd = C // H
    bias = torch.cat((sd[p + "q_bias"], sd[p + "zero_k_bias"], sd[p + "v_bias"]))
    qkv = F.linear(x, sd[p + "mat_qkv.weight"], bias).view(B, L, 3, H, hd)
    q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(dim=0)                       # BHLc
    if cos_attn:
        scale_mul = sd[p + "scale_mul_1H11"].clamp_max(math.log(100)).exp()  # :100 (max_scale_mul = log(100))
        q = F.normalize(q, dim=-1).mul(scale_mul)
        k = F.normalize(k, dim=-1)
    if cache is not None:
        if cache.get("k") is None:
            cache["k"], cache["v"] = k, v
        else:
            k = cache["k"] = torch.cat((cache["k"], k), dim=2)
            v = cache["v"] = torch.cat((cache["v"], v), dim=2)
    oup = F.scaled_dot_product_attention(query=q, key=k, value=v, scale=scale, attn_mask=attn_bias,
                                         dropout_p=0.0).transpose(1, 2).reshape(B, L, C)
    return F.linear(oup, sd[p + "proj.weight"], sd[p + "proj.bias"])


def ffn(x: Tensor, sd: Dict[str, Tensor], p: str) -> Tensor:
    """FFN.forward, unfused branch - models/basic_var.py:51."""
    return F.linear(F.gelu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"]), approximate="tanh"),
                    sd[p + "fc2.weight"], sd[p + "fc2.bias"])


def adaln_block(x: Tensor, cond_BD: Tensor, sd: Dict[str, Tensor], p: str, H: int, cache: Optional[dict],
                cos_attn: bool, scale: float, attn_bias: Optional[Tensor] = None) -> Tensor:
    """AdaLNSABlock.forward (drop_path is identity in eval) - models/basic_var.py:203-210."""
    C = x.shape[-1]
    g1, g2, s1, s2, b1, b2 = ada_lin(cond_BD, sd[p + "ada_lin.1.weight"], sd[p + "ada_lin.1.bias"]) \
        .view(-1, 1, 6, C).unbind(2)
    x = x + self_attention(ln_modulate(x, s1, b1), sd, p + "attn.", H, cache, cos_attn, scale, attn_bias).mul_(g1)
    x = x + ffn(ln_modulate(x, s2, b2), sd, p + "ffn.").mul(g2)
    return x


def get_logits(x: Tensor, cond_BD: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    """ControlVAR.get_logits + AdaLNBeforeHead.forward - models/control_var.py:215-221, 699-701."""
    C = x.shape[-1]
    scale, shift = ada_lin(cond_BD, sd["head_nm.ada_lin.1.weight"], sd["head_nm.ada_lin.1.bias"]) \
        .view(-1, 1, 2, C).unbind(2)
    return F.linear(ln_modulate(x.float(), scale, shift).float(), sd["head.weight"], sd["head.bias"]).float()


# --------------------------------------------------------------------------------------------- sampling
def cfg_combine(logits_2BlV: Tensor, B: int, t: float) -> Tensor:
    """(1+t) * logits[:B] - t * logits[B:] - models/control_var.py:501-502."""
    return (1 + t) * logits_2BlV[:B] - t * logits_2BlV[B:]


def mask_top_k_top_p_(logits_BlV: Tensor, top_k: int, top_p: float) -> Tensor:
    """In-place masking part of sample_with_top_k_top_p_ - models/helpers.py:8-15."""
    if top_k > 0:
        kth = logits_BlV.topk(top_k, largest=True, sorted=False, dim=-1)[0].amin(dim=-1, keepdim=True)
        logits_BlV.masked_fill_(logits_BlV < kth, -torch.inf)
    if top_p > 0:
        sorted_logits, sorted_idx = logits_BlV.sort(dim=-1, descending=False)
        rm = sorted_logits.softmax(dim=-1).cumsum_(dim=-1) <= (1 - top_p)
        rm[..., -1:] = False
        logits_BlV.masked_fill_(rm.scatter(sorted_idx.ndim - 1, sorted_idx, rm), -torch.inf)
    return logits_BlV


def multinomial1_with_noise(probs_NV: Tensor, q_NV: Tensor) -> Tensor:
    """torch.multinomial(probs, 1, replacement=True, generator=g) == argmax(probs / q), q ~ Exp(1) drawn as
    empty_like(probs).exponential_(1, g)  (ATen multinomial fast path; identity checked in SURVEY.md section 4
    and again by tests/test_oracle_golden.py).  Taking q explicitly lets CPU and GPU runs share one noise."""
    return torch.argmax(probs_NV / q_NV, dim=-1)


def sample_with_top_k_top_p_(logits_BlV: Tensor, top_k: int, top_p: float, q_NV: Tensor) -> Tensor:
    """sample_with_top_k_top_p_ with num_samples=1 - models/helpers.py:6-19.  Returns (B, l) int64."""
    B, l, V = logits_BlV.shape
    mask_top_k_top_p_(logits_BlV, top_k, top_p)
    return multinomial1_with_noise(logits_BlV.softmax(dim=-1).view(-1, V), q_NV).view(B, l)


def sampling_margin(logits_BlV_masked: Tensor, q_NV: Tensor) -> Tensor:
    """Relative gap between the best and the runner-up of probs/q per row (test helper: rows whose gap is below
    the numerical resolution of the path are 'ambiguous draws')."""
    V = logits_BlV_masked.shape[-1]
    r = logits_BlV_masked.softmax(dim=-1).view(-1, V) / q_NV
    top2 = r.topk(2, dim=-1)[0]
    return (top2[:, 0] - top2[:, 1]) / top2[:, 0]


# ------------------------------------------------------------------------------------- multi-scale VQ
def phi_index(si: int, SN: int, K: int = 4) -> int:
    """PhiPartiallyShared.__getitem__(si/(SN-1)) - models/quant.py:282-293."""
    ticks = np.linspace(1 / 3 / K, 1 - 1 / 3 / K, K) if K == 4 else np.linspace(1 / 2 / K, 1 - 1 / 2 / K, K)
    return int(np.argmin(np.abs(ticks - si / (SN - 1))).item())


def phi(h: Tensor, w: Tensor, b: Tensor, ratio: float = 0.5) -> Tensor:
    """Phi.forward - models/quant.py:269-270."""
    return h.mul(1 - ratio) + F.conv2d(h, w, b, stride=1, padding=1).mul_(ratio)


def get_next_autoregressive_input(si: int, patch_nums: Sequence[int], f_hat: Tensor, h_BChw: Tensor,
                                  vsd: Dict[str, Tensor], K: int = 4):
    """VectorQuantizer2.get_next_autoregressive_input - models/quant.py:243-260 (f_hat updated in place)."""
    SN = len(patch_nums)
    HW = patch_nums[-1]
    k = phi_index(si, SN, K)
    w, b = vsd[f"quantize.quant_resi.qresi_ls.{k}.weight"], vsd[f"quantize.quant_resi.qresi_ls.{k}.bias"]
    if si != SN - 1:
        h = phi(F.interpolate(h_BChw, size=(HW, HW), mode="bicubic"), w, b)
        f_hat.add_(h)
        return f_hat, F.interpolate(f_hat, size=(patch_nums[si + 1], patch_nums[si + 1]), mode="area")
    h = phi(h_BChw, w, b)
    f_hat.add_(h)
    return f_hat, f_hat


def vq_nearest(z_NC: Tensor, emb: Tensor) -> Tensor:
    """L2 nearest codebook entry - models/quant.py:203-206."""
    d = torch.sum(z_NC.square(), dim=1, keepdim=True) + torch.sum(emb.square(), dim=1, keepdim=False)
    d.addmm_(z_NC, emb.T, alpha=-2, beta=1)
    return torch.argmin(d, dim=1)


def f_to_idxBl(f_BChw: Tensor, patch_nums: Sequence[int], vsd: Dict[str, Tensor], K: int = 4) -> List[Tensor]:
    """VectorQuantizer2.f_to_idxBl_or_fhat(to_fhat=False) - models/quant.py:184-215."""
    B, C, H, W = f_BChw.shape
    f_rest = f_BChw.detach().clone()
    f_hat = torch.zeros_like(f_rest)
    emb = vsd["quantize.embedding.weight"]
    SN = len(patch_nums)
    out = []
    for si, pn in enumerate(patch_nums):
        z_NC = (F.interpolate(f_rest, size=(pn, pn), mode="area") if si != SN - 1 else f_rest) \
            .permute(0, 2, 3, 1).reshape(-1, C)
        idx_N = vq_nearest(z_NC, emb)
        idx_Bhw = idx_N.view(B, pn, pn)
        h = F.embedding(idx_Bhw, emb).permute(0, 3, 1, 2)
        h = F.interpolate(h, size=(H, W), mode="bicubic").contiguous() if si != SN - 1 else h.contiguous()
        k = phi_index(si, SN, K)
        h = phi(h, vsd[f"quantize.quant_resi.qresi_ls.{k}.weight"], vsd[f"quantize.quant_resi.qresi_ls.{k}.bias"])
        f_hat.add_(h)
        f_rest.sub_(h)
        out.append(idx_N.reshape(B, pn * pn))
    return out


# --------------------------------------------------------------------------------------------- decoder
def _gn(x: Tensor, vsd, p: str) -> Tensor:
    """Normalize = GroupNorm(32, eps=1e-6, affine) - models/vae_modules.py:18-19."""
    return F.group_norm(x, 32, vsd[p + ".weight"], vsd[p + ".bias"], 1e-6)


def _conv(x: Tensor, vsd, p: str, padding: int) -> Tensor:
    return F.conv2d(x, vsd[p + ".weight"], vsd[p + ".bias"], stride=1, padding=padding)


def resnet_block(x: Tensor, vsd, p: str) -> Tensor:
    """ResnetBlock.forward - models/vae_modules.py:57-60."""
    h = _conv(F.silu(_gn(x, vsd, p + "norm1"), inplace=True), vsd, p + "conv1", 1)
    h = _conv(F.silu(_gn(h, vsd, p + "norm2"), inplace=True), vsd, p + "conv2", 1)
    sc = _conv(x, vsd, p + "nin_shortcut", 0) if (p + "nin_shortcut.weight") in vsd else x
    return sc + h


def attn_block(x: Tensor, vsd, p: str) -> Tensor:
    """AttnBlock.forward (single head over H*W positions) - models/vae_modules.py:73-92."""
    qkv = _conv(_gn(x, vsd, p + "norm"), vsd, p + "qkv", 0)
    B, _, H, W = qkv.shape
    C = x.shape[1]
    q, k, v = qkv.reshape(B, 3, C, H, W).unbind(1)
    q = q.view(B, C, H * W).contiguous().permute(0, 2, 1).contiguous()
    k = k.view(B, C, H * W).contiguous()
    w = torch.bmm(q, k).mul_(int(C) ** (-0.5))
    w = F.softmax(w, dim=2)
    v = v.view(B, C, H * W).contiguous()
    w = w.permute(0, 2, 1).contiguous()
    h = torch.bmm(v, w).view(B, C, H, W).contiguous()
    return x + _conv(h, vsd, p + "proj_out", 0)


def decoder_forward(z: Tensor, vsd: Dict[str, Tensor], num_resolutions: int = 5, num_res_blocks: int = 2) -> Tensor:
    """Decoder.forward - models/vae_modules.py:210-226."""
    h = _conv(z, vsd, "decoder.conv_in", 1)
    h = resnet_block(h, vsd, "decoder.mid.block_1.")
    h = attn_block(h, vsd, "decoder.mid.attn_1.")
    h = resnet_block(h, vsd, "decoder.mid.block_2.")
    for lvl in reversed(range(num_resolutions)):
        for ib in range(num_res_blocks + 1):
            h = resnet_block(h, vsd, f"decoder.up.{lvl}.block.{ib}.")
            if f"decoder.up.{lvl}.attn.{ib}.norm.weight" in vsd:
                h = attn_block(h, vsd, f"decoder.up.{lvl}.attn.{ib}.")
        if lvl != 0:
            h = _conv(F.interpolate(h, scale_factor=2, mode="nearest"), vsd, f"decoder.up.{lvl}.upsample.conv", 1)
    return _conv(F.silu(_gn(h, vsd, "decoder.norm_out"), inplace=True), vsd, "decoder.conv_out", 1)


def fhat_to_img(f_hat: Tensor, vsd: Dict[str, Tensor]) -> Tensor:
    """VQVAE.fhat_to_img - models/vqvae.py:88-89."""
    return decoder_forward(_conv(f_hat, vsd, "post_quant_conv", 1), vsd).clamp_(-1, 1)


# --------------------------------------------------------------------------------------- encoder (f-1)
def encoder_forward(x: Tensor, vsd: Dict[str, Tensor], num_resolutions: int = 5, num_res_blocks: int = 2) -> Tensor:
    """Encoder.forward - models/vae_modules.py:145-160; Downsample2x = zero pad (right, bottom) + stride-2 conv, :37."""
    h = _conv(x, vsd, "encoder.conv_in", 1)
    for lvl in range(num_resolutions):
        for ib in range(num_res_blocks):
            h = resnet_block(h, vsd, f"encoder.down.{lvl}.block.{ib}.")
            if f"encoder.down.{lvl}.attn.{ib}.norm.weight" in vsd:
                h = attn_block(h, vsd, f"encoder.down.{lvl}.attn.{ib}.")
        if lvl != num_resolutions - 1:
            p = f"encoder.down.{lvl}.downsample.conv"
            h = F.conv2d(F.pad(h, pad=(0, 1, 0, 1), mode="constant", value=0), vsd[p + ".weight"], vsd[p + ".bias"],
                         stride=2, padding=0)
    h = resnet_block(h, vsd, "encoder.mid.block_1.")
    h = attn_block(h, vsd, "encoder.mid.attn_1.")
    h = resnet_block(h, vsd, "encoder.mid.block_2.")
    return _conv(F.silu(_gn(h, vsd, "encoder.norm_out"), inplace=True), vsd, "encoder.conv_out", 1)


def img_to_f(img: Tensor, vsd: Dict[str, Tensor]) -> Tensor:
    """quant_conv(encoder(img)) - models/vqvae.py:74."""
    return _conv(encoder_forward(img, vsd), vsd, "quant_conv", 1)


def img_to_idxBl(img: Tensor, vsd: Dict[str, Tensor], patch_nums: Sequence[int]) -> List[Tensor]:
    """VQVAE.img_to_idxBl - models/vqvae.py:73-75."""
    return f_to_idxBl(img_to_f(img, vsd), patch_nums, vsd)


def f_to_fhat_list(f_BChw: Tensor, patch_nums: Sequence[int], vsd: Dict[str, Tensor], K: int = 4) -> List[Tensor]:
    """VectorQuantizer2.f_to_idxBl_or_fhat(to_fhat=True) - models/quant.py:184-215: f_hat after every scale."""
    B, C, H, W = f_BChw.shape
    f_rest = f_BChw.detach().clone()
    f_hat = torch.zeros_like(f_rest)
    emb = vsd["quantize.embedding.weight"]
    SN = len(patch_nums)
    out = []
    for si, pn in enumerate(patch_nums):
        z_NC = (F.interpolate(f_rest, size=(pn, pn), mode="area") if si != SN - 1 else f_rest).permute(0, 2, 3, 1).reshape(-1, C)
        idx_Bhw = vq_nearest(z_NC, emb).view(B, pn, pn)
        h = F.embedding(idx_Bhw, emb).permute(0, 3, 1, 2)
        h = F.interpolate(h, size=(H, W), mode="bicubic").contiguous() if si != SN - 1 else h.contiguous()
        k = phi_index(si, SN, K)
        h = phi(h, vsd[f"quantize.quant_resi.qresi_ls.{k}.weight"], vsd[f"quantize.quant_resi.qresi_ls.{k}.bias"])
        f_hat.add_(h)
        f_rest.sub_(h)
        out.append(f_hat.clone())
    return out


def img_to_recon_last(img: Tensor, vsd: Dict[str, Tensor], patch_nums: Sequence[int]) -> Tensor:
    """VQVAE.img_to_recon(last_one=True) - models/vqvae.py:80-86: decoder(post_quant_conv(f_hat_last)), NOT clamped."""
    f_hat = f_to_fhat_list(img_to_f(img, vsd), patch_nums, vsd)[-1]
    return decoder_forward(_conv(f_hat, vsd, "post_quant_conv", 1), vsd)


def idxBl_to_fhat_list(ms_idx_Bl: List[Tensor], patch_nums: Sequence[int], vsd: Dict[str, Tensor], K: int = 4) -> List[Tensor]:
    """VQVAE.idxBl_to_img's embedding step + VectorQuantizer2.embed_to_fhat(all_to_max_scale=True) - models/vqvae.py:97-104,
    models/quant.py:156-170: f_hat after every scale from given token ids."""
    emb = vsd["quantize.embedding.weight"]
    B, Cv = ms_idx_Bl[0].shape[0], emb.shape[1]
    SN, HW = len(patch_nums), patch_nums[-1]
    f_hat = torch.zeros(B, Cv, HW, HW)
    out = []
    for si, pn in enumerate(patch_nums):
        h = F.embedding(ms_idx_Bl[si], emb).transpose(1, 2).view(B, Cv, pn, pn)
        if si < SN - 1:
            h = F.interpolate(h, size=(HW, HW), mode="bicubic")
        k = phi_index(si, SN, K)
        f_hat.add_(phi(h, vsd[f"quantize.quant_resi.qresi_ls.{k}.weight"], vsd[f"quantize.quant_resi.qresi_ls.{k}.bias"]))
        out.append(f_hat.clone())
    return out


def idxBl_to_img_last(ms_idx_Bl: List[Tensor], patch_nums: Sequence[int], vsd: Dict[str, Tensor]) -> Tensor:
    """VQVAE.idxBl_to_img(same_shape=True, last_one=True) - models/vqvae.py:97-104, 91-93 (clamped to [-1, 1])."""
    return fhat_to_img(idxBl_to_fhat_list(ms_idx_Bl, patch_nums, vsd)[-1], vsd)


def idxBl_to_var_input(ms_idx_Bl: List[Tensor], patch_nums: Sequence[int], vsd: Dict[str, Tensor], K: int = 4) -> List[Tensor]:
    """VectorQuantizer2.idxBl_to_var_input (= VQVAE.idxBl_to_h) - models/quant.py:217-241: the teacher-forcing inputs of
    VAR training, one (B, pn_next^2, Cvae) tensor per scale transition."""
    emb = vsd["quantize.embedding.weight"]
    B, Cv = ms_idx_Bl[0].shape[0], emb.shape[1]
    SN, HW = len(patch_nums), patch_nums[-1]
    f_hat = torch.zeros(B, Cv, HW, HW)
    out = []
    pn_next = patch_nums[0]
    for si in range(SN - 1):
        h = F.interpolate(F.embedding(ms_idx_Bl[si], emb).transpose_(1, 2).view(B, Cv, pn_next, pn_next), size=(HW, HW),
                          mode="bicubic")
        k = phi_index(si, SN, K)
        f_hat.add_(phi(h, vsd[f"quantize.quant_resi.qresi_ls.{k}.weight"], vsd[f"quantize.quant_resi.qresi_ls.{k}.bias"]))
        pn_next = patch_nums[si + 1]
        out.append(F.interpolate(f_hat, size=(pn_next, pn_next), mode="area").view(B, Cv, -1).transpose(1, 2))
    return out


def vq_nearest_margin(z_NC: Tensor, emb: Tensor) -> Tensor:
    """Test helper: gap between the two smallest code distances of every latent vector (ambiguous argmins)."""
    d = torch.sum(z_NC.square(), dim=1, keepdim=True) + torch.sum(emb.square(), dim=1, keepdim=False)
    d.addmm_(z_NC, emb.T, alpha=-2, beta=1)
    two = d.topk(2, dim=1, largest=False)[0]
    return two[:, 1] - two[:, 0]


# ----------------------------------------------------------------------------------- the sampler (a-1)
@torch.no_grad()
def autoregressive_infer_cfg(
    sd: Dict[str, Tensor], vsd: Dict[str, Tensor], patch_nums: Sequence[int], depth: int,
    B: int, label_B: Tensor, cond_type: Tensor, cfg: float, top_k: int, top_p: float,
    noise: Callable[[int, int, int], Tensor], decode: bool = True, trace: Optional[dict] = None,
    forced_idx: Optional[List[Tensor]] = None, embed_dim: int = 0, num_heads: int = 0, more_smooth: bool = False,
) -> Dict[str, object]:
    """ControlVAR.autoregressive_infer_cfg, released branch - models/control_var.py:373-409, 486-565.
    ``more_smooth`` (:511-515, visualisation only): ``noise`` is then called twice per scale - multinomial, then Gumbel.

    ``noise(si, n_rows, V)`` returns the Exp(1) tensor that torch.multinomial would have drawn at scale si.
    ``forced_idx`` (teacher forcing, test helper) replaces the sampled tokens after sampling so that later scales
    are conditioned on a given trajectory while the freely sampled tokens are still reported.
    """
    C = embed_dim or 64 * depth                                               # models/__init__.py:37-40
    H = num_heads or depth
    cos_attn = depth == 30                                                    # control_var.py:35
    scale = 1.0 if cos_attn else 1 / math.sqrt(C // H) / 4                      # basic_var.py:66-71
    SN = len(patch_nums)
    Cvae = vsd["quantize.embedding.weight"].shape[1]
    V = vsd["quantize.embedding.weight"].shape[0]
    num_classes = sd["class_emb.weight"].shape[0] - 1
    first_l = 2 * patch_nums[0] ** 2

    label_B = label_B.long()
    sos = cond_BD = F.embedding(torch.cat((label_B, torch.full_like(label_B, num_classes)), dim=0),
                                sd["class_emb.weight"])                       # :381
    lvl_pos = F.embedding(sd["lvl_1L"], sd["lvl_embed.weight"]) + sd["pos_1LC"]   # :383
    uncond_type = torch.full((B,), 4).long()                                  # :399
    ct = torch.concat([cond_type.long(), uncond_type], dim=0)                 # :400
    sos = sos.unsqueeze(1)
    cond_token = F.embedding(ct, sd["cond_embed.weight"]).unsqueeze(1)        # :402
    next_token_map = torch.concat([cond_token, sos], dim=1)                   # :405 (mask_first is always True)
    next_token_map = next_token_map + sd["pos_start"].expand(2 * B, first_l, -1) + lvl_pos[:, :first_l]   # :409

    caches = [dict() for _ in range(depth)]
    cur_L = 0
    HW = patch_nums[-1]
    f_hat = sos.new_zeros(B, Cvae, HW * 2, HW)                                # :489
    emb = vsd["quantize.embedding.weight"]
    idx_all: List[Tensor] = []
    for si, pn in enumerate(patch_nums):                                      # :490
        ratio = si / (SN - 1)
        cur_L += pn * pn * 2
        x = next_token_map
        for bi in range(depth):                                               # :496-498
            x = adaln_block(x, cond_BD, sd, f"blocks.{bi}.", H, caches[bi], cos_attn, scale)
        logits_BlV = get_logits(x, cond_BD, sd)                               # :499
        t = cfg * ratio
        logits_BlV = cfg_combine(logits_BlV, B, t)                            # :502
        logits_BlV = logits_BlV[:, :, :V]                                     # :504
        if trace is not None:
            trace.setdefault("x_last", []).append(x.clone())
            trace.setdefault("logits_cfg", []).append(logits_BlV.clone())
        q = noise(si, B * logits_BlV.shape[1], V)
        idx_Bl = sample_with_top_k_top_p_(logits_BlV, top_k, top_p, q)        # :505
        if trace is not None:
            trace.setdefault("logits_masked", []).append(logits_BlV.clone())
            trace.setdefault("q", []).append(q)
        idx_all.append(idx_Bl.clone())
        if forced_idx is not None:
            idx_Bl = forced_idx[si].clone()
        if not more_smooth:
            h_BChw = F.embedding(idx_Bl, emb)                                 # :512
        else:
            h_BChw = gumbel_soft_embedding(logits_BlV, ratio, noise(si, B * logits_BlV.shape[1], V), emb)   # :514-515
            if trace is not None:
                trace.setdefault("h_soft", []).append(h_BChw.clone())
        h_BChw = h_BChw.transpose_(1, 2)                                      # :522
        h1 = h_BChw[:, :, :pn * pn].reshape(B, Cvae, pn, pn)
        h2 = h_BChw[:, :, -pn * pn:].reshape(B, Cvae, pn, pn)
        f_hat_1 = f_hat[:, :, :HW, :]
        f_hat_2 = f_hat[:, :, HW:, :]
        f_hat_1, ntm1 = get_next_autoregressive_input(si, patch_nums, f_hat_1, h1, vsd)   # :527
        f_hat_2, ntm2 = get_next_autoregressive_input(si, patch_nums, f_hat_2, h2, vsd)   # :528
        f_hat = torch.concat((f_hat_1, f_hat_2), dim=2)                       # :529
        next_token_map = torch.concat((ntm1, ntm2), dim=2)                    # :530
        if si != SN - 1:                                                      # :534
            n1, n2 = next_token_map[:, :, :pn, :], next_token_map[:, :, pn:, :]     # :539 (uses *current* pn)
            n1 = n1.reshape(B, Cvae, -1).transpose(1, 2)
            n2 = n2.reshape(B, Cvae, -1).transpose(1, 2)
            next_token_map = torch.concat((n1, n2), dim=1)                    # :554
            if trace is not None:
                trace.setdefault("next_map_Cvae", []).append(next_token_map.clone())
            next_token_map = F.linear(next_token_map, sd["word_embed.weight"], sd["word_embed.bias"])   # :555
            next_token_map = next_token_map + lvl_pos[:, cur_L:cur_L + patch_nums[si + 1] ** 2 * 2]    # :557
            next_token_map = next_token_map.repeat(2, 1, 1)                   # :560
    out: Dict[str, object] = {"idx": idx_all, "f_hat": f_hat}
    if decode:
        img1 = fhat_to_img(f_hat_1, vsd).add_(1).mul_(0.5)                    # :563
        img2 = fhat_to_img(f_hat_2, vsd).add_(1).mul_(0.5)                    # :564
        out["img"] = torch.concat([img1, img2], dim=2)                        # :565
    return out


@torch.no_grad()
def forward_teacher_forced(sd: Dict[str, Tensor], patch_nums: Sequence[int], depth: int, label_B: Tensor,
                           x_BLCv_wo_first_l: Tensor, cond_type: Tensor, embed_dim: int = 0, num_heads: int = 0,
                           mask_first: bool = True) -> Tensor:
    """ControlVAR.forward, released branch (multi_cond, mask_factor 2, mask_first=True, prog_si=-1) -
    models/control_var.py:566-651: one full-sequence pass under the block-causal attn_bias_for_masking.
    The label / condition-type dropout of :577 and :584 (torch.rand < cond_drop_rate, active even in eval) is the
    caller's: pass the ids as they should enter the model (the goldens are made with cond_drop_rate = 0)."""
    C = embed_dim or 64 * depth
    H = num_heads or depth
    cos_attn = depth == 30
    scale = 1.0 if cos_attn else 1 / math.sqrt(C // H) / 4
    B = x_BLCv_wo_first_l.shape[0]
    first_l = 2 * patch_nums[0] ** 2
    sos = cond_BD = F.embedding(label_B.long(), sd["class_emb.weight"])                       # :578
    sos = sos.unsqueeze(1).expand(B, 1, -1)                                                   # :581
    cond_token = F.embedding(cond_type.long(), sd["cond_embed.weight"]).unsqueeze(1).expand(B, 1, -1)   # :585-586
    sos = torch.concat([cond_token, sos], dim=1) if mask_first else torch.concat([sos, cond_token], dim=1)   # :587
    sos = sos + sd["pos_start"].expand(B, first_l, -1)                                        # :588
    x_BLC = torch.cat((sos, F.linear(x_BLCv_wo_first_l.float(), sd["word_embed.weight"], sd["word_embed.bias"])), dim=1)  # :616
    x_BLC += F.embedding(sd["lvl_1L"].expand(B, -1), sd["lvl_embed.weight"]) + sd["pos_1LC"]   # :618
    attn_bias = sd["attn_bias_for_masking"]                                                   # :622
    for bi in range(depth):                                                                   # :634-635
        x_BLC = adaln_block(x_BLC, cond_BD, sd, f"blocks.{bi}.", H, None, cos_attn, scale, attn_bias)
    return get_logits(x_BLC.float(), cond_BD, sd)                                             # :636


def cfg_combine4(logits_4BlV: Tensor, B: int, t1: float, t2: float, t3: float) -> Tensor:
    """The four-way guidance mix of conditional_infer_cfg - models/control_var.py:295-298."""
    return (1 + t1) * logits_4BlV[:B] \
        + (t2 - t1) * logits_4BlV[B:2 * B] \
        + (t3 - t2) * logits_4BlV[2 * B:3 * B] \
        - t3 * logits_4BlV[-B:]


@torch.no_grad()
def conditional_infer_cfg(
    sd: Dict[str, Tensor], vsd: Dict[str, Tensor], patch_nums: Sequence[int], depth: int,
    B: int, label_B: Tensor, cond_type: Tensor, cfg: Sequence[float], top_k: int, top_p: float,
    noise: Callable[[int, int, int], Tensor], c_mask: Optional[List[Tensor]] = None,
    c_img: Optional[List[Tensor]] = None, decode: bool = True, trace: Optional[dict] = None,
    embed_dim: int = 0, num_heads: int = 0, more_smooth: bool = False,
) -> Dict[str, object]:
    """ControlVAR.conditional_infer_cfg - models/control_var.py:223-354 (pixel-level control: the condition map's
    and / or the image's tokens are teacher-forced into three of four guidance replicas).

    Rows are [class+type | type only | nothing | nothing] x B (label_B cat at :256, cond_type cat at :263), the
    mixed logits are repeated 4x and every replica row draws its own sample (:306-307; ``noise(si, 4*B*l, V)``).
    """
    C = embed_dim or 64 * depth
    H = num_heads or depth
    cos_attn = depth == 30
    scale = 1.0 if cos_attn else 1 / math.sqrt(C // H) / 4
    SN = len(patch_nums)
    emb = vsd["quantize.embedding.weight"]
    V, Cvae = emb.shape
    num_classes = sd["class_emb.weight"].shape[0] - 1
    first_l = 2 * patch_nums[0] ** 2
    HW = patch_nums[-1]

    lvl_pos = F.embedding(sd["lvl_1L"], sd["lvl_embed.weight"]) + sd["pos_1LC"]             # :245
    label_B = label_B.long()
    empty_cls = torch.full_like(label_B, num_classes)                                       # :252
    label_4B = torch.cat((label_B, empty_cls, empty_cls, empty_cls), dim=0)                  # :256
    sos = cond_BD = F.embedding(label_4B, sd["class_emb.weight"])                           # :257
    empty_ct = torch.full((B,), 4).long()                                                   # :259
    ct = torch.concat([cond_type.long(), cond_type.long(), empty_ct, empty_ct], dim=0)      # :263
    sos = sos.unsqueeze(1)
    cond_token = F.embedding(ct, sd["cond_embed.weight"]).unsqueeze(1)                      # :265
    next_token_map = torch.concat([cond_token, sos], dim=1)                                 # :266
    rep = label_4B.shape[0] // B                                                            # :268 (always 4)
    next_token_map = next_token_map + sd["pos_start"].expand(rep * B, first_l, -1) + lvl_pos[:, :first_l]   # :269

    caches = [dict() for _ in range(depth)]
    cur_L = 0
    f_hat = sos.new_zeros(rep * B, Cvae, HW * 2, HW)                                        # :275
    idx_all: List[Tensor] = []
    for si, pn in enumerate(patch_nums):                                                    # :276
        ratio = si / (SN - 1)
        cur_L += pn * pn * 2
        x = next_token_map
        for bi in range(depth):                                                             # :282-284 (bias slice is all zeros)
            x = adaln_block(x, cond_BD, sd, f"blocks.{bi}.", H, caches[bi], cos_attn, scale)
        logits = get_logits(x, cond_BD, sd)                                                 # :285
        t1, t2, t3 = cfg[0] * ratio, cfg[1] * ratio, cfg[2] * ratio                         # :288
        logits = cfg_combine4(logits, B, t1, t2, t3)                                        # :295-298
        if trace is not None:
            trace.setdefault("logits_cfg", []).append(logits.clone())
        logits = logits.repeat(rep, 1, 1)                                                   # :306
        q = noise(si, rep * B * logits.shape[1], V)
        idx_Bl = sample_with_top_k_top_p_(logits, top_k, top_p, q)                          # :307
        if trace is not None:
            trace.setdefault("logits_masked", []).append(logits.clone())
            trace.setdefault("q", []).append(q)
            trace.setdefault("idx_sampled", []).append(idx_Bl.clone())
        if c_mask is not None:                                                              # :309-313
            for g in range(3):
                idx_Bl[g * B:(g + 1) * B, :pn * pn] = c_mask[si]
        if c_img is not None:                                                               # :317-321
            for g in range(3):
                idx_Bl[g * B:(g + 1) * B, pn * pn:] = c_img[si]
        idx_all.append(idx_Bl.clone())
        if not more_smooth:
            h_BChw = F.embedding(idx_Bl, emb).transpose_(1, 2)                              # :327, :334
        else:   # :329-331: the mixture comes from the (repeated) masked logits; the forced tokens play no part in it
            h_BChw = gumbel_soft_embedding(logits, ratio, noise(si, logits.shape[0] * logits.shape[1], V), emb).transpose_(1, 2)
        h1 = h_BChw[:, :, :pn * pn].reshape(rep * B, Cvae, pn, pn)
        h2 = h_BChw[:, :, -pn * pn:].reshape(rep * B, Cvae, pn, pn)
        f_hat_1 = f_hat[:, :, :HW, :]
        f_hat_2 = f_hat[:, :, HW:, :]
        f_hat_1, ntm1 = get_next_autoregressive_input(si, patch_nums, f_hat_1, h1, vsd)     # :339
        f_hat_2, ntm2 = get_next_autoregressive_input(si, patch_nums, f_hat_2, h2, vsd)     # :340
        f_hat = torch.concat((f_hat_1, f_hat_2), dim=2)                                     # :341
        ntm1 = ntm1.view(rep * B, Cvae, -1).transpose(1, 2)                                 # :342
        ntm2 = ntm2.view(rep * B, Cvae, -1).transpose(1, 2)
        next_token_map = torch.concat((ntm1, ntm2), dim=1)                                  # :344
        next_token_map = F.linear(next_token_map, sd["word_embed.weight"], sd["word_embed.bias"])   # :345
        if si != SN - 1:
            next_token_map = next_token_map + lvl_pos[:, cur_L:cur_L + patch_nums[si + 1] ** 2 * 2]    # :347
    out: Dict[str, object] = {"idx": idx_all, "f_hat": f_hat}
    if decode:
        img1 = fhat_to_img(f_hat_1[:B], vsd).add_(1).mul_(0.5)                              # :349-352
        img2 = fhat_to_img(f_hat_2[:B], vsd).add_(1).mul_(0.5)
        out["img"] = torch.concat([img1, img2], dim=2)                                      # :354
    return out


def gumbel_soft_embedding(logits_masked: Tensor, ratio: float, e: Tensor, emb: Tensor) -> Tensor:
    """The ``more_smooth`` replacement for ``embedding(idx_Bl)`` - models/control_var.py:513-515 (and :329-331),
    models/helpers.py:22-36 with ``hard=False``: a Gumbel-softmax mixture of code vectors.

    ``logits_masked`` are the guidance-mixed logits AFTER sample_with_top_k_top_p_ has masked them in place
    (helpers.py:8-15: removed entries are -inf, so their weight is exactly 0); ``e`` is the Exp(1) tensor that
    ``torch.empty_like(logits).exponential_(generator=rng)`` draws right after torch.multinomial's (helpers.py:26)."""
    gum_t = max(0.27 * (1 - ratio * 0.95), 0.005)                             # control_var.py:514
    gumbels = -e.view_as(logits_masked).log()                                 # helpers.py:26
    gumbels = (logits_masked.mul(1 + ratio) + gumbels) / gum_t                # helpers.py:27, control_var.py:515
    return gumbels.softmax(-1) @ emb.unsqueeze(0)                             # helpers.py:28, control_var.py:515


def cpu_generator_noise(seed: int) -> Callable[[int, int, int], Tensor]:
    """The Exp(1) stream the reference's CPU ``self.rng`` produces inside torch.multinomial after
    ``rng.manual_seed(seed)`` (control_var.py:373-374, helpers.py:19)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)

    def draw(si: int, n_rows: int, V: int) -> Tensor:
        return torch.empty(n_rows, V, dtype=torch.float32).exponential_(1, generator=g)

    return draw
