"""Host mirror of the reference ``ControlVAR`` for the next-scale sampling hot path.

Same constructor arguments, ``state_dict`` keys and ``autoregressive_infer_cfg`` signature / return value as
/root/reference/models/control_var.py:23-67, 356-565, so it drops in for inference.
The Python here only sequences kernel launches of libcvar_sm100.so and owns the device memory:

  prologue                         cvar_lvl_pos, cvar_prologue_rows             control_var.py:381-409
  per block, per scale             cvar_ln_modulate, cvar_qkv_project, cvar_attn_kvcache, cvar_gemm x3
                                                                                 basic_var.py:203-210, 89-119, 43-51
  head + CFG + sampling            cvar_ln_modulate, cvar_gemm, cvar_cfg_sample  control_var.py:499-505, helpers.py:6-19
  VQ step                          cvar_vq_step                                  control_var.py:512-560, quant.py:243-270
  decode (both streams, one pass)  VQVAE._fhat_to_img -> cvar_conv2d / cvar_gn_stats / ...   control_var.py:563-565

``conditional_infer_cfg`` (control_var.py:223-354) runs the same loop on four guidance
replicas with ``cvar_cfg_sample_multi`` (4-way mix, one draw per replica row, teacher-forced tokens).

Differences from the reference that do not change results: ``ada_lin`` (constant across scales) is evaluated once
per call instead of once per block per scale; the KV cache is a pre-allocated arena written in place instead of
``torch.cat`` growth; the control and image halves of ``f_hat`` are decoded in one decoder pass over 2B maps
(``VQVAE._fhat_halves_to_img``) instead of two passes over B.
Branches no released configuration uses (``separator``, ``type_pos``, ``bidirectional``, ``separate_decoding``,
``shared_aln``, ``aln < 0``, ``mask_factor == 1``) raise NotImplementedError.  ``more_smooth`` (the reference's
visualisation-only Gumbel-softmax mixture, control_var.py:511-515) runs eagerly: cvar_cfg_sample_masked + cvar_gumbel_embed.
"""
from __future__ import annotations

import math
import os
import random
import struct
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .config import PathConfig, DEFAULT_PATCH_NUMS
from .vqvae import VQVAE, bicubic_matrix, register_tree
from .weights import attn_bias_for_masking, lvl_1L, var_key_shapes

_BUFFERS = ("lvl_1L", "attn_bias_for_masking", "zero_k_bias")


def _f32(v: float) -> float:
    """A python double rounded to fp32 (what a python scalar becomes when it multiplies an fp32 tensor)."""
    return struct.unpack("f", struct.pack("f", v))[0]


class ControlVAR(nn.Module):
    def __init__(
        self, vae_local: VQVAE,
        num_classes=1000, norm_eps=1e-6, aln=1, aln_gamma_init=1e-3, shared_aln=False, cond_drop_rate=0.1,
        depth=16, embed_dim=1024, num_heads=16, mlp_ratio=4., drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
        layer_scale=-1., tau=4, cos_attn=False,
        patch_nums=DEFAULT_PATCH_NUMS,
        flash_if_available=True, fused_if_available=True, mask_factor=2, bidirectional=False,
        separate_decoding=False, separator=False, type_pos=False, indep=True, multi_cond=False,
    ):
        super().__init__()
        if shared_aln or aln < 0:
            raise NotImplementedError("only AdaLNSABlock with per-block ada_lin (aln >= 0, shared_aln=False) is implemented")
        if separator or type_pos or bidirectional or separate_decoding:
            raise NotImplementedError("separator / type_pos / bidirectional / separate_decoding are not used by any "
                                      "released configuration and are not implemented")
        if mask_factor != 2:
            raise NotImplementedError("mask_type='replace' (mask_factor=1) ends in a NameError in the reference sampler "
                                      "(control_var.py:563); only 'interleave_append' is implemented")
        if embed_dim % num_heads != 0 or embed_dim // num_heads != 64:
            raise NotImplementedError("the attention kernels are specialised for head_dim = 64")
        if mlp_ratio != 4.0 or tau != 4:
            raise NotImplementedError("mlp_ratio must be 4 and tau 4 (the released models)")
        self.cfg = PathConfig(depth=depth, patch_nums=tuple(patch_nums), num_classes=num_classes,
                              vocab_size=vae_local.vocab_size, Cvae=vae_local.Cvae, norm_eps=norm_eps,
                              multi_cond=multi_cond, embed_dim=embed_dim if embed_dim != 64 * depth else 0,
                              heads=num_heads if num_heads != depth else 0)
        cfg = self.cfg
        self.Cvae, self.V = vae_local.Cvae, vae_local.vocab_size
        self.depth, self.C, self.D, self.num_heads = depth, embed_dim, embed_dim, num_heads
        self.cos_attn = cfg.cos_attn                       # control_var.py:35: forced by depth == 30
        self.multi_cond, self.indep, self.mask_factor = multi_cond, indep, mask_factor
        self.cond_drop_rate, self.prog_si = cond_drop_rate, -1
        self.patch_nums = tuple(patch_nums)
        self.L, self.first_l = cfg.L, cfg.first_l
        self.num_stages_minus_1 = len(self.patch_nums) - 1
        self.num_classes = num_classes
        self.norm_eps = norm_eps
        self.vae_proxy: Tuple[VQVAE] = (vae_local,)
        self.vae_quant_proxy = (vae_local.quantize,)
        for key, shape in var_key_shapes(cfg).items():
            last = key.split(".")[-1]
            if key == "lvl_1L":
                t = lvl_1L(cfg)
            elif key == "attn_bias_for_masking":
                t = attn_bias_for_masking(cfg)
            else:
                t = torch.zeros(shape)
            register_tree(self, key, t, last in _BUFFERS)
        # RNG: the reference binds self.rng to dist.get_device() (control_var.py:68).  'cuda' reproduces the
        # reference's stream on a GPU host; 'cpu' reproduces the stream of the reference run on CPU (used by the
        # parity tests against the CPU oracle): the Exp(1) noise is then drawn on the host and copied over.
        self.rng_device = "cuda"
        # engine 4 only: q / K / V^T as FP16 pairs end to end (cvar_qkv_project16 + cvar_attn_kvcache16) instead of
        # fp32 q + TF32-split cache + 3xTF32 attention.  Half the KV arena, twice the attention MMA rate.
        self.kv16 = True
        self._rng: Optional[torch.Generator] = None
        self._ws: Dict[Tuple, torch.Tensor] = {}
        # CUDA graphs: the launch sequence of one sampling call (~2100 launches at d24) is
        # static per (batch, guidance schedule, top-k / top-p, engine), so the second call with the same key is captured
        # and later calls replay it: one graph launch instead of ~2100 ctypes launches from a Python loop, which matters
        # where kernels are shorter than a launch (the five small scales, small batches, 8 ranks sharing one host).
        # Inputs of a replay: labels / condition types / forced tokens (copied into static buffers) and the Exp(1) noise,
        # drawn from the caller's generator BEFORE the replay in the same order as the eager path (bit-identical tokens).
        # forward(): one masked full-sequence pass (True, engine 4) or the scales one after the other against the growing
        # KV cache (False; also what the other engines do) - the same numbers (tests compare them)
        self.forward_single_pass = True
        self.use_graphs = os.environ.get("CVAR_GRAPHS", "1") != "0"
        self._graphs: Dict[Tuple, dict] = {}
        self._ws_gen = 0                                  # bumped whenever a workspace is (re)allocated: graphs hold pointers
        self._consts: Dict[str, object] = {}
        self.last_idx: List[torch.Tensor] = []           # tokens of the last call, per scale (diagnostics / tests)
        # Parity instruments (tests only; SURVEY.md section 7.2 "margin-aware + teacher-forced"):
        #   debug_forced_idx     per-scale (B, l) tokens that REPLACE the sampled ones after sampling, so that the
        #                        trajectory is pinned to the oracle's while last_idx still reports what was sampled;
        #   debug_capture_logits keep a copy of the raw (2B, l, V) logits of every scale in last_logits.
        #   debug_noise_fn       callable (scale index, rows, V) -> (rows, V) Exp(1) tensor replacing the generator draw
        #                        (lets a test feed a sample the very same noise in differently shaped batches).
        self.debug_forced_idx: Optional[List[torch.Tensor]] = None
        self.debug_capture_logits = False
        self.debug_noise_fn = None
        self.last_logits: List[torch.Tensor] = []
        self.eval()

    # ------------------------------------------------------------------------------------------ plumbing
    def _apply(self, fn, recurse=True):
        self._graphs.clear()
        self._ws.clear()
        self._consts.clear()
        return super()._apply(fn, recurse)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        self._graphs.clear()         # derived operand-form weights are rebuilt: captured pointers go stale
        self._consts.clear()
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    @property
    def device(self) -> torch.device:
        return self.pos_1LC.device

    def _buf(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        key = (name, dtype)
        n = 1
        for s in shape:
            n *= s
        t = self._ws.get(key)
        if t is None or t.numel() < n or t.device != self.device:
            t = torch.empty(n, device=self.device, dtype=dtype)
            self._ws[key] = t
            self._ws_gen += 1
        return t[:n].view(shape)

    def _pair(self, name: str, shape) -> "ops.F16Pair":
        return ops.F16Pair(self._buf(name + ".hi", shape, torch.float16), self._buf(name + ".lo", shape, torch.float16))

    def _kv_caches(self, depth, R, H, T):
        """Per-block KV arenas (cached across calls; zero-filled once - every call rewrites the keys it reads)."""
        key = ("kv", depth, R, H, T)
        c = self._ws.get(key)
        if c is None or c[0].k_hi.device != self.device:
            for k in [k for k in self._ws if isinstance(k, tuple) and k and k[0] in ("kv", "kv16")]:
                del self._ws[k]
            n = ops.KVCache.numel(R, H, T)
            arena = torch.zeros(depth * n, dtype=torch.float32, device=self.device)
            c = [ops.KVCache(R, H, T, self.device, arena[i * n:(i + 1) * n]) for i in range(depth)]
            self._ws[key] = c
            self._ws_gen += 1
        return c

    def _kv_caches16(self, depth, R, H, T):
        """The same arena as FP16 pairs (engine 4 with kv16): half the bytes of the TF32 split."""
        key = ("kv16", depth, R, H, T)
        c = self._ws.get(key)
        if c is None or c[0].k_hi.device != self.device:
            for k in [k for k in self._ws if isinstance(k, tuple) and k and k[0] in ("kv", "kv16")]:
                del self._ws[k]
            n = ops.KVCache16.numel(R, H, T)
            arena = torch.zeros(depth * n, dtype=torch.float16, device=self.device)
            c = [ops.KVCache16(R, H, T, self.device, arena[i * n:(i + 1) * n]) for i in range(depth)]
            self._ws[key] = c
            self._ws_gen += 1
        return c

    def release_workspace(self):
        """Free the KV arena and activation scratch (they are cached across calls)."""
        self._graphs.clear()
        self._ws.clear()
        self._ws_gen += 1
        self.vae_proxy[0].release_workspace()

    def _generator(self, seed: Optional[int]) -> Optional[torch.Generator]:
        if seed is None:
            return None
        dev = "cpu" if self.rng_device == "cpu" else self.device
        if self._rng is None or self._rng.device != torch.device(dev):
            self._rng = torch.Generator(device=dev)
        self._rng.manual_seed(seed)
        return self._rng

    def _noise(self, rows: int, V: int, rng: Optional[torch.Generator]) -> torch.Tensor:
        """The Exp(1) tensor torch.multinomial(num_samples=1) draws internally (helpers.py:19)."""
        if self.rng_device == "cpu":
            return torch.empty(rows, V, dtype=torch.float32).exponential_(1, generator=rng).to(self.device, non_blocking=True)
        return torch.empty(rows, V, dtype=torch.float32, device=self.device).exponential_(1, generator=rng)

    def _constants(self):
        c = self._consts
        f16 = ops.get_gemm_engine() == ops.ENGINE_TC_F16X3
        if c and c.get("f16") != f16:        # the engine changed since the weights were put in operand form
            c.clear()
        if not c:
            c["f16"] = f16
            dev, cfg = self.device, self.cfg
            hw = self.patch_nums[-1]
            c["U"] = {pn: bicubic_matrix(pn, hw).to(dev) for pn in set(self.patch_nums) if pn != hw}
            c["lvl_pos"] = ops.lvl_pos(self.get_parameter("lvl_embed.weight"), self.lvl_1L, self.pos_1LC,
                                       torch.empty(self.L, self.C, device=dev))
            P = self.get_parameter
            blocks = []
            for i in range(self.depth):
                p = f"blocks.{i}."
                # operand form of the tensor-core engine, made once per weight: TF32 hi/lo split, or FP16 pairs (engine 4)
                SW = (lambda w: ops.SplitWeight(w, f16=True)) if f16 else ops.SplitWeight
                blocks.append(dict(
                    qkv_w=SW(P(p + "attn.mat_qkv.weight")), q_bias=P(p + "attn.q_bias"), v_bias=P(p + "attn.v_bias"),
                    k_bias=self.get_buffer(p + "attn.zero_k_bias"),
                    proj_w=SW(P(p + "attn.proj.weight")), proj_b=P(p + "attn.proj.bias"),
                    fc1_w=SW(P(p + "ffn.fc1.weight")), fc1_b=P(p + "ffn.fc1.bias"),
                    fc2_w=SW(P(p + "ffn.fc2.weight")), fc2_b=P(p + "ffn.fc2.bias"),
                    ada_w=SW(P(p + "ada_lin.1.weight")), ada_b=P(p + "ada_lin.1.bias"),
                    scale_mul=P(p + "attn.scale_mul_1H11").reshape(-1).contiguous() if self.cos_attn else None,
                ))
            c["blocks"] = blocks
            c["head_ada_w"] = ops.SplitWeight(P("head_nm.ada_lin.1.weight"), f16=True) if f16 else P("head_nm.ada_lin.1.weight")
            c["head_ada_b"] = P("head_nm.ada_lin.1.bias")
            c["head_w"] = ops.SplitWeight(P("head.weight"), f16=f16)
            vq = self.vae_proxy[0]
            c["phi"] = [(vq.get_parameter(f"quantize.quant_resi.qresi_ls.{k}.weight"),
                         vq.get_parameter(f"quantize.quant_resi.qresi_ls.{k}.bias"))
                        for k in range(self.cfg.share_quant_resi)]
            c["codebook"] = vq.get_parameter("quantize.embedding.weight")
        return c

    # -------------------------------------------------------------------------------- argument validation
    # The kernels index embedding tables with these ids (class_emb[label], cond_embed[type], codebook[token]); the
    # reference fails with an embedding device-assert / a shape error on bad input, a raw kernel would read out of bounds.
    def _check_ids(self, label_B: torch.Tensor, cond_type: torch.Tensor, B: int) -> None:
        if tuple(label_B.shape) != (B,) or tuple(cond_type.shape) != (B,):
            raise ValueError(f"label_B / cond_type must have shape ({B},), got {tuple(label_B.shape)} / {tuple(cond_type.shape)}")
        bad = ((label_B < 0) | (label_B > self.num_classes)).any() | ((cond_type < 0) | (cond_type > 4)).any()
        if bool(bad):
            raise ValueError(f"label_B must be in [0, {self.num_classes}] and cond_type in [0, 4] "
                             f"(ids {self.num_classes} / 4 are the unconditional entries)")

    def _check_forced(self, lst, B: int, what: str):
        if lst is None:
            return None
        if len(lst) != len(self.patch_nums):
            raise ValueError(f"{what}: expected {len(self.patch_nums)} token maps (one per scale), got {len(lst)}")
        out, bad = [], None
        for si, (pn, t) in enumerate(zip(self.patch_nums, lst)):
            if not torch.is_tensor(t) or tuple(t.shape) != (B, pn * pn):
                raise ValueError(f"{what}[{si}] must be a ({B}, {pn * pn}) token tensor (VQVAE.img_to_idxBl with the model's "
                                 f"patch_nums), got {tuple(t.shape) if torch.is_tensor(t) else type(t)}")
            t = t.to(device=self.device, dtype=torch.int64).contiguous()
            b = ((t < 0) | (t >= self.V)).any()
            bad = b if bad is None else (bad | b)
            out.append(t)
        if bool(bad):
            raise ValueError(f"{what}: token ids must be in [0, {self.V})")
        return out

    # ------------------------------------------------------------------------------------------- sampler
    @torch.no_grad()
    def autoregressive_infer_cfg(
        self, B: int, label_B: Optional[Union[int, torch.LongTensor]],
        g_seed: Optional[int] = None, cfg=1.5, top_k=0, top_p=0.0,
        more_smooth=False, cond_type=None,
    ) -> torch.Tensor:   # (B, 3, 2*H, W) in [0, 1]: control map on top, image below
        """Drop-in for ControlVAR.autoregressive_infer_cfg (control_var.py:356-565), released branch."""
        if not self.pos_1LC.is_cuda:
            raise RuntimeError("controlvar_b200.ControlVAR runs on CUDA only (no CPU fallback); call .cuda() first")
        dev = self.device
        rng = self._generator(g_seed)
        rng_dev = "cpu" if self.rng_device == "cpu" else dev

        # ---- host-side argument handling, as in control_var.py:376-403
        if label_B is None:
            sel = torch.full((1, self.num_classes), 1 / self.num_classes, dtype=torch.float32, device=rng_dev)
            label_B = torch.multinomial(sel, num_samples=B, replacement=True, generator=rng).reshape(B)
        elif isinstance(label_B, int):
            label_B = torch.full((B,), self.num_classes if label_B < 0 else label_B)
        label_B = label_B.to(device=dev, dtype=torch.long).contiguous()
        if self.multi_cond:
            if cond_type is None:
                if B == 4:
                    cond_type = torch.tensor([0, 1, 2, 3])
                else:
                    cond_idx = torch.full((1, 4), 1 / 4, dtype=torch.float32, device=rng_dev)
                    cond_type = torch.multinomial(cond_idx, num_samples=B, replacement=True, generator=rng).reshape(B)
            elif isinstance(cond_type, int):
                assert cond_type <= 3 and cond_type > 0
                cond_type = torch.full((B,), cond_type)
            cond_type = cond_type.to(device=dev, dtype=torch.long).contiguous()
            random.random()      # control_var.py:403 draws from Python's global RNG even when bidirectional=False
        else:
            cond_type = torch.zeros(B, dtype=torch.long, device=dev)
        self._check_ids(label_B, cond_type, B)

        label_R = torch.cat((label_B, torch.full_like(label_B, self.num_classes)))         # control_var.py:381
        cond_R = torch.cat((cond_type, torch.full_like(cond_type, 4)))                     # control_var.py:399-400
        return self._sample(B, label_R, cond_R, groups=2, mix=lambda ratio: (cfg * ratio,), replicas=1,
                            top_k=top_k, top_p=top_p, rng=rng, more_smooth=bool(more_smooth))

    @torch.no_grad()
    def conditional_infer_cfg(
        self, B: int, label_B: Optional[Union[int, torch.LongTensor]],
        g_seed: Optional[int] = None, cfg=(1.5, 1.5, 1.5), top_k=0, top_p=0.0,
        more_smooth=False, cond_type=None, c_mask=None, c_img=None,
    ) -> torch.Tensor:   # (B, 3, 2*H, W) in [0, 1]
        """Drop-in for ControlVAR.conditional_infer_cfg (control_var.py:223-354): pixel-level control.  Four guidance
        replicas [class + type | type | none | none] of every sample run side by side (4B rows); c_mask / c_img
        (List[(B, pn*pn)] from VQVAE.img_to_idxBl) teacher-force the control / image tokens of the first three; the
        image is decoded from the first replica."""
        if not self.multi_cond:
            raise NotImplementedError("conditional_infer_cfg needs multi_cond=True (it reads cond_embed, control_var.py:265)")
        if not self.pos_1LC.is_cuda:
            raise RuntimeError("controlvar_b200.ControlVAR runs on CUDA only (no CPU fallback); call .cuda() first")
        dev = self.device
        rng = self._generator(g_seed)
        rng_dev = "cpu" if self.rng_device == "cpu" else dev
        if label_B is None:
            sel = torch.full((1, self.num_classes), 1 / self.num_classes, dtype=torch.float32, device=rng_dev)
            label_B = torch.multinomial(sel, num_samples=B, replacement=True, generator=rng).reshape(B)
        elif isinstance(label_B, int):
            label_B = torch.full((B,), self.num_classes if label_B < 0 else label_B)
        if cond_type is None:
            raise TypeError("conditional_infer_cfg: cond_type must be given (the reference concatenates it, control_var.py:263)")
        if isinstance(cond_type, int):
            cond_type = torch.full((B,), cond_type)
        label_B = label_B.to(device=dev, dtype=torch.long).contiguous()
        cond_type = cond_type.to(device=dev, dtype=torch.long).contiguous()
        self._check_ids(label_B, cond_type, B)
        cfg = tuple(float(c) for c in cfg)
        assert len(cfg) == 3
        empty_cls = torch.full_like(label_B, self.num_classes)                             # control_var.py:252
        empty_ct = torch.full_like(cond_type, 4)                                           # control_var.py:259
        label_R = torch.cat((label_B, empty_cls, empty_cls, empty_cls))                    # control_var.py:256
        cond_R = torch.cat((cond_type, cond_type, empty_ct, empty_ct))                     # control_var.py:263

        return self._sample(B, label_R, cond_R, groups=4,
                            mix=lambda ratio: (cfg[0] * ratio, cfg[1] * ratio, cfg[2] * ratio), replicas=4,
                            top_k=top_k, top_p=top_p, rng=rng, c_mask=self._check_forced(c_mask, B, "c_mask"),
                            c_img=self._check_forced(c_img, B, "c_img"), more_smooth=bool(more_smooth))

    # ------------------------------------------------------------------------------------ the shared scale loop
    def _sample(self, B: int, label_R: torch.Tensor, cond_R: torch.Tensor, *, groups: int, mix, replicas: int, top_k, top_p,
                rng, c_mask=None, c_img=None, more_smooth: bool = False) -> torch.Tensor:
        """Eager launch sequence, or capture / replay of it as a CUDA graph (see __init__).  Eager whenever a debug hook or
        the per-kernel profiler is active (both put host logic between launches)."""
        SN = len(self.patch_nums)
        ts_all = tuple(tuple(float(t) for t in mix(si / self.num_stages_minus_1)) for si in range(SN))
        Bf, V, lens = B * replicas, self.V, self.cfg.scale_lens
        graphable = (self.use_graphs and self.debug_forced_idx is None and not self.debug_capture_logits
                     and self.debug_noise_fn is None and not ops.profiling() and not more_smooth)
        if not graphable:      # (more_smooth - the reference's visualisation-only mode - draws twice per scale: kept eager)
            return self._sample_body(B, label_R, cond_R, groups, ts_all, replicas, top_k, top_p,
                                     lambda si, rows: self._noise_for_scale(si, rows, rng), c_mask, c_img,
                                     more_smooth=more_smooth)
        vae = self.vae_proxy[0]
        key = (B, groups, replicas, int(top_k), float(top_p), ts_all, c_mask is not None, c_img is not None,
               ops.get_gemm_engine(), ops.get_fast_mode(), self.kv16, self.rng_device, vae.tc_min_hw, vae.ksplit_min_k)
        ent = self._graphs.get(key)
        gens = (self._ws_gen, vae._ws_gen)
        if ent is not None and ent["gens"] != gens:        # a workspace moved since: the captured pointers are stale
            ent = None
            self._graphs.pop(key, None)
        if ent is None:
            # first call with this key: eager (it also allocates every workspace the capture will use)
            img = self._sample_body(B, label_R, cond_R, groups, ts_all, replicas, top_k, top_p,
                                    lambda si, rows: self._noise_for_scale(si, rows, rng), c_mask, c_img)
            self._graphs[key] = dict(graph=None, gens=(self._ws_gen, vae._ws_gen))
            return img
        dev = self.device
        # ---- static inputs
        rows_all = [Bf * l for l in lens]
        noise = self._buf("g_noise", (sum(rows_all), V))
        lab_s, cond_s = self._buf("g_label", (label_R.numel(),), torch.int64), self._buf("g_cond", (cond_R.numel(),), torch.int64)
        lab_s.copy_(label_R)
        cond_s.copy_(cond_R)
        forced = {}
        for name, lst in (("c_mask", c_mask), ("c_img", c_img)):
            if lst is not None:
                bufs = [self._buf(f"g_{name}{si}", tuple(t.shape), torch.int64) for si, t in enumerate(lst)]
                for b_, t in zip(bufs, lst):
                    b_.copy_(t)
                forced[name] = bufs
        # ---- the Exp(1) noise of all scales, in the order the eager path draws it
        off = 0
        for si, rows in enumerate(rows_all):
            self._noise_for_scale(si, rows, rng, out=noise[off:off + rows])
            off += rows

        def noise_view(si, rows):
            o = sum(rows_all[:si])
            return noise[o:o + rows]

        if ent["graph"] is None:
            n0 = ops.launch_count()
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                img_static = self._sample_body(B, lab_s, cond_s, groups, ts_all, replicas, top_k, top_p, noise_view,
                                               forced.get("c_mask"), forced.get("c_img"))
            ent.update(graph=g, img=img_static, last_idx=self.last_idx, f_hat=self.last_f_hat,
                       launches=ops.launch_count() - n0, gens=(self._ws_gen, vae._ws_gen))
        ent["graph"].replay()
        ops.add_launch_count(ent["launches"])              # the replay launched that many of our kernels
        self.last_idx, self.last_f_hat, self.last_logits = ent["last_idx"], ent["f_hat"], []
        return ent["img"].clone()

    def _noise_for_scale(self, si: int, rows: int, rng, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            return self._noise(rows, self.V, rng)
        if self.rng_device == "cpu":
            out.copy_(torch.empty(rows, self.V, dtype=torch.float32).exponential_(1, generator=rng))
        else:
            out.exponential_(1, generator=rng)
        return out

    def _sample_body(self, B: int, label_R: torch.Tensor, cond_R: torch.Tensor, groups: int, ts_all, replicas: int, top_k,
                     top_p, noise_fn, c_mask=None, c_img=None, more_smooth: bool = False) -> torch.Tensor:
        """The 10-scale loop both entry points share.  groups: guidance row groups in the transformer batch (R = groups*B
        rows).  replicas = 1: autoregressive_infer_cfg - one f_hat per sample, the next map is written to both CFG halves.
        replicas = groups = 4: conditional_infer_cfg - every replica row keeps its own samples and f_hat.
        ts_all[si]: the guidance scalars of scale si; noise_fn(si, rows) -> (rows, V) Exp(1) noise."""
        dev = self.device
        vae = self.vae_proxy[0]
        C, V, Cvae = self.C, self.V, self.Cvae
        R, hw = groups * B, self.patch_nums[-1]
        Bf = B * replicas                                   # samples that own an f_hat / a token row
        SN = len(self.patch_nums)
        lens = self.cfg.scale_lens
        lmax = max(lens)
        tr = self._transformer(R)
        cst, lvl_pos, x, logits = tr.cst, tr.lvl_pos, tr.x, tr.logits
        idx = self._buf("idx", (Bf * lmax,), torch.int64)
        f_hat = self._buf("f_hat", (Bf, Cvae, 2 * hw, hw))
        f_hat.zero_()
        tr.prologue(label_R, cond_R)

        self.last_idx = []
        self.last_logits = []
        cur_L = 0
        for si, pn in enumerate(self.patch_nums):
            l = lens[si]
            M = R * l
            L_prev = cur_L
            cur_L += l
            tr.scale(l, L_prev)                               # blocks + head: logits (R*l, V)
            # guidance mix + top-k/top-p + multinomial
            ts = ts_all[si]
            if self.debug_noise_fn is not None:
                q_noise = self.debug_noise_fn(si, Bf * l, V).to(device=dev, dtype=torch.float32).contiguous()
                assert q_noise.shape == (Bf * l, V)
            else:
                q_noise = noise_fn(si, Bf * l)
            if self.debug_capture_logits:
                self.last_logits.append(logits[:M].view(R, l, V).clone())
            if more_smooth:
                # control_var.py:511-515 / 326-331: the next map is built from a Gumbel-softmax MIXTURE of code vectors taken
                # from the logits as sample_with_top_k_top_p_ left them, not from the sampled (or forced) tokens.  Two draws
                # per scale from the same generator: torch.multinomial's, then helpers.py:26's.
                masked = self._buf("masked_logits", (B * lmax, V))
                coef = ((_f32(1 + ts[0]), -_f32(ts[0])) if groups == 2 else
                        (_f32(1 + ts[0]), _f32(ts[1] - ts[0]), _f32(ts[2] - ts[1]), -_f32(ts[2])))
                ops.cfg_sample_masked(logits, q_noise, idx, masked, B, l, V, coef, replicas, top_k, top_p,
                                      forced_first=None if c_mask is None else c_mask[si],
                                      forced_second=None if c_img is None else c_img[si],
                                      forced_replicas=3 if groups == 4 else 0)
                if self.debug_noise_fn is not None:
                    e_noise = self.debug_noise_fn(si, Bf * l, V).to(device=dev, dtype=torch.float32).contiguous()
                else:
                    e_noise = noise_fn(si, Bf * l)
                ratio = si / self.num_stages_minus_1
                h_soft = self._buf("h_soft", (Bf * lmax, Cvae))
                ops.gumbel_embed(masked, e_noise, cst["codebook"], h_soft, B * l, Bf * l, V, Cvae, 1 + ratio,
                                 max(0.27 * (1 - ratio * 0.95), 0.005))
            elif groups == 2:
                ops.cfg_sample(logits, q_noise, idx, B, l, V, ts[0], top_k, top_p)
            else:
                # (1 + t1)*L0 + (t2 - t1)*L1 + (t3 - t2)*L2 - t3*L3: python forms the scalars in double precision and
                # they meet the fp32 tensors as fp32 values (control_var.py:288-298)
                t1, t2, t3 = ts
                coef = (_f32(1 + t1), _f32(t2 - t1), _f32(t3 - t2), -_f32(t3))
                ops.cfg_sample_multi(logits, q_noise, idx, B, l, V, coef, replicas, top_k, top_p,
                                     forced_first=None if c_mask is None else c_mask[si],
                                     forced_second=None if c_img is None else c_img[si], forced_replicas=3)
            self.last_idx.append(idx[:Bf * l].view(Bf, l).clone())
            if self.debug_forced_idx is not None:
                idx[:Bf * l].copy_(self.debug_forced_idx[si].to(device=dev, dtype=torch.int64).reshape(-1))
            # VQ step + next-scale input
            phi_w, phi_b = cst["phi"][self.cfg.phi_index(si)]
            pn_next = self.patch_nums[si + 1] if si != SN - 1 else 0
            if more_smooth:
                # the VQ step gathers embedding[idx]: row r of the mixture table through the identity index = the mixture itself
                vq_idx, vq_table = self._arange_idx(Bf * lmax)[:Bf * l], h_soft
            else:
                vq_idx, vq_table = idx, cst["codebook"]
            ops.vq_step(vq_idx, vq_table, cst["U"].get(pn), phi_w, phi_b,
                        self.get_parameter("word_embed.weight"), self.get_parameter("word_embed.bias"),
                        lvl_pos[cur_L:] if pn_next else None, f_hat, x if pn_next else None,
                        Bf, pn, pn_next, hw, Cvae, C, streams=2, x_replicas=(2 if replicas == 1 else 1))

        # ---- decode both halves (control rows on top, image rows below): control_var.py:563-565 / 349-354
        img = vae._fhat_halves_to_img(f_hat, B, out_mode=1)       # one decoder pass over the 2B maps
        self.last_f_hat = f_hat
        return img

    def _start_rows_class_first(self, label_B: torch.Tensor, cond_type: torch.Tensor, tr: "_Transformer") -> torch.Tensor:
        """forward(mask_first=False), control_var.py:587: the class token BEFORE the condition-type token.  Two rows per sample
        of an option no released configuration uses: formed with torch ops in the reference's order of additions -
        (token + pos_start) + (lvl_embed + pos_1LC), :588 and :618."""
        import torch.nn.functional as F
        sos = F.embedding(label_B, self.get_parameter("class_emb.weight")).unsqueeze(1)
        ct = F.embedding(cond_type, self.get_parameter("cond_embed.weight")).unsqueeze(1)
        first = torch.cat([sos, ct], dim=1) + self.pos_start.expand(label_B.shape[0], self.first_l, -1)
        return first + tr.lvl_pos[:self.first_l]

    def _arange_idx(self, n: int) -> torch.Tensor:
        """0 .. n-1 as int64 on the device (the identity index of the more_smooth VQ step), rebuilt when the workspace moved."""
        t = self._buf("arange_idx", (n,), torch.int64)
        if getattr(self, "_arange_state", None) != (t.data_ptr(), n):
            t.copy_(torch.arange(n, device=t.device))
            self._arange_state = (t.data_ptr(), n)
        return t

    def _transformer(self, R: int, lmax: Optional[int] = None) -> "_Transformer":
        return _Transformer(self, R, lmax)

    # ----------------------------------------------------------------------------------- teacher-forced pass
    @torch.no_grad()
    def forward(self, label_B: torch.LongTensor, x_BLCv_wo_first_l: torch.Tensor, cond_type, mask_first=True) -> torch.Tensor:
        """Drop-in for ControlVAR.forward (control_var.py:566-651), inference only: logits (B, L, V) of the teacher-forced
        token pyramid under the block-causal mask of control_var.py:168 (a query sees the keys of its own and of all
        coarser scales).  Default (engine 4): ONE full-sequence pass like the reference's, the mask applied inside a single
        attention launch per block (cvar_attn_blockcausal16); ``forward_single_pass = False`` (and the other engines) runs
        the scales one after the other against the growing KV cache, which is the same computation (SURVEY.md section 4
        identity; tests/test_oracle_golden.py checks it on the oracle, tests/test_gpu_conditional.py compares the two).
        As in the reference, labels and condition types are dropped with probability cond_drop_rate by torch.rand draws
        even in eval mode (:577, :584): set ``cond_drop_rate = 0`` for deterministic logits."""
        if not self.pos_1LC.is_cuda:
            raise RuntimeError("controlvar_b200.ControlVAR runs on CUDA only (no CPU fallback); call .cuda() first")
        if not self.multi_cond:
            raise NotImplementedError("forward is implemented for the released configuration (multi_cond=True)")
        dev = self.device
        B = x_BLCv_wo_first_l.shape[0]
        C, V, Cvae = self.C, self.V, self.Cvae
        assert x_BLCv_wo_first_l.shape == (B, self.L - self.first_l, Cvae)
        label_B = label_B.to(device=dev, dtype=torch.long)
        cond_type = cond_type.to(device=dev, dtype=torch.long)
        self._check_ids(label_B, cond_type, B)
        label_B = torch.where(torch.rand(B, device=dev) < self.cond_drop_rate, self.num_classes, label_B)      # :577
        cond_type = torch.where(torch.rand(B, device=dev) < self.cond_drop_rate, 4, cond_type)                # :584
        xin = x_BLCv_wo_first_l.to(device=dev, dtype=torch.float32).contiguous()
        ww, wb = self.get_parameter("word_embed.weight"), self.get_parameter("word_embed.bias")
        Lin = self.L - self.first_l
        single_pass = (self.forward_single_pass and ops.get_gemm_engine() == ops.ENGINE_TC_F16X3 and self.kv16)
        if single_pass:
            # ONE masked full-sequence pass, as the reference runs it (:622-636): every dense layer sees all B*L rows, the
            # block-causal mask (:168, query level >= key level) is applied inside one attention launch per block
            # (cvar_attn_blockcausal16: the mask is a per-scale key limit, no L x L bias tensor is built or read).
            L = self.L
            tr = self._transformer(B, lmax=L)
            x0 = self._buf("fwd_x0", (B * self.first_l, C))
            ops.prologue_rows(self.get_parameter("class_emb.weight"), self.get_parameter("cond_embed.weight"), self.pos_start,
                              tr.lvl_pos, label_B.contiguous(), cond_type.contiguous(), tr.cond_BD, tr.silu_cond, x0)
            tr.prologue_ada()
            xv = tr.x[:B * L].view(B, L, C)
            xv[:, :self.first_l].copy_(x0.view(B, self.first_l, C) if mask_first else self._start_rows_class_first(label_B, cond_type, tr))
            # x = word_embed(teacher tokens) + (lvl_embed + pos_1LC) for every later token: one batched K = 32 GEMM (:616-618)
            ops.gemm(xin, ww, wb, xv[:, self.first_l:], Lin, C, Cvae, lda=Cvae, epilogue=ops.EPI_BIAS_RESID,
                     resid=tr.lvl_pos[self.first_l:], ldr=C, strideR=0, batch=B, strideA=Lin * Cvae, strideW=0, strideO=L * C)
            tr.scale(L, 0, block_causal_lens=self.cfg.scale_lens)
            return tr.logits[:B * L].view(B, L, V).clone()
        tr = self._transformer(B)
        tr.prologue(label_B.contiguous(), cond_type.contiguous())
        if not mask_first:
            tr.x[:B * self.first_l].view(B, self.first_l, C).copy_(self._start_rows_class_first(label_B, cond_type, tr))
        out = torch.empty(B, self.L, V, device=dev, dtype=torch.float32)
        cur_L = 0
        for si, l in enumerate(self.cfg.scale_lens):
            if si > 0:
                # x = word_embed(teacher tokens) + (lvl_embed + pos_1LC): one batched K = 32 GEMM, the positional rows
                # enter as a residual shared by the batch (:616, :618)
                a0 = cur_L - self.first_l
                ops.gemm(xin[:, a0:a0 + l], ww, wb, tr.x, l, C, Cvae, lda=Cvae, epilogue=ops.EPI_BIAS_RESID,
                         resid=tr.lvl_pos[cur_L:cur_L + l], ldr=C, strideR=0, batch=B, strideA=Lin * Cvae, strideW=0,
                         strideO=l * C)
            tr.scale(l, cur_L)
            out[:, cur_L:cur_L + l].copy_(tr.logits[:B * l].view(B, l, V))
            cur_L += l
        return out


class _Transformer:
    """Workspaces and launch sequence of the AdaLN transformer for R rows (the part autoregressive_infer_cfg,
    conditional_infer_cfg and forward share): prologue -> per scale, depth x AdaLNSABlock + head -> logits."""

    def __init__(self, m: "ControlVAR", R: int, lmax: Optional[int] = None):
        """lmax: rows per sample the workspaces must hold (default: the longest scale; the whole pyramid for the
        single-pass teacher-forced forward)."""
        self.m, self.R = m, R
        cst = self.cst = m._constants()
        C, H, depth, V, T = m.C, m.num_heads, m.depth, m.V, m.L
        lmax = max(m.cfg.scale_lens) if lmax is None else lmax
        self.lvl_pos = cst["lvl_pos"]
        # ---- workspaces (cached across calls)
        self.cond_BD = m._buf("cond_BD", (R, C))
        self.silu_cond = m._buf("silu_cond", (R, C))
        self.ada = m._buf("ada", (depth, R, 6 * C))
        self.ada_head = m._buf("ada_head", (R, 2 * C))
        self.x = m._buf("x", (R * lmax, C))
        # engine 4 (f16x3): every dense-layer input is produced as an FP16 pair and never exists in fp32
        f16 = self.f16 = cst["f16"]
        kv16 = self.kv16 = f16 and m.kv16
        self.qbuf = None if kv16 else m._buf("q", (R * H * lmax * 64,))
        self.q16 = m._pair("q16", (R * H * lmax * 64,)) if kv16 else None
        self.xn16 = m._pair("xn16", (R * lmax, C)) if f16 else None
        self.attn_o16 = m._pair("attn_o16", (R * lmax, C)) if f16 else None
        self.hid16 = m._pair("hid16", (R * lmax, 4 * C)) if f16 else None
        self.xn = None if f16 else m._buf("xn", (R * lmax, C))
        self.attn_o = None if f16 else m._buf("attn_o", (R * lmax, C))
        self.hid = None if f16 else m._buf("hid", (R * lmax, 4 * C))
        # engine 3 (2-CTA all-TMA GEMM): every GEMM input is produced already split hi/lo; the '*_lo' halves live here
        split = ops.get_gemm_engine() == ops.ENGINE_TC_2CTA
        self.xn_lo = m._buf("xn_lo", (R * lmax, C)) if split else None
        self.attn_o_lo = m._buf("attn_o_lo", (R * lmax, C)) if split else None
        self.hid_lo = m._buf("hid_lo", (R * lmax, 4 * C)) if split else None
        self.logits = m._buf("logits", (R * lmax, V))
        self.caches = m._kv_caches16(depth, R, H, T) if kv16 else m._kv_caches(depth, R, H, T)

    def prologue(self, label_R: torch.Tensor, cond_R: torch.Tensor) -> None:
        """Start tokens of scale 0 into x, cond_BD, and every ada_lin = Linear(SiLU(cond)) (constant across scales)."""
        m, R, C, cst = self.m, self.R, self.m.C, self.cst
        ops.prologue_rows(m.get_parameter("class_emb.weight"),
                          m.get_parameter("cond_embed.weight") if m.multi_cond else None, m.pos_start,
                          self.lvl_pos, label_R, cond_R, self.cond_BD, self.silu_cond, self.x)
        self.prologue_ada()

    def prologue_ada(self) -> None:
        """Every ada_lin = Linear(SiLU(cond)) of the call (silu_cond must be filled)."""
        m, R, C, cst = self.m, self.R, self.m.C, self.cst
        silu16 = ops.F16Pair.from_tensor(self.silu_cond, out=m._pair("silu16", (R, C))) if self.f16 else None
        for bi, blk in enumerate(cst["blocks"]):
            ops.gemm(self.silu_cond, blk["ada_w"], blk["ada_b"], self.ada[bi], R, 6 * C, C, A16=silu16)
        ops.gemm(self.silu_cond, cst["head_ada_w"], cst["head_ada_b"], self.ada_head, R, 2 * C, C, A16=silu16)

    def scale(self, l: int, L_prev: int, block_causal_lens=None) -> None:
        """x (R*l, C) of one scale through all blocks (keys / values appended at L_prev) and the head -> self.logits.
        block_causal_lens: x holds the WHOLE pyramid (l = sum of the lens, L_prev = 0) and attention runs under the
        block-causal mask in one launch per block (cvar_attn_blockcausal16) - the single-pass ControlVAR.forward."""
        m, R, cst = self.m, self.R, self.cst
        C, H, V = m.C, m.num_heads, m.V
        M, cur_L = R * l, L_prev + l
        x, xn, xn_lo, xn16 = self.x, self.xn, self.xn_lo, self.xn16
        attn_o, attn_o_lo, attn_o16 = self.attn_o, self.attn_o_lo, self.attn_o16
        hid, hid_lo, hid16 = self.hid, self.hid_lo, self.hid16
        attn_scale = m.cfg.attn_scale
        for bi, blk in enumerate(cst["blocks"]):
            a = self.ada[bi]                               # (R, 6C): gamma1, gamma2, scale1, scale2, shift1, shift2
            g1, g2, s1, s2, b1, b2 = (a[:, k * C:(k + 1) * C] for k in range(6))
            ops.ln_modulate(x, s1, b1, 6 * C, xn, M, C, l, m.norm_eps, out_lo=xn_lo, out16=xn16)
            if self.kv16:
                ops.qkv_project16(xn16, blk["qkv_w"], blk["q_bias"], blk["k_bias"], blk["v_bias"], self.q16,
                                  self.caches[bi], R, l, L_prev, H, m.cos_attn, blk["scale_mul"])
                if block_causal_lens is not None:
                    ops.attn_blockcausal16(self.q16, self.caches[bi], None, R, H, block_causal_lens, attn_scale, out16=attn_o16)
                else:
                    ops.attn_kvcache16(self.q16, self.caches[bi], None, R, H, l, cur_L, attn_scale, out16=attn_o16)
            else:
                ops.qkv_project(xn, blk["qkv_w"], blk["q_bias"], blk["k_bias"], blk["v_bias"], self.qbuf, self.caches[bi],
                                R, l, L_prev, H, m.cos_attn, blk["scale_mul"], A_lo=xn_lo, A16=xn16)
                ops.attn_kvcache(self.qbuf, self.caches[bi], attn_o, R, H, l, cur_L, attn_scale, out_lo=attn_o_lo,
                                 out16=attn_o16)
            ops.gemm(attn_o, blk["proj_w"], blk["proj_b"], x, M, C, C, A_lo=attn_o_lo, A16=attn_o16,
                     epilogue=ops.EPI_BIAS_GAMMA_RESID, gamma=g1, gamma_row_stride=6 * C, rows_per_sample=l)
            ops.ln_modulate(x, s2, b2, 6 * C, xn, M, C, l, m.norm_eps, out_lo=xn_lo, out16=xn16)
            ops.gemm(xn, blk["fc1_w"], blk["fc1_b"], hid, M, 4 * C, C, epilogue=ops.EPI_BIAS_GELU, A_lo=xn_lo,
                     out_lo=hid_lo, A16=xn16, out16=hid16)
            ops.gemm(hid, blk["fc2_w"], blk["fc2_b"], x, M, C, 4 * C, A_lo=hid_lo, A16=hid16,
                     epilogue=ops.EPI_BIAS_GAMMA_RESID, gamma=g2, gamma_row_stride=6 * C, rows_per_sample=l)
        # head: AdaLNBeforeHead + Linear(C, V)
        ah = self.ada_head
        ops.ln_modulate(x, ah[:, :C], ah[:, C:], 2 * C, xn, M, C, l, m.norm_eps, out_lo=xn_lo, out16=xn16)
        ops.gemm(xn, cst["head_w"], m.get_parameter("head.bias"), self.logits, M, V, C, A_lo=xn_lo, A16=xn16)
