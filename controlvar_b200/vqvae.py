"""Host mirror of the reference VQVAE for the decode side of the sampling path.

Same constructor arguments, attribute names (``quantize``, ``Cvae``, ``vocab_size``) and ``state_dict`` keys as
/root/reference/models/vqvae.py:16-48, so checkpoints load with ``load_state_dict``.  ``fhat_to_img``
(models/vqvae.py:88-89; alias ``decode``) runs the whole Decoder (models/vae_modules.py:163-226) through
libcvar_sm100.so: NHWC activations, GroupNorm folded to a per-(sample, channel) affine that the next convolution
applies (with SiLU) while it loads its operand, nearest-x2 upsampling folded into the convolution's gather, residual
adds and the final clamp / de-normalise / NCHW store folded into epilogues.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .config import PathConfig, DEFAULT_PATCH_NUMS
from .weights import decoder_plan, encoder_plan, vae_key_shapes

_BUFFER_KEYS = ("ema_vocab_hit_SV",)


def bicubic_matrix(n_in: int, n_out: int) -> torch.Tensor:
    """(n_out, n_in) matrix U of F.interpolate(mode='bicubic', align_corners=False) along one axis, read off impulse
    responses so that the coefficients (A = -0.75, border clamping) are exactly ATen's.  Host-side constant."""
    eye = torch.eye(n_in, dtype=torch.float32).view(n_in, 1, 1, n_in)          # one impulse per batch entry
    out = F.interpolate(eye, size=(1, n_out), mode="bicubic")                  # height 1 -> 1 is the identity
    return out[:, 0, 0, :].t().contiguous()                                    # U[X, j]


def phi_index(si: int, SN: int, K: int = 4) -> int:
    """PhiPartiallyShared.__getitem__(si / (SN - 1)) - quant.py:282-293."""
    ticks = np.linspace(1 / 3 / K, 1 - 1 / 3 / K, K) if K == 4 else np.linspace(1 / 2 / K, 1 - 1 / 2 / K, K)
    return int(np.argmin(np.abs(ticks - si / (SN - 1))).item())


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's dotted parameter names."""


def register_tree(root: nn.Module, key: str, tensor: torch.Tensor, is_buffer: bool) -> None:
    parts = key.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            mod.add_module(p, _Node())
        mod = getattr(mod, p)
    if is_buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class VQVAE(nn.Module):
    def __init__(
        self, vocab_size=4096, z_channels=32, ch=128, dropout=0.0, beta=0.25, using_znorm=False, quant_conv_ks=3,
        quant_resi=0.5, share_quant_resi=4, default_qresi_counts=0, v_patch_nums=DEFAULT_PATCH_NUMS, test_mode=True,
    ):
        super().__init__()
        if using_znorm or quant_conv_ks != 3 or share_quant_resi != 4 or abs(quant_resi - 0.5) > 1e-9:
            raise NotImplementedError("only the released VQVAE configuration (vae_ch160v4096z32) is implemented")
        if z_channels != 32:
            raise NotImplementedError("the sm_100a kernels are specialised for Cvae = 32")
        self.test_mode = test_mode
        self.V, self.Cvae = vocab_size, z_channels
        self.vocab_size = vocab_size
        self.cfg = PathConfig(patch_nums=tuple(v_patch_nums), vocab_size=vocab_size, Cvae=z_channels, vae_ch=ch,
                              share_quant_resi=share_quant_resi)
        self.downsample = 2 ** (len(self.cfg.vae_ch_mult) - 1)
        for key, shape in vae_key_shapes(self.cfg, with_encoder=True).items():
            is_buf = key.split(".")[-1] in _BUFFER_KEYS
            register_tree(self, key, torch.zeros(shape), is_buf)
        # attribute surface of VectorQuantizer2 that callers read (quant.py:15-37)
        q = self.quantize
        q.vocab_size, q.Cvae, q.v_patch_nums, q.share_quant_resi = vocab_size, z_channels, tuple(v_patch_nums), share_quant_resi
        self._plan = decoder_plan(self.cfg)
        self._enc_plan = encoder_plan(self.cfg)
        self._U: Dict[Tuple[int, int], torch.Tensor] = {}
        # test instrument (as ControlVAR.debug_forced_idx): per-scale (B, pn*pn) tokens that replace the argmin result
        # before the residual update of f_to_idxBl, pinning the residual trajectory to the oracle's
        self.debug_forced_idx: Optional[List[torch.Tensor]] = None
        self.last_idx: List[torch.Tensor] = []
        # Accuracy policy of the convolutions (measured on B200, profiles/r01_decoder_policy.md; bound: 1e-4 abs per pixel).
        # The tensor core truncates its fp32 accumulator after every MMA, so one long accumulation (K = 9 * 640 = 360
        # steps) is what costs accuracy.  Engine 4 (f16x3): every layer runs on tensor cores and 3x3 layers with
        # 9 * Cin >= ksplit_min_k run their three kernel rows as three accumulations summed in fp32 (cvar_conv_args.ksplit):
        # worst pixel 4.3e-5 over 32 realistic images, 125 ms per 64-image decode (without the split: 1.09e-4, over the
        # bound; with the 16x16 layers on SIMT instead: 7.1e-5, 157 ms).  The 3xTF32 engines (1, 3) have no split and keep
        # layers with an output side below 64 on the SIMT fp32 engine (8.0e-5).
        # tc_min_hw: None = that per-engine default (16 / 64); an int overrides it (tools/diag_decoder_policy.py).
        self.tc_min_hw: Optional[int] = None
        self.ksplit_min_k = 2880          # 0 switches the K-split off
        self._packed: Dict[str, torch.Tensor] = {}
        self._packed16: Dict[str, "ops.F16Pair"] = {}
        self._ws: Dict[Tuple, torch.Tensor] = {}
        self._pending_stats: Dict[int, tuple] = {}   # data_ptr of a conv output -> (GroupNorm partials, B, HW, C)
        self.fuse_gn_stats = True                     # GroupNorm statistics from the producing conv's epilogue
        self._ws_gen = 0       # bumped when a workspace or a derived weight is (re)made: ControlVAR's CUDA graphs hold pointers
        self.eval()

    # -------------------------------------------------------------------------------------------- weights
    def load_state_dict(self, state_dict, strict=True, assign=False):
        # models/vqvae.py:106-109: tolerate a different number of scales in the usage statistics buffer
        sd = dict(state_dict)
        k = "quantize.ema_vocab_hit_SV"
        if k in sd and sd[k].shape[0] != self.quantize.ema_vocab_hit_SV.shape[0]:
            sd[k] = self.quantize.ema_vocab_hit_SV
        self._packed.clear()
        self._packed16.clear()
        self._ws_gen += 1
        return super().load_state_dict(sd, strict=strict, assign=assign)

    def _apply(self, fn, recurse=True):
        self._packed.clear()
        self._packed16.clear()
        self._ws.clear()
        self._U.clear()
        self._ws_gen += 1
        return super()._apply(fn, recurse)

    def _w(self, key: str) -> torch.Tensor:
        return self.get_parameter(key)

    def _conv_w(self, prefix: str, cin_pad: int = 0) -> torch.Tensor:
        """Conv weight repacked (Cout, ks*ks*Cin) tap-major for the implicit-GEMM kernels (cached).  cin_pad: zero
        weights for padding input channels (Encoder.conv_in: the 3-channel image runs as a 16-channel NHWC tensor)."""
        key = prefix + ".weight"
        t = self._packed.get(key)
        if t is None:
            w = self._w(key)
            cin = max(cin_pad, w.shape[1])
            t = torch.empty(w.shape[0], cin * w.shape[2] * w.shape[3], device=w.device, dtype=torch.float32)
            if cin != w.shape[1]:
                ops.repack_conv_weight_pad(w.contiguous(), t, cin)
            else:
                ops.repack_conv_weight(w.contiguous(), t)
            if t.numel() % 4 == 0:
                t = ops.SplitWeight(t)       # + TF32 hi/lo split for the tcgen05 engine
            self._packed[key] = t
        return t

    def _conv_w16(self, prefix: str) -> "ops.F16Pair":
        """FP16 pair of the repacked weight, for the f16x3 convolution kernel (cached)."""
        key = prefix + ".weight"
        p = self._packed16.get(key)
        if p is None:
            w = self._conv_w(prefix)
            p = ops.F16Pair.from_tensor(w.w if isinstance(w, ops.SplitWeight) else w)
            self._packed16[key] = p
        return p

    def _gemm_w16(self, prefix: str) -> "ops.SplitWeight":
        """A 1x1 conv weight (AttnBlock q / k / v / proj_out) as the FP16-pair GEMM weight of the f16x3 engine (cached)."""
        key = prefix + ".weight/gemm16"
        w = self._packed.get(key)
        if w is None:
            base = self._conv_w(prefix)
            w = ops.SplitWeight(base.w if isinstance(base, ops.SplitWeight) else base, f16=True)
            self._packed[key] = w
        return w

    def _min_hw(self) -> int:
        if self.tc_min_hw is not None:
            return self.tc_min_hw
        return 16 if ops.get_gemm_engine() == ops.ENGINE_TC_F16X3 else 64

    def _f16_layer(self, Hout, Wout, Cin, Cout, ks) -> bool:
        """Does this layer run on the FP16-pair TMA kernel?  Engine 4, accuracy policy (tc_min_hw), supported shape."""
        return (ops.get_gemm_engine() == ops.ENGINE_TC_F16X3 and Hout >= self._min_hw()
                and ops.conv2d_f16_supported(Hout, Wout, Cin, Cout, ks))

    def _conv(self, x, prefix, out, B, Hin, Win, Cin, Cout, ks, **kw):
        """x: fp32 NHWC tensor, or an F16Pair (normalised / upsampled by its producer) for a layer _f16_layer() accepts."""
        bias = self._w(prefix + ".bias")
        self._pending_stats.pop(out.data_ptr(), None)
        stats_role = kw.pop("stats_role", None)
        if isinstance(x, ops.F16Pair):
            split = 3 if (self.ksplit_min_k and ks == 3 and 9 * Cin >= self.ksplit_min_k and not kw.get("out_mode")) else 0
            part = None
            if (stats_role is not None and self.fuse_gn_stats and not kw.get("out_mode")
                    and ops.conv2d_gn_fusable(Hin, Win, Cin, Cout, ks)):
                # the consumer of this output is a GroupNorm: its statistics come out of this conv's epilogue
                part = self._buf("gn_part_" + stats_role, (2 * B * 32 * (Hin * Win // 32),), torch.float64)
                self._pending_stats[out.data_ptr()] = (part, B, Hin * Win, Cout)
            return ops.conv2d(None, self._conv_w(prefix), bias, out, B, Hin, Win, Cin, Cout, ks, x16=x,
                              w16=self._conv_w16(prefix), ksplit=split, gn_part=part, **kw)
        hout = Hin * (2 if kw.get("upsample2x") else 1)
        return ops.conv2d(x, self._conv_w(prefix), bias, out, B, Hin, Win, Cin, Cout, ks,
                          engine=(-1 if hout >= self._min_hw() else 0), **kw)

    def _pair(self, name: str, numel: int, shape) -> "ops.F16Pair":
        n = 1
        for s_ in shape:
            n *= s_
        return ops.F16Pair(self._buf(name + ".hi", (numel,), torch.float16)[:n].view(shape),
                           self._buf(name + ".lo", (numel,), torch.float16)[:n].view(shape))

    def _buf(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        """Named workspace, grown to the largest size ever asked for (one allocation per name, not one per shape: a new
        batch size re-uses or replaces the buffer instead of adding to it)."""
        dev = self._w("post_quant_conv.weight").device
        key = (name, dtype)
        n = 1
        for s_ in shape:
            n *= int(s_)
        t = self._ws.get(key)
        if t is None or t.device != dev or t.numel() < n:
            t = torch.empty(max(n, 1), device=dev, dtype=dtype)
            self._ws[key] = t
            self._ws_gen += 1          # captured CUDA graphs (ControlVAR._graphs) hold workspace pointers
        return t[:n].view(tuple(shape))

    def release_workspace(self):
        self._ws.clear()
        self._ws_gen += 1

    # -------------------------------------------------------------------------------------------- decoder
    def _gn(self, x, prefix: str, B, HW, Cn, slot: int):
        a = self._buf(f"gn_a{slot}", (B, Cn))
        b = self._buf(f"gn_b{slot}", (B, Cn))
        pend = self._pending_stats.pop(x.data_ptr(), None)
        if pend is not None and pend[1:] == (B, HW, Cn):
            # partial sums written by the epilogue of the convolution that produced x (cvar_conv_args.gn_part)
            ops.gn_finalize_parts(pend[0], self._w(prefix + ".weight"), self._w(prefix + ".bias"), a, b, B, HW, Cn)
            return a, b
        scratch = self._buf("gn_scratch", (2 * B * 32 * ops.gn_chunks(HW),), torch.float64)
        ops.gn_stats(x, self._w(prefix + ".weight"), self._w(prefix + ".bias"), a, b, scratch, B, HW, Cn)
        return a, b

    def _norm_act(self, x, prefix, B, H, W, Cn, slot: int, pair: bool = False):
        """silu(GroupNorm(x)) materialised once (vae_modules.py:58-59).  One HBM-bound pass instead of re-evaluating
        the normalisation + SiLU for each of the 9 taps inside the convolution's operand gather.  pair: write it as the
        FP16 pair the f16x3 convolution fetches by TMA (same bytes as fp32)."""
        a, b = self._gn(x, prefix, B, H * W, Cn, slot)
        if pair:
            y16 = self._pair("normact16", self._act_numel, (B, H, W, Cn))
            ops.affine_nc(x, a, b, None, B, H * W, Cn, silu=True, out16=y16)
            return y16
        y = self._buf("normact", (self._act_numel,))[:B * H * W * Cn].view(B, H, W, Cn)
        ops.affine_nc(x, a, b, y, B, H * W, Cn, silu=True)
        return y

    def _resblock(self, x, prefix, B, H, W, cin, cout, bufs):
        """ResnetBlock.forward (vae_modules.py:57-60); x is never written."""
        h1, out = bufs
        self._conv(self._norm_act(x, prefix + "norm1", B, H, W, cin, 0, pair=self._f16_layer(H, W, cin, cout, 3)),
                   prefix + "conv1", h1, B, H, W, cin, cout, 3, stats_role="h")
        if cin != cout:
            sc = self._buf("shortcut", (B, H, W, cout))
            self._conv(x, prefix + "nin_shortcut", sc, B, H, W, cin, cout, 1)
        else:
            sc = x
        self._conv(self._norm_act(h1, prefix + "norm2", B, H, W, cout, 1, pair=self._f16_layer(H, W, cout, cout, 3)),
                   prefix + "conv2", out, B, H, W, cout, cout, 3, resid=sc, stats_role="o")
        return out

    def _attnblock(self, x, prefix, B, H, W, Cn, out):
        """AttnBlock.forward (vae_modules.py:73-92) on NHWC: single head over HW positions."""
        HW = H * W
        a, b = self._gn(x, prefix + "norm", B, HW, Cn, 0)
        f16 = ops.get_gemm_engine() == ops.ENGINE_TC_F16X3 and Cn % 64 == 0
        qkv = self._buf("attn_qkv", (B, HW, 3 * Cn))
        if f16:       # the dense layers of the block on the f16x3 2-CTA kernel, like the transformer's (round 1: 3xTF32 1-CTA)
            xn16 = self._pair("attn_xn16", B * HW * Cn, (B * HW, Cn))
            ops.affine_nc(x, a, b, None, B, HW, Cn, silu=False, out16=xn16)
            ops.gemm(None, self._gemm_w16(prefix + "qkv"), self._w(prefix + "qkv.bias"), qkv, B * HW, 3 * Cn, Cn, A16=xn16)
        else:
            xn = self._buf("attn_xn", (B, HW, Cn))
            ops.affine_nc(x, a, b, xn, B, HW, Cn, silu=False)
            ops.gemm(xn, self._conv_w(prefix + "qkv"), self._w(prefix + "qkv.bias"), qkv, B * HW, 3 * Cn, Cn)
        S = self._buf("attn_S", (B, HW, HW))
        # w = bmm(q, k).mul_(C ** -0.5)
        ops.gemm(qkv, qkv[:, :, Cn:], None, S, HW, HW, Cn, lda=3 * Cn, ldw=3 * Cn, ldo=HW, alpha=int(Cn) ** (-0.5),
                 batch=B, strideA=HW * 3 * Cn, strideW=HW * 3 * Cn, strideO=HW * HW)
        ops.softmax_rows(S, B * HW, HW)
        hbuf = self._buf("attn_h", (B, HW, Cn))
        # h[i, c] = sum_j P[i, j] v[j, c]
        ops.gemm(S, qkv[:, :, 2 * Cn:], None, hbuf, HW, Cn, HW, lda=HW, ldw=3 * Cn, ldo=Cn, w_is_kn=True, batch=B,
                 strideA=HW * HW, strideW=HW * 3 * Cn, strideO=HW * Cn)
        if f16:
            h16 = ops.F16Pair.from_tensor(hbuf, out=self._pair("attn_h16", B * HW * Cn, (B, HW, Cn)))
            ops.gemm(None, self._gemm_w16(prefix + "proj_out"), self._w(prefix + "proj_out.bias"), out, B * HW, Cn, Cn,
                     epilogue=ops.EPI_BIAS_RESID, resid=x, A16=h16)
        else:
            ops.gemm(hbuf, self._conv_w(prefix + "proj_out"), self._w(prefix + "proj_out.bias"), out, B * HW, Cn, Cn,
                     epilogue=ops.EPI_BIAS_RESID, resid=x)
        return out

    def _decode_nhwc(self, z_nhwc: torch.Tensor, B: int, hw: int, img_out: torch.Tensor, rows_total: int,
                     row_offset: int, out_mode: int = 1, out_samples: int = 0) -> None:
        """post_quant_conv + Decoder.forward + clamp + (x+1)/2, image planes written into img_out (NCHW).
        out_samples > 0: the B maps are stacked groups of out_samples (cvar_conv_args.out_samples)."""
        cfg = self.cfg
        H = W = hw
        zq = self._buf("z_pq", (B, H, W, cfg.Cvae))
        if self._f16_layer(H, W, cfg.Cvae, cfg.Cvae, 3):
            z16 = ops.F16Pair.from_tensor(z_nhwc, out=self._pair("z_in16", B * H * W * cfg.Cvae, (B, H, W, cfg.Cvae)))
            self._conv(z16, "post_quant_conv", zq, B, H, W, cfg.Cvae, cfg.Cvae, 3)
        else:
            self._conv(z_nhwc, "post_quant_conv", zq, B, H, W, cfg.Cvae, cfg.Cvae, 3)
        cur = zq
        ring = self._ring_setup(self._plan, B, hw)

        for op, prefix, cin, cout in self._plan:
            if op == "conv3":
                out = ring((B, H, W, cout))
                if self._f16_layer(H, W, cin, cout, 3):
                    c16 = ops.F16Pair.from_tensor(cur, out=self._pair("z_pq16", B * H * W * cin, (B, H, W, cin)))
                    self._conv(c16, prefix, out, B, H, W, cin, cout, 3, stats_role="o")
                else:
                    self._conv(cur, prefix, out, B, H, W, cin, cout, 3)
                cur = out
            elif op == "res":
                h1 = ring((B, H, W, cout))
                out = ring((B, H, W, cout))
                cur = self._resblock(cur, prefix, B, H, W, cin, cout, (h1, out))
            elif op == "attn":
                out = ring((B, H, W, cout))
                cur = self._attnblock(cur, prefix, B, H, W, cin, out)
            elif op == "up":
                out = ring((B, 2 * H, 2 * W, cout))
                if self._f16_layer(2 * H, 2 * W, cin, cout, 3):
                    # Upsample2x (vae_modules.py:27-28): nearest x2 written once as the FP16 pair the conv fetches by TMA
                    up16 = self._pair("normact16", self._act_numel, (B, 2 * H, 2 * W, cin))
                    ops.upsample2x_split_f16(cur, up16, B, H, W, cin)
                    self._conv(up16, prefix, out, B, 2 * H, 2 * W, cin, cout, 3, stats_role="o")
                else:
                    self._conv(cur, prefix, out, B, H, W, cin, cout, 3, upsample2x=True)
                H, W = 2 * H, 2 * W
                cur = out
            elif op == "out":
                self._conv(self._norm_act(cur, "decoder.norm_out", B, H, W, cin, 0), "decoder.conv_out", img_out, B, H, W,
                           cin, 3, 3, out_mode=out_mode, out_rows_total=rows_total, row_offset=row_offset,
                           out_samples=out_samples)
            else:
                raise AssertionError(op)

    @torch.no_grad()
    def fhat_to_img(self, f_hat: torch.Tensor) -> torch.Tensor:
        """decoder(post_quant_conv(f_hat)).clamp_(-1, 1) - models/vqvae.py:88-89.  f_hat: (B, Cvae, h, w) NCHW
        (may be a strided view such as the halves of control_var.py:525-526).  Returns (B, 3, 16h, 16w) in [-1, 1]."""
        return self._fhat_to_img(f_hat, out_mode=2)

    decode = fhat_to_img   # the name BASELINE.json's north_star uses for this entry point

    @torch.no_grad()
    def _fhat_to_img(self, f_hat: torch.Tensor, out: Optional[torch.Tensor] = None, rows_total: int = 0,
                     row_offset: int = 0, out_mode: int = 1) -> torch.Tensor:
        """out_mode 1: clamp + (x+1)/2 fused (what control_var.py:563-565 does right after); 2: clamp only."""
        if not f_hat.is_cuda:
            raise RuntimeError("controlvar_b200.VQVAE runs on CUDA only (no CPU fallback)")
        B, Cz, h, w = f_hat.shape
        assert Cz == self.Cvae and h == w
        if f_hat.stride(3) != 1 or f_hat.stride(2) != w or f_hat.stride(0) != Cz * f_hat.stride(1):
            f_hat = f_hat.contiguous()
        z = self._buf("z_nhwc", (B, h, w, Cz))
        ops.nchw_to_nhwc(f_hat, z, B, Cz, h, w, f_hat.stride(0))
        side = h * self.downsample
        if out is None:
            out = torch.empty(B, 3, side, side, device=f_hat.device, dtype=torch.float32)
            rows_total, row_offset = side, 0
        self._decode_nhwc(z, B, h, out, rows_total, row_offset, out_mode)
        return out

    @torch.no_grad()
    def _fhat_halves_to_img(self, f_hat2: torch.Tensor, B: int, out_mode: int = 1) -> torch.Tensor:
        """control_var.py:563-565 in ONE decoder pass: f_hat2 (>= B, Cvae, 2*hw, hw) holds the control map (rows [0, hw))
        and the image map (rows [hw, 2hw)) of every sample; both are decoded as one batch of 2B maps (control maps first)
        and written as (B, 3, 2*side, side), control on top.  Each map is an independent decoder input (GroupNorm is per
        image), so the pixels equal two fhat_to_img calls; twice the rows per launch, half the launches."""
        if not f_hat2.is_cuda:
            raise RuntimeError("controlvar_b200.VQVAE runs on CUDA only (no CPU fallback)")
        _, Cz, h2, w = f_hat2.shape
        assert Cz == self.Cvae and h2 == 2 * w and f_hat2.is_contiguous()
        z = self._buf("z_nhwc", (2 * B, w, w, Cz))
        ops.nchw_to_nhwc(f_hat2, z[:B], B, Cz, w, w, f_hat2.stride(0))
        ops.nchw_to_nhwc(f_hat2[:, :, w:, :], z[B:], B, Cz, w, w, f_hat2.stride(0))
        side = w * self.downsample
        out = torch.empty(B, 3, 2 * side, side, device=f_hat2.device, dtype=torch.float32)
        self._decode_nhwc(z, 2 * B, w, out, 2 * side, 0, out_mode, out_samples=B)
        return out

    # -------------------------------------------------------------------------------------------- encoder
    def _ring_setup(self, plan, B: int, hw0: int):
        """Three flat activation buffers sized for the largest tensor of the plan, handed out round-robin."""
        act_numel, hh = 0, hw0
        for op, _, cin, cout in plan:
            if op == "up":
                hh *= 2
            elif op == "down":
                hh //= 2
            act_numel = max(act_numel, B * hh * hh * max(cin if op == "out" else cout, 1))
        self._act_numel = act_numel
        state = {"i": 0}

        def ring(shape):
            state["i"] = (state["i"] + 1) % 3
            n = 1
            for s_ in shape:
                n *= s_
            t = self._buf(f"act{state['i']}", (act_numel,))[:n].view(shape)
            self._pending_stats.pop(t.data_ptr(), None)       # the buffer is about to be overwritten
            return t
        return ring

    @torch.no_grad()
    def _img_to_f(self, img: torch.Tensor) -> torch.Tensor:
        """quant_conv(encoder(img)) - vqvae.py:74, Encoder.forward vae_modules.py:145-160.  img (B, 3, H, W) in [-1, 1];
        returns f (B, Cvae, H/16, W/16) NCHW (a fresh tensor)."""
        if not img.is_cuda:
            raise RuntimeError("controlvar_b200.VQVAE runs on CUDA only (no CPU fallback)")
        B, Ci, H, W = img.shape
        assert Ci == 3 and H == W and H % self.downsample == 0, "img must be (B, 3, S, S) with S a multiple of 16"
        img = img.to(torch.float32).contiguous()
        cfg = self.cfg
        CPAD = 16
        ring = self._ring_setup(self._enc_plan, B, H)
        cur = None
        for op, prefix, cin, cout in self._enc_plan:
            if op == "conv_in":
                x0 = self._buf("enc_x0", (B, H, W, CPAD))
                ops.nchw_to_nhwc_pad(img, x0, B, 3, H, W, CPAD)
                out = ring((B, H, W, cout))
                ops.conv2d(x0, self._conv_w(prefix, cin_pad=CPAD), self._w(prefix + ".bias"), out, B, H, W, CPAD, cout, 3,
                           engine=0)
                cur = out
            elif op == "res":
                h1 = ring((B, H, W, cout))
                out = ring((B, H, W, cout))
                cur = self._resblock(cur, prefix, B, H, W, cin, cout, (h1, out))
            elif op == "attn":
                out = ring((B, H, W, cout))
                cur = self._attnblock(cur, prefix, B, H, W, cin, out)
            elif op == "down":
                out = ring((B, H // 2, W // 2, cout))
                ops.conv2d(cur, self._conv_w(prefix), self._w(prefix + ".bias"), out, B, H, W, cin, cout, 3,
                           downsample2x=True, engine=0)
                H, W = H // 2, W // 2
                cur = out
            elif op == "out":
                z = self._buf("enc_z", (B, H, W, cout))
                self._conv(self._norm_act(cur, "encoder.norm_out", B, H, W, cin, 0), "encoder.conv_out", z, B, H, W, cin,
                           cout, 3)
                f = torch.empty(B, cout, H, W, device=img.device, dtype=torch.float32)
                # quant_conv, written NCHW (the layout of f_rest / f_hat in quant.py:184-215)
                self._conv(z, "quant_conv", f, B, H, W, cout, cout, 3, out_mode=3, out_rows_total=H, row_offset=0)
                return f
            else:
                raise AssertionError(op)
        raise AssertionError("encoder plan without an output record")

    @torch.no_grad()
    def _f_to_idxBl(self, f: torch.Tensor, v_patch_nums: Sequence[int], to_fhat: bool = False) -> List[torch.Tensor]:
        """VectorQuantizer2.f_to_idxBl_or_fhat - quant.py:184-215: per scale, area-pool the residual, nearest code, then
        f_hat += phi(bicubic(E[idx])), f_rest -= the same, in one kernel.  to_fhat: return f_hat after every scale
        (clones) instead of the token ids."""
        B, Cz, H, W = f.shape
        pns = [int(pn) for pn in v_patch_nums]
        assert Cz == self.Cvae and H == W and pns[-1] == H, f"patch_nums[-1]={pns[-1]} != H={H}"
        SN = len(pns)
        f_rest = f.detach().clone().contiguous()
        f_hat = torch.zeros_like(f_rest)
        emb = self._w("quantize.embedding.weight")
        out: List[torch.Tensor] = []
        self.last_idx = []
        for si, pn in enumerate(pns):
            N = B * pn * pn
            z = self._buf("vq_z", (B * H * W, Cz))[:N]
            ops.area_pool_nc(f_rest, z, B, Cz, H, pn)
            idx = torch.empty(N, dtype=torch.int64, device=f.device)
            ops.vq_nearest(z, emb, idx)
            self.last_idx.append(idx.view(B, pn * pn).clone())
            if self.debug_forced_idx is not None:
                idx.copy_(self.debug_forced_idx[si].to(device=f.device, dtype=torch.int64).reshape(-1))
            k = phi_index(si, SN, self.cfg.share_quant_resi) if SN > 1 else 0
            U = None
            if pn != H:
                U = self._U.get((pn, H))
                if U is None:
                    U = self._U[(pn, H)] = bicubic_matrix(pn, H).to(f.device)
            ops.vq_step(idx, emb, U, self._w(f"quantize.quant_resi.qresi_ls.{k}.weight"),
                        self._w(f"quantize.quant_resi.qresi_ls.{k}.bias"), None, None, None, f_hat, None, B, pn, 0, H, Cz,
                        0, streams=1, x_replicas=1, f_rest=f_rest)
            out.append(f_hat.clone() if to_fhat else idx.view(B, pn * pn))
        return out

    def _phi_U(self, si: int, SN: int, pn: int, H: int, device):
        k = phi_index(si, SN, self.cfg.share_quant_resi) if SN > 1 else 0
        U = None
        if pn != H:
            U = self._U.get((pn, H))
            if U is None:
                U = self._U[(pn, H)] = bicubic_matrix(pn, H).to(device)
        return (self._w(f"quantize.quant_resi.qresi_ls.{k}.weight"), self._w(f"quantize.quant_resi.qresi_ls.{k}.bias"), U)

    def _check_tokens(self, ms_idx_Bl) -> Tuple[List[int], int, int]:
        pns = [int(round(t.shape[1] ** 0.5)) for t in ms_idx_Bl]          # vqvae.py:101
        assert all(pn * pn == t.shape[1] for pn, t in zip(pns, ms_idx_Bl)), "token maps must be square"
        if not ms_idx_Bl[0].is_cuda:
            raise RuntimeError("controlvar_b200.VQVAE runs on CUDA only (no CPU fallback)")
        B = ms_idx_Bl[0].shape[0]
        # the kernels index the codebook with these ids: a bad id must be an error, not an out-of-bounds read
        bad = None
        for t in ms_idx_Bl:
            if t.dim() != 2 or t.shape[0] != B:
                raise ValueError(f"token maps must all be (B, pn*pn) with the same B, got {tuple(t.shape)}")
            b = ((t < 0) | (t >= self.vocab_size)).any()
            bad = b if bad is None else (bad | b)
        if bool(bad):
            raise ValueError(f"token ids must be in [0, {self.vocab_size})")
        return pns, B, self.cfg.patch_nums[-1]

    @torch.no_grad()
    def idxBl_to_img(self, ms_idx_Bl: List[torch.Tensor], same_shape: bool, last_one=False):
        """Drop-in for VQVAE.idxBl_to_img (vqvae.py:97-104) with same_shape=True (embed_to_fhat(all_to_max_scale=True),
        quant.py:156-170): tokens -> f_hat accumulated scale by scale -> decoded image(s) in [-1, 1].  last_one: only
        the final image, else one image per scale.  same_shape=False is the reference's 'experimental visualisation'
        branch and is not implemented."""
        if not same_shape:
            raise NotImplementedError("idxBl_to_img(same_shape=False) is not implemented")
        pns, B, H = self._check_tokens(ms_idx_Bl)
        SN = len(self.cfg.patch_nums)
        dev = ms_idx_Bl[0].device
        emb = self._w("quantize.embedding.weight")
        f_hat = torch.zeros(B, self.Cvae, H, H, device=dev, dtype=torch.float32)
        imgs = []
        for si, (pn, idx) in enumerate(zip(pns, ms_idx_Bl)):
            pw, pb, U = self._phi_U(si, SN, pn, H, dev)
            ops.vq_step(idx.to(torch.int64).contiguous().view(-1), emb, U, pw, pb, None, None, None, f_hat, None, B, pn, 0, H,
                        self.Cvae, 0, streams=1, x_replicas=1)
            if not last_one:
                imgs.append(self.fhat_to_img(f_hat))
        return self.fhat_to_img(f_hat) if last_one else imgs

    @torch.no_grad()
    def idxBl_to_h(self, gt_ms_idx_Bl: List[torch.Tensor]) -> List[torch.Tensor]:
        """Drop-in for VQVAE.idxBl_to_h = VectorQuantizer2.idxBl_to_var_input (vqvae.py:77-78, quant.py:217-241): the
        teacher-forcing inputs of ControlVAR.forward, one (B, pn_next^2, Cvae) tensor per scale transition."""
        pns, B, H = self._check_tokens(gt_ms_idx_Bl)
        SN = len(self.cfg.patch_nums)
        dev = gt_ms_idx_Bl[0].device
        emb = self._w("quantize.embedding.weight")
        f_hat = torch.zeros(B, self.Cvae, H, H, device=dev, dtype=torch.float32)
        out = []
        for si in range(SN - 1):
            pn, pn_next = pns[si], self.cfg.patch_nums[si + 1]
            pw, pb, U = self._phi_U(si, SN, pn, H, dev)
            ops.vq_step(gt_ms_idx_Bl[si].to(torch.int64).contiguous().view(-1), emb, U, pw, pb, None, None, None, f_hat, None,
                        B, pn, 0, H, self.Cvae, 0, streams=1, x_replicas=1)
            z = torch.empty(B * pn_next * pn_next, self.Cvae, device=dev, dtype=torch.float32)
            ops.area_pool_nc(f_hat, z, B, self.Cvae, H, pn_next)      # area pool + (B, C, n) -> (B, n, C)
            out.append(z.view(B, pn_next * pn_next, self.Cvae))
        return out

    @torch.no_grad()
    def img_to_recon(self, x: torch.Tensor, v_patch_nums: Optional[Sequence[int]] = None, last_one=False):
        """Drop-in for VQVAE.img_to_recon (vqvae.py:80-86): decoder(post_quant_conv(f_hat)) of the quantised encoder
        output - NOT clamped, as in the reference - after the last scale (last_one) or after every scale."""
        pns = self.cfg.patch_nums if v_patch_nums is None else v_patch_nums
        fhats = self._f_to_idxBl(self._img_to_f(x), pns, to_fhat=True)
        if last_one:
            return self._fhat_to_img(fhats[-1], out_mode=3)
        return [self._fhat_to_img(fh, out_mode=3) for fh in fhats]

    @torch.no_grad()
    def img_to_idxBl(self, inp_img_no_grad: torch.Tensor,
                     v_patch_nums: Optional[Sequence[int]] = None) -> List[torch.Tensor]:
        """Drop-in for VQVAE.img_to_idxBl (vqvae.py:73-75): the multi-scale token ids, List[(B, pn*pn) int64], of an
        image in [-1, 1].  v_patch_nums defaults to the constructor's (the reference's callers always pass it)."""
        pns = self.cfg.patch_nums if v_patch_nums is None else v_patch_nums
        return self._f_to_idxBl(self._img_to_f(inp_img_no_grad), pns)

    def forward(self, *a, **k):
        raise NotImplementedError("VQVAE.forward is training-only in the reference and out of scope")
