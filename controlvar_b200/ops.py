"""Thin tensor-level wrappers over the C ABI (include/cvar.h).  PyTorch supplies device memory and the current
stream; every computation below happens inside libcvar_sm100.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import ConvArgs, GemmArgs, check

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_GAMMA_RESID, EPI_BIAS_RESID = 0, 1, 2, 3
ENGINE_SIMT, ENGINE_TC_3XTF32, ENGINE_TC_2CTA, ENGINE_TC_F16X3 = 0, 1, 3, 4
F16_LO_SCALE = 2048.0      # an FP16 pair represents hi + lo / 2048 (include/cvar.h: cvar_split_f16)
QK_SCALE = 16.0            # q / K of the f16 attention path are "qk pairs": 16 x = hi + lo (include/cvar.h)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.CvarError("controlvar_b200 ops need CUDA tensors (there is no CPU fallback)")
        if t.dtype not in (torch.float32, torch.int64, torch.float64):
            raise _lib.CvarError(f"unsupported dtype {t.dtype}")


class F16Pair:
    """An fp32 array held as two IEEE-half arrays, value = hi + lo / 2048: the operand format of the f16x3 engine."""
    __slots__ = ("hi", "lo")

    def __init__(self, hi: torch.Tensor, lo: torch.Tensor):
        for t in (hi, lo):
            if not t.is_cuda or t.dtype != torch.float16 or not t.is_contiguous():
                raise _lib.CvarError("F16Pair needs contiguous CUDA float16 tensors")
        if hi.shape != lo.shape:
            raise _lib.CvarError("F16Pair: hi / lo shapes differ")
        self.hi, self.lo = hi, lo

    @staticmethod
    def empty(shape, device) -> "F16Pair":
        return F16Pair(torch.empty(shape, dtype=torch.float16, device=device),
                       torch.empty(shape, dtype=torch.float16, device=device))

    @staticmethod
    def from_tensor(x: torch.Tensor, out: Optional["F16Pair"] = None) -> "F16Pair":
        _chk(x)
        x = x.contiguous()
        if x.numel() % 4 != 0:
            raise _lib.CvarError("F16Pair: number of elements must be a multiple of 4")
        p = out if out is not None else F16Pair.empty(x.shape, x.device)
        check(_lib.load().cvar_split_f16(_p(x), _p(p.hi), _p(p.lo), x.numel(), _stream()), "cvar_split_f16")
        return p

    def float(self) -> torch.Tensor:
        return self.hi.float() + self.lo.float() / F16_LO_SCALE

    # ---- "qk pairs" (q and K of cvar_qkv_project16 / cvar_attn_kvcache16): 16 x = hi + lo, residual NOT scaled.
    # The library writes them in the QKV epilogue; these two helpers exist for tests and tools only.
    @staticmethod
    def from_tensor_qk(x: torch.Tensor) -> "F16Pair":
        _chk(x)
        xs = x.contiguous() * QK_SCALE
        hi = xs.to(torch.float16)
        return F16Pair(hi, (xs - hi.float()).to(torch.float16))

    def float_qk(self) -> torch.Tensor:
        return (self.hi.float() + self.lo.float()) / QK_SCALE


def _p16(p: Optional[F16Pair]):
    return (None, None) if p is None else (p.hi.data_ptr(), p.lo.data_ptr())


# ---- optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline object) -------
_prof = None


def profile_begin():
    global _prof
    _prof = []


def profile_end():
    """-> {tag: dict(launches, work, bytes, ms)}; call after torch.cuda.synchronize()."""
    global _prof
    rec, _prof = _prof, None
    out = {}
    for tag, work, nbytes, e0, e1 in rec or []:
        d = out.setdefault(tag, dict(launches=0, work=0.0, bytes=0.0, ms=0.0))
        d["launches"] += 1
        d["work"] += work
        d["bytes"] += nbytes
        d["ms"] += e0.elapsed_time(e1)
    return out


class _Timed:
    __slots__ = ("tag", "work", "nbytes", "e0")

    def __init__(self, tag, work, nbytes=0.0):
        self.tag, self.work, self.nbytes = tag, work, nbytes

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.append((self.tag, self.work, self.nbytes, self.e0, e1))
        return False


def profiling() -> bool:
    return _prof is not None


def launch_count() -> int:
    return int(_lib.load().cvar_launch_count())


def add_launch_count(n: int) -> None:
    """Account for kernels launched by a CUDA-graph replay (the library only counts the launches it issues itself)."""
    _lib.load().cvar_add_launch_count(int(n))


def set_fast_mode(on: bool) -> int:
    """NOT a parity mode: single-MMA FP16 operands (hi halves only) in the tcgen05 GEMM / conv / attention kernels."""
    return int(_lib.load().cvar_set_fast_mode(int(bool(on))))


def get_fast_mode() -> int:
    return int(_lib.load().cvar_get_fast_mode())


def set_gemm_engine(engine: int) -> int:
    return int(_lib.load().cvar_set_gemm_engine(int(engine)))


def set_epilogue_overlap(on: bool) -> int:
    return int(_lib.load().cvar_set_epilogue_overlap(int(bool(on))))


def set_tc_kblock(bk: int) -> int:
    return int(_lib.load().cvar_set_tc_kblock(int(bk)))


def get_gemm_engine() -> int:
    return int(_lib.load().cvar_get_gemm_engine())


def lvl_pos(lvl_embed, lvl_1L, pos_1LC, out):
    _chk(lvl_embed, lvl_1L, pos_1LC, out)
    T, Cdim = pos_1LC.shape[-2], pos_1LC.shape[-1]
    check(_lib.load().cvar_lvl_pos(_p(lvl_embed), _p(lvl_1L), _p(pos_1LC), _p(out), T, Cdim, _stream()), "cvar_lvl_pos")
    return out


def prologue(class_emb, cond_embed, pos_start, lvl_pos_t, label_B, cond_type_B, num_classes, cond_BD, silu_cond, x0):
    _chk(class_emb, cond_embed, pos_start, lvl_pos_t, label_B, cond_type_B, cond_BD, silu_cond, x0)
    B, Cdim = label_B.shape[0], class_emb.shape[1]
    check(_lib.load().cvar_prologue(_p(class_emb), _p(cond_embed), _p(pos_start), _p(lvl_pos_t), _p(label_B),
                                    _p(cond_type_B), B, Cdim, num_classes, _p(cond_BD), _p(silu_cond), _p(x0),
                                    _stream()), "cvar_prologue")


def prologue_rows(class_emb, cond_embed, pos_start, lvl_pos_t, label_R, cond_type_R, cond_BD, silu_cond, x0):
    """Explicit per-row class / condition-type ids (the four guidance replicas of conditional_infer_cfg)."""
    _chk(class_emb, cond_embed, pos_start, lvl_pos_t, label_R, cond_type_R, cond_BD, silu_cond, x0)
    R, Cdim = label_R.shape[0], class_emb.shape[1]
    check(_lib.load().cvar_prologue_rows(_p(class_emb), _p(cond_embed), _p(pos_start), _p(lvl_pos_t), _p(label_R),
                                         _p(cond_type_R), R, Cdim, _p(cond_BD), _p(silu_cond), _p(x0), _stream()),
          "cvar_prologue_rows")


def ln_modulate(x, scale, shift, mod_row_stride, out, M, Cdim, rows_per_sample, eps, out_lo=None,
                out16: Optional[F16Pair] = None):
    """scale/shift: views into an ada_lin output; only their data pointers and the common row stride are used.
    out_lo: write the result as a TF32 hi/lo split (out = hi) - the operand format of the 2-CTA GEMM.
    out16: write the result as an FP16 pair (f16x3 engine); ``out`` may then be None."""
    _chk(x, scale, shift, out, out_lo)
    h16, l16 = _p16(out16)
    check(_lib.load().cvar_ln_modulate(_p(x), _p(scale), _p(shift), mod_row_stride, _p(out), _p(out_lo), h16, l16, M,
                                       Cdim, rows_per_sample, float(eps), _stream()), "cvar_ln_modulate")
    return out if out is not None else out16


class SplitWeight:
    """A weight matrix with its tensor-core operand form (made once, consumed by TMA): the TF32 hi/lo split, or with
    ``f16=True`` the FP16 pair of the f16x3 engine (``.h16``; no TF32 copy is kept then)."""
    __slots__ = ("w", "hi", "lo", "h16")

    def __init__(self, w: torch.Tensor, f16: bool = False):
        self.w = w.contiguous()
        self.hi = self.lo = self.h16 = None
        if self.w.numel() % 4 != 0:
            raise _lib.CvarError("SplitWeight: number of elements must be a multiple of 4")
        if f16:
            self.h16 = F16Pair.from_tensor(self.w)
            return
        self.hi = torch.empty_like(self.w)
        self.lo = torch.empty_like(self.w)
        check(_lib.load().cvar_split_tf32(_p(self.w), _p(self.hi), _p(self.lo), self.w.numel(), _stream()),
              "cvar_split_tf32")


def _wparts(W):
    if isinstance(W, SplitWeight):
        return W.w, W.hi, W.lo, W.h16
    return W, None, None, None


def gemm(A, W, bias, out, M, N, K, *, lda=None, ldw=None, ldo=None, epilogue=EPI_BIAS, alpha=1.0, w_is_kn=False,
         batch=1, strideA=0, strideW=0, strideO=0, gamma=None, gamma_row_stride=0, rows_per_sample=1,
         resid=None, ldr=None, strideR=0, A_lo=None, out_lo=None, A16: Optional[F16Pair] = None,
         out16: Optional[F16Pair] = None):
    """A16 (with W a SplitWeight(f16=True)): FP16-pair operands, A may be None; out16: FP16-pair result, out may be None."""
    W, W_hi, W_lo, W16 = _wparts(W)
    _chk(A, W, bias, out, gamma, resid, A_lo, out_lo)
    if A16 is not None and W16 is None:
        raise _lib.CvarError("gemm: an FP16-pair activation needs a SplitWeight(f16=True) weight")
    a = GemmArgs()
    a.A16_hi, a.A16_lo = _p16(A16)
    a.W16_hi, a.W16_lo = _p16(W16 if A16 is not None else None)
    a.out16_hi, a.out16_lo = _p16(out16)
    a.W_hi, a.W_lo = _p(W_hi), _p(W_lo)
    a.A_lo, a.out_lo = _p(A_lo), _p(out_lo)
    a.A, a.lda, a.strideA = _p(A), (K if lda is None else lda), strideA
    a.W, a.ldw, a.strideW, a.w_is_kn = _p(W), ((N if w_is_kn else K) if ldw is None else ldw), strideW, int(w_is_kn)
    a.bias = _p(bias)
    a.out, a.ldo, a.strideO = _p(out), (N if ldo is None else ldo), strideO
    a.M, a.N, a.K, a.batch = M, N, K, batch
    a.epilogue, a.alpha = epilogue, float(alpha)
    a.gamma, a.gamma_row_stride, a.rows_per_sample = _p(gamma), gamma_row_stride, rows_per_sample
    a.resid, a.ldr, a.strideR = _p(resid), (N if ldr is None else ldr), strideR
    with _Timed("gemm", 2.0 * M * N * K * batch, 4.0 * batch * (M * K + N * K + M * N)):
        check(_lib.load().cvar_gemm(C.byref(a), _stream()), "cvar_gemm")
    return out if out is not None else out16


class KVCache:
    """KV arena of one transformer block in the operand format of the tensor-core attention kernel:
    K split hi/lo (R, H, T, 64) and V transposed + split (R, H, 64, T); T padded to a multiple of 4; zero-initialised
    once (stale tail keys are masked but must be finite).  Replaces the reference's torch.cat growth (basic_var.py:106-108)."""
    __slots__ = ("k_hi", "k_lo", "vt_hi", "vt_lo", "R", "H", "T")

    def __init__(self, R: int, H: int, T: int, device, storage: Optional[torch.Tensor] = None):
        self.R, self.H, self.T = R, H, (T + 3) // 4 * 4
        n = R * H * self.T * 64
        if storage is None:
            storage = torch.zeros(4 * n, dtype=torch.float32, device=device)
        self.k_hi, self.k_lo, self.vt_hi, self.vt_lo = (storage[i * n:(i + 1) * n] for i in range(4))

    @staticmethod
    def numel(R: int, H: int, T: int) -> int:
        return 4 * R * H * ((T + 3) // 4 * 4) * 64

    def keys(self, L: int) -> torch.Tensor:
        """(R, H, L, 64) view of the cached keys (hi + lo), for tests / debugging."""
        return (self.k_hi + self.k_lo).view(self.R, self.H, self.T, 64)[:, :, :L]

    def values(self, L: int) -> torch.Tensor:
        return (self.vt_hi + self.vt_lo).view(self.R, self.H, 64, self.T)[:, :, :, :L].transpose(2, 3)


class KVCache16:
    """KV arena of one transformer block as FP16 pairs (engine 4): K (R, H, T, 64) as qk pairs (16 k = hi + lo) and V^T
    (R, H, 64, T) as standard pairs (v = hi + lo / 2048), T padded to a multiple of 8; zero-initialised once.  Same bytes
    as one fp32 copy, half of KVCache's TF32 split."""
    __slots__ = ("k_hi", "k_lo", "vt_hi", "vt_lo", "R", "H", "T")

    def __init__(self, R: int, H: int, T: int, device, storage: Optional[torch.Tensor] = None):
        self.R, self.H, self.T = R, H, (T + 7) // 8 * 8
        n = R * H * self.T * 64
        if storage is None:
            storage = torch.zeros(4 * n, dtype=torch.float16, device=device)
        assert storage.dtype == torch.float16 and storage.numel() >= 4 * n
        self.k_hi, self.k_lo, self.vt_hi, self.vt_lo = (storage[i * n:(i + 1) * n] for i in range(4))

    @staticmethod
    def numel(R: int, H: int, T: int) -> int:
        return 4 * R * H * ((T + 7) // 8 * 8) * 64

    def keys(self, L: int) -> torch.Tensor:
        """(R, H, L, 64) fp32 view of the cached keys, for tests / debugging."""
        return ((self.k_hi.float() + self.k_lo.float()) / QK_SCALE).view(self.R, self.H, self.T, 64)[:, :, :L]

    def values(self, L: int) -> torch.Tensor:
        return (self.vt_hi.float() + self.vt_lo.float() / F16_LO_SCALE).view(self.R, self.H, 64, self.T)[:, :, :, :L] \
            .transpose(2, 3)


def qkv_project16(A16: F16Pair, Wqkv: "SplitWeight", q_bias, k_bias, v_bias, q16: F16Pair, cache: KVCache16, R, l, L_prev,
                  H, cos_attn, scale_mul_H):
    """cvar_qkv_project16: pair operands in, q / K / V^T out as FP16 pairs (q16: (R, H, l, 64))."""
    _chk(q_bias, k_bias, v_bias, scale_mul_H)
    if Wqkv.h16 is None:
        raise _lib.CvarError("qkv_project16 needs a SplitWeight(f16=True) weight")
    Cd = H * 64
    with _Timed("gemm", 2.0 * R * l * 3 * Cd * Cd, 4.0 * (R * l * Cd + 3 * Cd * Cd + R * l * 3 * Cd)):
        check(_lib.load().cvar_qkv_project16(_p(A16.hi), _p(A16.lo), _p(Wqkv.h16.hi), _p(Wqkv.h16.lo), _p(q_bias),
                                             _p(k_bias), _p(v_bias), _p(q16.hi), _p(q16.lo), _p(cache.k_hi),
                                             _p(cache.k_lo), _p(cache.vt_hi), _p(cache.vt_lo), R, l, L_prev, cache.T, H,
                                             int(cos_attn), _p(scale_mul_H), _stream()), "cvar_qkv_project16")


def attn_kvcache16(q16: F16Pair, cache: KVCache16, out, R, H, l, L, scale, engine: int = -1,
                   out16: Optional[F16Pair] = None):
    _chk(out)
    o16h, o16l = _p16(out16)
    with _Timed("attn", 4.0 * l * L * 64 * R * H, (2.0 * l + 2.0 * L) * 64 * 4 * R * H):
        check(_lib.load().cvar_attn_kvcache16(_p(q16.hi), _p(q16.lo), _p(cache.k_hi), _p(cache.k_lo), _p(cache.vt_hi),
                                              _p(cache.vt_lo), _p(out), o16h, o16l, R, H, l, L, cache.T, float(scale),
                                              int(engine), _stream()), "cvar_attn_kvcache16")
    return out if out is not None else out16


def attn_blockcausal16(q16: F16Pair, cache: KVCache16, out, R, H, scale_lens, scale, out16: Optional[F16Pair] = None):
    """Block-causal attention over the whole pyramid in one launch (ControlVAR.forward); scale_lens: tokens per scale."""
    _chk(out)
    o16h, o16l = _p16(out16)
    lens = [int(x) for x in scale_lens]
    arr = (C.c_int * len(lens))(*lens)
    l_total = sum(lens)
    flops, L = 0.0, 0
    for ls in lens:
        L += ls
        flops += 4.0 * ls * L * 64
    with _Timed("attn", flops * R * H, (2.0 * l_total + 2.0 * l_total) * 64 * 4 * R * H):
        check(_lib.load().cvar_attn_blockcausal16(_p(q16.hi), _p(q16.lo), _p(cache.k_hi), _p(cache.k_lo), _p(cache.vt_hi),
                                                  _p(cache.vt_lo), _p(out), o16h, o16l, R, H, l_total, cache.T, float(scale),
                                                  len(lens), arr, _stream()), "cvar_attn_blockcausal16")
    return out if out is not None else out16


def qkv_project(A, Wqkv, q_bias, k_bias, v_bias, q_out, cache: KVCache, R, l, L_prev, H, cos_attn, scale_mul_H,
                A_lo=None, A16: Optional[F16Pair] = None):
    Wqkv, W_hi, W_lo, W16 = _wparts(Wqkv)
    _chk(A, Wqkv, q_bias, k_bias, v_bias, q_out, cache.k_hi, scale_mul_H, A_lo)
    if A16 is not None and W16 is None:
        raise _lib.CvarError("qkv_project: an FP16-pair activation needs a SplitWeight(f16=True) weight")
    a16h, a16l = _p16(A16)
    w16h, w16l = _p16(W16 if A16 is not None else None)
    Cd = H * 64
    with _Timed("gemm", 2.0 * R * l * 3 * Cd * Cd, 4.0 * (R * l * Cd + 3 * Cd * Cd + R * l * 3 * Cd)):
        check(_lib.load().cvar_qkv_project(_p(A), _p(A_lo), a16h, a16l, _p(Wqkv), _p(W_hi), _p(W_lo), w16h, w16l,
                                           _p(q_bias), _p(k_bias), _p(v_bias),
                                           _p(q_out), _p(cache.k_hi), _p(cache.k_lo), _p(cache.vt_hi), _p(cache.vt_lo),
                                           R, l, L_prev, cache.T, H, int(cos_attn), _p(scale_mul_H), _stream()),
              "cvar_qkv_project")


def attn_kvcache(q, cache: KVCache, out, R, H, l, L, scale, engine: int = -1, out_lo=None,
                 out16: Optional[F16Pair] = None):
    _chk(q, cache.k_hi, out, out_lo)
    o16h, o16l = _p16(out16)
    # algorithmic work of SURVEY.md section 8d: 4*l*L*64 flop and (2l + 2L)*64*4 bytes per (row, head)
    with _Timed("attn", 4.0 * l * L * 64 * R * H, (2.0 * l + 2.0 * L) * 64 * 4 * R * H):
        check(_lib.load().cvar_attn_kvcache(_p(q), _p(cache.k_hi), _p(cache.k_lo), _p(cache.vt_hi), _p(cache.vt_lo),
                                            _p(out), _p(out_lo), o16h, o16l, R, H, l, L, cache.T, float(scale),
                                            int(engine), _stream()),
              "cvar_attn_kvcache")
    return out if out is not None else out16


def cfg_sample(logits, q_noise, idx_out, B, l, V, t, top_k, top_p):
    _chk(logits, q_noise, idx_out)
    with _Timed("sample", 0.0, 4.0 * B * l * V * 3 + 8.0 * B * l):
        check(_lib.load().cvar_cfg_sample(_p(logits), _p(q_noise), _p(idx_out), B, l, V, float(t), int(top_k),
                                          float(top_p), _stream()), "cvar_cfg_sample")
    return idx_out


def cfg_sample_multi(logits, q_noise, idx_out, B, l, V, coef, replicas, top_k, top_p, forced_first=None,
                     forced_second=None, forced_replicas=0):
    """Guidance mix over len(coef) logit groups, `replicas` independent draws per row, optional teacher forcing
    (control_var.py:288-321).  coef: python floats, already rounded the way the reference's scalars are."""
    _chk(logits, q_noise, idx_out, forced_first, forced_second)
    G = len(coef)
    arr = (C.c_float * G)(*[float(c) for c in coef])
    with _Timed("sample", 0.0, 4.0 * B * l * V * (G + replicas) + 8.0 * B * l * replicas):
        check(_lib.load().cvar_cfg_sample_multi(_p(logits), _p(q_noise), _p(idx_out), B, l, V, G, arr, int(replicas),
                                                int(top_k), float(top_p), _p(forced_first), _p(forced_second),
                                                int(forced_replicas), _stream()), "cvar_cfg_sample_multi")
    return idx_out


def cfg_sample_masked(logits, q_noise, idx_out, masked_out, B, l, V, coef, replicas, top_k, top_p, forced_first=None,
                      forced_second=None, forced_replicas=0):
    """cfg_sample_multi that also writes the mixed logits as sample_with_top_k_top_p_ leaves them (masked_out: (B*l, V),
    removed entries -inf) - the input of gumbel_embed (more_smooth, control_var.py:513-515)."""
    _chk(logits, q_noise, idx_out, masked_out, forced_first, forced_second)
    G = len(coef)
    arr = (C.c_float * G)(*[float(c) for c in coef])
    with _Timed("sample", 0.0, 4.0 * B * l * V * (G + replicas + 1) + 8.0 * B * l * replicas):
        check(_lib.load().cvar_cfg_sample_masked(_p(logits), _p(q_noise), _p(idx_out), _p(masked_out), B, l, V, G, arr,
                                                 int(replicas), int(top_k), float(top_p), _p(forced_first), _p(forced_second),
                                                 int(forced_replicas), _stream()), "cvar_cfg_sample_masked")
    return idx_out


def gumbel_embed(masked_logits, e_noise, embedding, h_out, rows_in, rows_out, V, Cvae, mul, tau):
    """h = softmax((masked_logits * mul + (-log e_noise)) / tau) @ embedding (helpers.py:22-36 with hard=False); output row r
    reads logits row r % rows_in."""
    _chk(masked_logits, e_noise, embedding, h_out)
    with _Timed("sample", 2.0 * rows_out * V * Cvae, 4.0 * V * (rows_in + rows_out) + 4.0 * rows_out * Cvae):
        check(_lib.load().cvar_gumbel_embed(_p(masked_logits), _p(e_noise), _p(embedding), _p(h_out), int(rows_in),
                                            int(rows_out), V, Cvae, float(mul), float(tau), _stream()), "cvar_gumbel_embed")
    return h_out


def vq_step(idx, embedding, U, phi_w, phi_b, word_w, word_b, lvl_pos_next, f_hat, x_next, B, pn, pn_next, hw, Cvae, Cdim,
            streams=2, x_replicas=2, f_rest=None):
    """streams / x_replicas / f_rest: see cvar_vq_step_ex (defaults = the autoregressive_infer_cfg step)."""
    _chk(idx, embedding, U, phi_w, phi_b, word_w, word_b, lvl_pos_next, f_hat, x_next, f_rest)
    check(_lib.load().cvar_vq_step_ex(_p(idx), _p(embedding), _p(U), _p(phi_w), _p(phi_b), _p(word_w), _p(word_b),
                                      _p(lvl_pos_next), _p(f_hat), _p(f_rest), _p(x_next), B, int(streams),
                                      int(x_replicas), pn, pn_next, hw, Cvae, Cdim, _stream()), "cvar_vq_step_ex")


def area_pool_nc(f_nchw, z_NC, B, Cvae, hw, pn):
    _chk(f_nchw, z_NC)
    check(_lib.load().cvar_area_pool_nc(_p(f_nchw), _p(z_NC), B, Cvae, hw, pn, _stream()), "cvar_area_pool_nc")
    return z_NC


def vq_nearest(z_NC, embedding, idx_out):
    _chk(z_NC, embedding, idx_out)
    N, Cv = z_NC.shape
    check(_lib.load().cvar_vq_nearest(_p(z_NC), _p(embedding), _p(idx_out), N, Cv, embedding.shape[0], _stream()),
          "cvar_vq_nearest")
    return idx_out


def nchw_to_nhwc(x_view, out, B, Cdim, H, W, in_batch_stride):
    _chk(x_view, out)
    check(_lib.load().cvar_nchw_to_nhwc(_p(x_view), _p(out), B, Cdim, H, W, in_batch_stride, _stream()),
          "cvar_nchw_to_nhwc")
    return out


def nchw_to_nhwc_pad(x, out, B, Cdim, H, W, Cpad):
    _chk(x, out)
    check(_lib.load().cvar_nchw_to_nhwc_pad(_p(x), _p(out), B, Cdim, H, W, Cpad, _stream()), "cvar_nchw_to_nhwc_pad")
    return out


def gn_chunks(HW: int) -> int:
    return int(_lib.load().cvar_gn_chunks(HW))


def gn_stats(x_nhwc, gamma, beta, a_out, b_out, scratch, B, HW, Cdim, groups=32, eps=1e-6):
    _chk(x_nhwc, gamma, beta, a_out, b_out, scratch)
    check(_lib.load().cvar_gn_stats(_p(x_nhwc), _p(gamma), _p(beta), _p(a_out), _p(b_out), _p(scratch), B, HW, Cdim,
                                    groups, float(eps), _stream()), "cvar_gn_stats")


def conv2d(x, w_packed, bias, out, B, Hin, Win, Cin, Cout, ks, *, in_a=None, in_b=None, in_silu=False, resid=None,
           upsample2x=False, out_mode=0, out_rows_total=0, row_offset=0, engine=-1, x16: Optional[F16Pair] = None,
           w16: Optional[F16Pair] = None, downsample2x=False, ksplit=0, out_samples=0, gn_part=None, gn_groups=32):
    """x16 + w16: FP16-pair input (already normalised / upsampled) and weight -> the 2-CTA TMA kernel; x may be None.
    downsample2x: the encoder's pad-(0,1,0,1) + stride-2 convolution (vae_modules.py:31-37)."""
    w_packed, w_hi, w_lo, _ = _wparts(w_packed)
    _chk(x, w_packed, bias, out, in_a, in_b, resid)
    if (x16 is None) != (w16 is None):
        raise _lib.CvarError("conv2d: x16 and w16 must come together")
    a = ConvArgs()
    a.x16_hi, a.x16_lo = _p16(x16)
    a.w16_hi, a.w16_lo = _p16(w16)
    a.x, a.w, a.bias, a.out = _p(x), _p(w_packed), _p(bias), _p(out)
    a.w_hi, a.w_lo = _p(w_hi), _p(w_lo)
    a.in_a, a.in_b, a.in_silu = _p(in_a), _p(in_b), int(in_silu)
    a.resid = _p(resid)
    a.B, a.Hin, a.Win, a.Cin, a.Cout, a.ks, a.upsample2x = B, Hin, Win, Cin, Cout, ks, int(upsample2x)
    a.out_mode, a.out_rows_total, a.row_offset = out_mode, out_rows_total, row_offset
    a.engine = int(engine)
    a.downsample2x = int(downsample2x)
    a.ksplit = int(ksplit)
    a.out_samples = int(out_samples)
    a.gn_part, a.gn_groups = _p(gn_part), (int(gn_groups) if gn_part is not None else 0)
    up = 2 if upsample2x else 1
    Mo = B * Hin * up * Win * up // (4 if downsample2x else 1)
    with _Timed("conv", 2.0 * Mo * Cout * ks * ks * Cin, 4.0 * (B * Hin * Win * Cin + Mo * Cout + Cout * ks * ks * Cin)):
        check(_lib.load().cvar_conv2d(C.byref(a), _stream()), "cvar_conv2d")
    return out


def repack_conv_weight(w_oihw, out):
    _chk(w_oihw, out)
    Cout, Cin, ks, _ = w_oihw.shape
    check(_lib.load().cvar_repack_conv_weight(_p(w_oihw), _p(out), Cout, Cin, ks, _stream()), "cvar_repack_conv_weight")
    return out


def repack_conv_weight_pad(w_oihw, out, Cin_pad):
    _chk(w_oihw, out)
    Cout, Cin, ks, _ = w_oihw.shape
    check(_lib.load().cvar_repack_conv_weight_pad(_p(w_oihw), _p(out), Cout, Cin, ks, int(Cin_pad), _stream()),
          "cvar_repack_conv_weight_pad")
    return out


def affine_nc(x, a, b, out, B, HW, Cdim, silu=False, out16: Optional[F16Pair] = None):
    _chk(x, a, b, out)
    h16, l16 = _p16(out16)
    check(_lib.load().cvar_affine_nc(_p(x), _p(a), _p(b), _p(out), h16, l16, B, HW, Cdim, int(silu), _stream()),
          "cvar_affine_nc")
    return out if out is not None else out16


def conv2d_gn_fusable(H, W, Cin, Cout, ks, groups=32) -> bool:
    return bool(_lib.load().cvar_conv2d_gn_fusable(int(H), int(W), int(Cin), int(Cout), int(ks), int(groups)))


def gn_finalize_parts(gn_part, gamma, beta, a_out, b_out, B, HW, Cdim, groups=32, eps=1e-6):
    _chk(gn_part, gamma, beta, a_out, b_out)
    check(_lib.load().cvar_gn_finalize_parts(_p(gn_part), _p(gamma), _p(beta), _p(a_out), _p(b_out), B, HW, Cdim, groups,
                                             float(eps), _stream()), "cvar_gn_finalize_parts")


def conv2d_f16_supported(H, W, Cin, Cout, ks) -> bool:
    return bool(_lib.load().cvar_conv2d_f16_supported(int(H), int(W), int(Cin), int(Cout), int(ks)))


def upsample2x_split_f16(x, out16: F16Pair, B, H, W, Cdim):
    _chk(x)
    check(_lib.load().cvar_upsample2x_split_f16(_p(x), _p(out16.hi), _p(out16.lo), B, H, W, Cdim, _stream()),
          "cvar_upsample2x_split_f16")
    return out16


def softmax_rows(x, rows, cols):
    _chk(x)
    check(_lib.load().cvar_softmax_rows(_p(x), rows, cols, _stream()), "cvar_softmax_rows")
    return x
