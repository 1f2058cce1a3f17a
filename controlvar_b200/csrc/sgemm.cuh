// SIMT fp32 GEMM core (FFMA, round-to-nearest accumulate): out = epilogue(A[M,K] * B[N,K]^T).
// This is the exact-arithmetic engine of the library: it serves the shapes the tcgen05 engine does not take
// (tiny M, K = 32, odd N) and is the fp32 yard-stick the tensor-core engine is validated against on the GPU.
// Operand access and the epilogue are functors, so dense layers, the fused QKV/KV-cache append and the
// implicit-GEMM decoder convolutions share one main loop.
#pragma once
#include "common.cuh"

namespace cvar {

// ---------------------------------------------------------------------------------------------- loaders
// A loader contract:   void prep(int slot, long long m, int batch);  float4 fetch(int slot, int k) (k % 4 == 0)
struct DenseALoader {
  const float* A;
  long long lda, strideA;
  int M, K;
  int vec;
  const float* A_lo = nullptr;   // optional: A holds the TF32 'hi' part and A_lo the remainder (hi + lo == x exactly)   // 1: K, lda multiples of 4 and 16-byte aligned base -> 128-bit loads; 0: scalar loads (ragged shapes)
  static constexpr int kMaxSlots = 8;
  const float* ptr[kMaxSlots];
  __device__ __forceinline__ void prep(int slot, long long m, int batch) {
    ptr[slot] = (m < M) ? A + (long long)batch * strideA + m * lda : nullptr;
  }
  __device__ __forceinline__ float4 fetch(int slot, int k) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* p = ptr[slot];
    if (p == nullptr || k >= K) return v;
    if (vec) {
      v = ld4(p + k);
      if (A_lo != nullptr) {
        float4 w = ld4(A_lo + (p - A) + k);
        v.x += w.x, v.y += w.y, v.z += w.z, v.w += w.w;
      }
      return v;
    }
    v.x = p[k];
    if (k + 1 < K) v.y = p[k + 1];
    if (k + 2 < K) v.z = p[k + 2];
    if (k + 3 < K) v.w = p[k + 3];
    return v;
  }
  // tcgen05 engine: k0 = first column of the K-block (multiple of the block size), koff = offset inside the block
  __device__ __forceinline__ void begin_block(int /*k0*/) {}
  __device__ __forceinline__ float4 fetch_blk(int slot, int k0, int koff) const { return fetch(slot, k0 + koff); }
};

// Implicit-GEMM view of a stride-1 'same' convolution over an NHWC activation (vae_modules.py Conv2d call sites):
// row m = (n, y, x) of the OUTPUT grid, column k = (tap, ci).  Optional nearest x2 upsampling of the input
// (Upsample2x, vae_modules.py:27-28) and optional fused GroupNorm-affine (+SiLU) on the input (vae_modules.py:58-59).
// stride = 2 with pad = 0 is the encoder's Downsample2x (vae_modules.py:31-37): F.pad(x, (0,1,0,1)) + a stride-2 'valid'
// 3x3 conv, i.e. output (y, x) reads input rows 2y..2y+2, zero beyond the bottom / right edge only.
struct ConvALoader {
  const float* x;
  const float* in_a;
  const float* in_b;
  int in_silu;
  int Hin, Win, Cin, ks, up;      // up: 0 or 1 (log2 of the upsampling factor)
  int Hout, Wout;                 // output grid (row decode)
  int Hv, Wv;                     // input grid the taps are bounds-checked against (Hin << up, Win << up)
  int stride, pad;                // 1, ks/2 for the 'same' convolutions; 2, 0 for Downsample2x
  long long Mtot;
  int K;
  static constexpr int kMaxSlots = 8;
  int sn[kMaxSlots], sy[kMaxSlots], sx[kMaxSlots];   // sy / sx hold the top-left tap position (y*stride - pad)
  int blk_dy, blk_dx, blk_ci;   // tap offset and first input channel of the current K-block (tcgen05 engine)
  // A K-block never straddles two taps when Cin is a multiple of the block size, so the tap decode (two integer
  // divisions) is done once per block instead of once per 16-byte fetch.
  __device__ __forceinline__ void begin_block(int k0) {
    int tap = k0 / Cin;
    blk_ci = k0 - tap * Cin;
    int ky = tap / ks;
    blk_dy = ky;
    blk_dx = tap - ky * ks;
  }
  __device__ __forceinline__ float4 fetch_blk(int slot, int /*k0*/, int koff) const {
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sn[slot] < 0) return z;
    int yy = sy[slot] + blk_dy, xx = sx[slot] + blk_dx;
    if (yy < 0 || yy >= Hv || xx < 0 || xx >= Wv) return z;
    int ci = blk_ci + koff;
    int n = sn[slot];
    float4 v = ld4(x + (((long long)n * Hin + (yy >> up)) * Win + (xx >> up)) * Cin + ci);
    if (in_a != nullptr) {
      float4 a = ld4(in_a + (long long)n * Cin + ci), b = ld4(in_b + (long long)n * Cin + ci);
      v.x = fmaf(v.x, a.x, b.x), v.y = fmaf(v.y, a.y, b.y), v.z = fmaf(v.z, a.z, b.z), v.w = fmaf(v.w, a.w, b.w);
      if (in_silu) v.x = silu_f(v.x), v.y = silu_f(v.y), v.z = silu_f(v.z), v.w = silu_f(v.w);
    }
    return v;
  }
  __device__ __forceinline__ void prep(int slot, long long m, int /*batch*/) {
    if (m < Mtot) {
      int xw = (int)(m % Wout);
      long long q = m / Wout;
      sy[slot] = (int)(q % Hout) * stride - pad;
      sn[slot] = (int)(q / Hout);
      sx[slot] = xw * stride - pad;
    } else {
      sn[slot] = -1;
    }
  }
  __device__ __forceinline__ float4 fetch(int slot, int k) const {
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sn[slot] < 0 || k >= K) return z;
    int tap = k / Cin;
    int ci = k - tap * Cin;
    int ky = tap / ks, kx = tap - ky * ks;
    int yy = sy[slot] + ky, xx = sx[slot] + kx;
    if (yy < 0 || yy >= Hv || xx < 0 || xx >= Wv) return z;          // zero padding is applied AFTER norm+act
    int n = sn[slot];
    const float* p = x + (((long long)n * Hin + (yy >> up)) * Win + (xx >> up)) * Cin + ci;
    float4 v = ld4(p);
    if (in_a != nullptr) {
      float4 a = ld4(in_a + (long long)n * Cin + ci), b = ld4(in_b + (long long)n * Cin + ci);
      v.x = fmaf(v.x, a.x, b.x);
      v.y = fmaf(v.y, a.y, b.y);
      v.z = fmaf(v.z, a.z, b.z);
      v.w = fmaf(v.w, a.w, b.w);
      if (in_silu) {
        v.x = silu_f(v.x);
        v.y = silu_f(v.y);
        v.z = silu_f(v.z);
        v.w = silu_f(v.w);
      }
    }
    return v;
  }
};

// B operand: W[N,K] row-major (nn.Linear / repacked conv weight), or W[K,N] row-major when kn != 0.
struct DenseBLoader {
  const float* W;
  long long ldw, strideW;
  int N, K, kn;
  int vec;   // as DenseALoader::vec
};

// -------------------------------------------------------------------------------------------- epilogues
// Epilogue contract: void store(long long m, int n, const float v[4], int nvalid, int batch)
struct DenseEpilogue {
  float* out;
  long long ldo, strideO;
  const float* bias;
  int mode;
  float alpha;
  const float* gamma;
  long long gamma_row_stride;
  int rows_per_sample;
  const float* resid;
  long long ldr, strideR;
  float* out_lo = nullptr;       // optional (BIAS / BIAS_GELU): write the result split, out = hi, out_lo = lo
  __half* out16_hi = nullptr;    // optional (BIAS / BIAS_GELU): write the result as an FP16 pair; `out` may then be null
  __half* out16_lo = nullptr;
  __device__ __forceinline__ void store(long long m, int n, const float* v, int nvalid, int batch) const {
    float* o = out + (long long)batch * strideO + m * ldo + n;
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float b = (bias != nullptr && j < nvalid) ? bias[n + j] : 0.f;
      float t = (mode == CVAR_EPI_BIAS) ? __fadd_rn(__fmul_rn(v[j], alpha), b) : __fadd_rn(v[j], b);
      if (mode == CVAR_EPI_BIAS_GELU) t = gelu_tanh_f(t);
      r[j] = t;
    }
    if (mode == CVAR_EPI_BIAS_GAMMA_RESID) {
      const float* g = gamma + (m / rows_per_sample) * gamma_row_stride + n;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nvalid) r[j] = __fadd_rn(o[j], __fmul_rn(r[j], g[j]));      // x + branch.mul(gamma)
    } else if (mode == CVAR_EPI_BIAS_RESID) {
      const float* rs = resid + (long long)batch * strideR + m * ldr + n;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nvalid) r[j] = __fadd_rn(rs[j], r[j]);                      // shortcut + h
    }
    if (out16_hi != nullptr) {
      const long long off = (long long)batch * strideO + m * ldo + n;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nvalid) split_f16(r[j], out16_hi[off + j], out16_lo[off + j]);
      if (out == nullptr) return;
    }
    if (out_lo != nullptr) {
      float* ol = out_lo + (long long)batch * strideO + m * ldo + n;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nvalid) {
          float hi = __uint_as_float(__float_as_uint(r[j]) & 0xFFFFE000u);
          ol[j] = r[j] - hi;
          r[j] = hi;
        }
    }
    if (nvalid == 4 && ((((uintptr_t)o) & 15) == 0)) {
      st4(o, make_float4(r[0], r[1], r[2], r[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nvalid) o[j] = r[j];
    }
  }
};

// ---- two-phase form used by the tcgen05 epilogue: all global READS of a pass are issued first (load_aux), then the
// math and the stores (store_aux), so the loads of independent rows overlap instead of serialising behind stores.
struct EpiAux {
  float4 a, b;
};

__device__ __forceinline__ EpiAux dense_load_aux(const DenseEpilogue& e, long long m, int n, int nvalid, int batch) {
  EpiAux x;
  x.a = make_float4(0.f, 0.f, 0.f, 0.f);
  x.b = x.a;
  if (nvalid != 4) return x;            // ragged tail: store_aux falls back to the scalar path
  if (e.mode == CVAR_EPI_BIAS_GAMMA_RESID) {
    x.a = ld4(e.out + (long long)batch * e.strideO + m * e.ldo + n);
    x.b = ld4(e.gamma + (m / e.rows_per_sample) * e.gamma_row_stride + n);
  } else if (e.mode == CVAR_EPI_BIAS_RESID) {
    x.a = ld4(e.resid + (long long)batch * e.strideR + m * e.ldr + n);
  }
  return x;
}
// fast, fp32-class GELU(tanh) for the tensor-core epilogue: x * sigmoid(2u) with ex2.approx / rcp.approx (<= ~1e-6 rel)
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return __fdividef(x, 1.0f + __expf(-2.0f * u));
}
__device__ __forceinline__ void dense_store_aux(const DenseEpilogue& e, long long m, int n, const float* v, int nvalid,
                                                int batch, const EpiAux& x) {
  const long long ooff = (long long)batch * e.strideO + m * e.ldo + n;
  float* o = e.out + ooff;
  if (nvalid != 4 || ((((uintptr_t)o) & 15) != 0) || (ooff & 3) != 0) {
    e.store(m, n, v, nvalid, batch);
    return;
  }
  float4 b4 = e.bias != nullptr ? ld4(e.bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
  float r[4] = {v[0], v[1], v[2], v[3]};
  const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float t = (e.mode == CVAR_EPI_BIAS) ? __fadd_rn(__fmul_rn(r[j], e.alpha), bb[j]) : __fadd_rn(r[j], bb[j]);
    if (e.mode == CVAR_EPI_BIAS_GELU) t = gelu_tanh_fast(t);
    r[j] = t;
  }
  if (e.mode == CVAR_EPI_BIAS_GAMMA_RESID) {
    r[0] = __fadd_rn(x.a.x, __fmul_rn(r[0], x.b.x)), r[1] = __fadd_rn(x.a.y, __fmul_rn(r[1], x.b.y));
    r[2] = __fadd_rn(x.a.z, __fmul_rn(r[2], x.b.z)), r[3] = __fadd_rn(x.a.w, __fmul_rn(r[3], x.b.w));
  } else if (e.mode == CVAR_EPI_BIAS_RESID) {
    r[0] = __fadd_rn(x.a.x, r[0]), r[1] = __fadd_rn(x.a.y, r[1]), r[2] = __fadd_rn(x.a.z, r[2]), r[3] = __fadd_rn(x.a.w, r[3]);
  }
  if (e.out16_hi != nullptr) {
    st4_split_f16(e.out16_hi + ooff, e.out16_lo + ooff, r);
    if (e.out == nullptr) return;
  }
  if (e.out_lo != nullptr) {
    float l4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float hi = __uint_as_float(__float_as_uint(r[j]) & 0xFFFFE000u);
      l4[j] = r[j] - hi;
      r[j] = hi;
    }
    st4(e.out_lo + (long long)batch * e.strideO + m * e.ldo + n, make_float4(l4[0], l4[1], l4[2], l4[3]));
  }
  st4(o, make_float4(r[0], r[1], r[2], r[3]));
}

// qkv = x W^T + [q_bias, k_bias, v_bias]; q to (R,H,l,64), k/v appended to the cache      (basic_var.py:92-108)
struct QkvEpilogue {
  const float* q_bias;
  const float* k_bias;
  const float* v_bias;
  float* q_out;
  // KV cache in the operand format of the tensor-core attention kernel (split once here, consumed by TMA there):
  //   K   : k_hi / k_lo   [R*H][T_max][64]   (TF32 hi/lo split, hi + lo == k exactly)
  //   V^T : vt_hi / vt_lo [R*H][64][T_max]   (transposed so that keys are the contiguous, K-major dimension of P @ V)
  float* k_hi;
  float* k_lo;
  float* vt_hi;
  float* vt_lo;
  int C, H, l, L_prev, T_max;
  // FP16-pair form of the same three outputs (cvar_qkv_project16): when q16_hi is set, q / K / V^T are written as
  // pairs with the same index layout and the fp32 pointers above are ignored: V^T as a standard pair (hi + lo * 2^-11,
  // cvar_split_f16), q and K as "qk pairs" (16 x = hi + lo, split_f16_qk in common.cuh).
  // This is the operand format of the f16 tensor-core attention kernel (q, K, V^T tiles fetched by TMA as they are).
  __half* q16_hi = nullptr;
  __half* q16_lo = nullptr;
  __half* k16_hi = nullptr;
  __half* k16_lo = nullptr;
  __half* vt16_hi = nullptr;
  __half* vt16_lo = nullptr;
  __device__ __forceinline__ void store(long long m, int n, const float* v, int nvalid, int /*batch*/) const {
    int which = n / C;
    int c = n - which * C;
    int h = c >> 6, d = c & 63;
    int r = (int)(m / l), t = (int)(m - (long long)r * l);
    const float* bias = which == 0 ? q_bias : (which == 1 ? k_bias : v_bias);
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = __fadd_rn(v[j], bias[c + j]);
    const long long rh = (long long)r * H + h;
    if (q16_hi != nullptr) {
      if (which == 0) {
        const long long off = ((rh * l + t) << 6) + d;
        st4_split_f16_qk(q16_hi + off, q16_lo + off, o);
      } else if (which == 1) {
        const long long off = ((rh * T_max + L_prev + t) << 6) + d;
        st4_split_f16_qk(k16_hi + off, k16_lo + off, o);
      } else {
        const long long off = (rh * 64 + d) * T_max + L_prev + t;
#pragma unroll
        for (int j = 0; j < 4; ++j) split_f16(o[j], vt16_hi[off + (long long)j * T_max], vt16_lo[off + (long long)j * T_max]);
      }
      return;
    }
    if (which == 0) {
      st4(q_out + ((rh * l + t) << 6) + d, make_float4(o[0], o[1], o[2], o[3]));
    } else {
      float hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hi[j] = __uint_as_float(__float_as_uint(o[j]) & 0xFFFFE000u);
        lo[j] = o[j] - hi[j];
      }
      if (which == 1) {
        const long long off = ((rh * T_max + L_prev + t) << 6) + d;
        st4(k_hi + off, make_float4(hi[0], hi[1], hi[2], hi[3]));
        st4(k_lo + off, make_float4(lo[0], lo[1], lo[2], lo[3]));
      } else {
        const long long off = (rh * 64 + d) * T_max + L_prev + t;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          vt_hi[off + (long long)j * T_max] = hi[j];
          vt_lo[off + (long long)j * T_max] = lo[j];
        }
      }
    }
    (void)nvalid;
  }
};

// conv output: NHWC (+bias, +residual) or the final image plane write                    (vqvae.py:88-89)
// out_mode 3: NCHW planes without the clamp (quant_conv output feeding the residual quantiser, vqvae.py:74)
struct ConvEpilogue {
  float* out;
  const float* bias;
  const float* resid;
  int Cout, out_mode;
  int Hout, Wout, out_rows_total, row_offset;
  int accumulate = 0;     // K-split continuation (out_mode 0): out += acc, no bias, no residual
  // image output (out_mode != 0) of a STACKED batch: image n = s * out_samples + b of the batch lands in sample b of the
  // (out_samples, Cout, out_rows_total, Wout) tensor at rows s * Hout + row_offset (the control / image halves of
  // control_var.py:563-565 decoded in ONE pass: s = 0 control on top, s = 1 image below).  0: image n is sample n.
  int out_samples = 0;
  // GroupNorm statistics of the OUTPUT from the epilogue (row epilogue of the 2-CTA kernel only; cvar_conv_args.gn_part):
  // every epilogue warp writes (sum, sum of squares) of its 32 pixels x one group as two doubles to
  // gn_part[((image * gn_groups + group) * gn_slots + slot) * 2], slot = (pixel index in the image) / 32.
  double* gn_part = nullptr;
  int gn_groups = 0, gn_cpg = 0, gn_slots = 0, gn_HW = 0;
  __device__ __forceinline__ void store(long long m, int n, const float* v, int nvalid, int /*batch*/) const {
    if (out_mode == 0) {
      float* o = out + m * Cout + n;
      float r[4];
      if (accumulate) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nvalid) o[j] = __fadd_rn(o[j], v[j]);
        return;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = __fadd_rn(v[j], (j < nvalid) ? bias[n + j] : 0.f);
      if (resid != nullptr) {
        const float* rs = resid + m * Cout + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nvalid) r[j] = __fadd_rn(rs[j], r[j]);
      }
      if (nvalid == 4) {
        st4(o, make_float4(r[0], r[1], r[2], r[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nvalid) o[j] = r[j];
      }
    } else {
      int xw = (int)(m % Wout);
      long long q = m / Wout;
      int y = (int)(q % Hout);
      long long nimg = q / Hout;
      int roff = row_offset;
      if (out_samples > 0) {
        roff += (int)(nimg / out_samples) * Hout;
        nimg = nimg % out_samples;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nvalid) {
          float t = __fadd_rn(v[j], bias[n + j]);
          if (out_mode != 3) t = fminf(fmaxf(t, -1.f), 1.f);  // .clamp_(-1, 1)           vqvae.py:89
          if (out_mode == 1) t = __fmul_rn(__fadd_rn(t, 1.f), 0.5f);   // .add_(1).mul_(0.5)  control_var.py:563
          out[((nimg * Cout + (n + j)) * out_rows_total + roff + y) * Wout + xw] = t;
        }
      }
    }
  }
};

__device__ __forceinline__ EpiAux epi_load_aux(const DenseEpilogue& e, long long m, int n, int nv, int b) {
  return dense_load_aux(e, m, n, nv, b);
}
__device__ __forceinline__ void epi_store_aux(const DenseEpilogue& e, long long m, int n, const float* v, int nv, int b,
                                              const EpiAux& x) {
  dense_store_aux(e, m, n, v, nv, b, x);
}
__device__ __forceinline__ EpiAux epi_load_aux(const QkvEpilogue&, long long, int, int, int) {
  EpiAux x;
  x.a = make_float4(0.f, 0.f, 0.f, 0.f);
  x.b = x.a;
  return x;
}
__device__ __forceinline__ void epi_store_aux(const QkvEpilogue& e, long long m, int n, const float* v, int nv, int b,
                                              const EpiAux&) {
  e.store(m, n, v, nv, b);
}
__device__ __forceinline__ EpiAux epi_load_aux(const ConvEpilogue& e, long long m, int n, int nv, int) {
  EpiAux x;
  x.a = make_float4(0.f, 0.f, 0.f, 0.f);
  x.b = x.a;
  if (e.out_mode == 0 && e.accumulate && nv == 4) x.a = ld4(e.out + m * e.Cout + n);
  else if (e.out_mode == 0 && e.resid != nullptr && nv == 4) x.a = ld4(e.resid + m * e.Cout + n);
  return x;
}
__device__ __forceinline__ void epi_store_aux(const ConvEpilogue& e, long long m, int n, const float* v, int nv, int b,
                                              const EpiAux& x) {
  if (e.out_mode != 0 || nv != 4) {
    e.store(m, n, v, nv, b);
    return;
  }
  if (e.accumulate) {
    st4(e.out + m * e.Cout + n, make_float4(__fadd_rn(x.a.x, v[0]), __fadd_rn(x.a.y, v[1]), __fadd_rn(x.a.z, v[2]),
                                            __fadd_rn(x.a.w, v[3])));
    return;
  }
  float4 b4 = ld4(e.bias + n);
  float4 r = make_float4(__fadd_rn(v[0], b4.x), __fadd_rn(v[1], b4.y), __fadd_rn(v[2], b4.z), __fadd_rn(v[3], b4.w));
  if (e.resid != nullptr)
    r = make_float4(__fadd_rn(x.a.x, r.x), __fadd_rn(x.a.y, r.y), __fadd_rn(x.a.z, r.z), __fadd_rn(x.a.w, r.w));
  st4(e.out + m * e.Cout + n, r);
}

// ------------------------------------------------------------------------------------------ main kernel
template <int BM, int BN, int FM, int FN, class AL, class EP>
__global__ void __launch_bounds__((BM / (4 * FM)) * (BN / (4 * FN)))
sgemm_kernel(AL al, DenseBLoader bl, EP ep, long long M, int N, int K) {
  constexpr int BK = 16;
  constexpr int TX = BN / (4 * FN), TY = BM / (4 * FM), NT = TX * TY;
  constexpr int SA = (BM * BK / 4) / NT;
  constexpr int SB = (BN * BK / 4) / NT;
  static_assert(SA >= 1 && SA <= 4 && SB >= 1 && SB <= 4, "tile/thread configuration");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int bz = blockIdx.z;

#pragma unroll
  for (int s = 0; s < SA; ++s) al.prep(s, m0 + (s * NT + tid) / 4, bz);
  const float* Wb = bl.W + (long long)bz * bl.strideW;

  float4 ra[SA], rb[SB];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int s = 0; s < SA; ++s) ra[s] = al.fetch(s, k0 + ((s * NT + tid) & 3) * 4);
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      int item = s * NT + tid;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!bl.kn) {
        int n = n0 + (item >> 2), k = k0 + (item & 3) * 4;
        if (n < N && k < K) {
          const float* p = Wb + (long long)n * bl.ldw + k;
          if (bl.vec) {
            v = ld4(p);
          } else {
            v.x = p[0];
            if (k + 1 < K) v.y = p[1];
            if (k + 2 < K) v.z = p[2];
            if (k + 3 < K) v.w = p[3];
          }
        }
      } else {
        int k = k0 + item / (BN / 4), n = n0 + (item % (BN / 4)) * 4;
        if (k < K && n < N) {
          const float* p = Wb + (long long)k * bl.ldw + n;
          if (bl.vec) {
            v = ld4(p);
          } else {
            v.x = p[0];
            if (n + 1 < N) v.y = p[1];
            if (n + 2 < N) v.z = p[2];
            if (n + 3 < N) v.w = p[3];
          }
        }
      }
      rb[s] = v;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int s = 0; s < SA; ++s) {
      int item = s * NT + tid, row = item >> 2, kq = (item & 3) * 4;
      As[buf][kq + 0][row] = ra[s].x;
      As[buf][kq + 1][row] = ra[s].y;
      As[buf][kq + 2][row] = ra[s].z;
      As[buf][kq + 3][row] = ra[s].w;
    }
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      int item = s * NT + tid;
      if (!bl.kn) {
        int row = item >> 2, kq = (item & 3) * 4;
        Bs[buf][kq + 0][row] = rb[s].x;
        Bs[buf][kq + 1][row] = rb[s].y;
        Bs[buf][kq + 2][row] = rb[s].z;
        Bs[buf][kq + 3][row] = rb[s].w;
      } else {
        int k = item / (BN / 4), nq = (item % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][k][nq]) = rb[s];
      }
    }
  };

  float acc[4 * FM][4 * FN];
#pragma unroll
  for (int i = 0; i < 4 * FM; ++i)
#pragma unroll
    for (int j = 0; j < 4 * FN; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4 * FM], b[4 * FN];
#pragma unroll
      for (int f = 0; f < FM; ++f) {
        float4 t = *reinterpret_cast<const float4*>(&As[cur][k][f * (BM / FM) + ty * 4]);
        a[4 * f + 0] = t.x, a[4 * f + 1] = t.y, a[4 * f + 2] = t.z, a[4 * f + 3] = t.w;
      }
#pragma unroll
      for (int g = 0; g < FN; ++g) {
        float4 t = *reinterpret_cast<const float4*>(&Bs[cur][k][g * (BN / FN) + tx * 4]);
        b[4 * g + 0] = t.x, b[4 * g + 1] = t.y, b[4 * g + 2] = t.z, b[4 * g + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < 4 * FM; ++i)
#pragma unroll
        for (int j = 0; j < 4 * FN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) stash(cur ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int f = 0; f < FM; ++f)
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      long long m = m0 + f * (BM / FM) + ty * 4 + ii;
      if (m >= M) continue;
#pragma unroll
      for (int g = 0; g < FN; ++g) {
        int n = n0 + g * (BN / FN) + tx * 4;
        if (n >= N) continue;
        int nvalid = min(4, N - n);
        ep.store(m, n, &acc[4 * f + ii][4 * g], nvalid, bz);
      }
    }
}

template <class AL, class EP>
static int launch_sgemm(AL al, DenseBLoader bl, EP ep, long long M, int N, int K, int batch, cudaStream_t s,
                        const char* name) {
  // tile choice: 128x128 for bulk work, 128x32 when N is narrow / not a multiple of 64 (decoder 160-channel
  // layers, the 3-channel image conv), 64x64 when the 128-tiles would leave most of the 148 SMs idle.
  long long tiles_L = (long long)cdiv(M, 128) * cdiv(N, 128) * batch;
  bool narrow = (N <= 32) || (N % 64 != 0 && N <= 192);
  if (narrow) {
    dim3 grid(cdiv(M, 128), cdiv(N, 32), batch);
    sgemm_kernel<128, 32, 2, 1, AL, EP><<<grid, 128, 0, s>>>(al, bl, ep, M, N, K);
  } else if (tiles_L < 148) {
    dim3 grid(cdiv(M, 64), cdiv(N, 64), batch);
    sgemm_kernel<64, 64, 1, 1, AL, EP><<<grid, 256, 0, s>>>(al, bl, ep, M, N, K);
  } else {
    dim3 grid(cdiv(M, 128), cdiv(N, 128), batch);
    sgemm_kernel<128, 128, 2, 2, AL, EP><<<grid, 256, 0, s>>>(al, bl, ep, M, N, K);
  }
  CVAR_CHECK_LAUNCH(name);
  return 0;
}

}  // namespace cvar
