// Shared helpers for libcvar_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/cvar.h"

namespace cvar {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
extern int g_gemm_engine;
extern int g_fast_mode;     // 1: single-MMA FP16 operands (hi halves only) - NOT a parity mode (cvar_set_fast_mode)

#define CVAR_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      cvar::set_error(__VA_ARGS__);             \
      return -1;                                \
    }                                           \
  } while (0)

// After a launch: catch configuration errors synchronously (asynchronous faults surface at the caller's next sync).
#define CVAR_CHECK_LAUNCH(name)                                                           \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      cvar::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));            \
      return -2;                                                                          \
    }                                                                                     \
    cvar::count_launch();                                                                 \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ln_stream.cu: the persistent streaming form of cvar_ln_modulate (FP16-pair output, large M)
bool ln_stream_usable(int M, int C);
int launch_ln_stream(const float* x, const float* scale, const float* shift, long long mod_stride, __half* y16_hi,
                     __half* y16_lo, int M, int C, int rows_per_sample, float eps, cudaStream_t s);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// x * sigmoid(x), the form ATen evaluates: x / (1 + exp(-x))      (vae_modules.py:14-15, F.silu)
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

// GELU(approximate='tanh')                                         (basic_var.py:39)
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float kBeta = 0.7978845608028654f;   // sqrt(2/pi)
  const float kKappa = 0.044715f;
  float inner = kBeta * (x + kKappa * x * x * x);
  // 0.5 * x * (1 + tanh(u)) == x * sigmoid(2u) == x / (1 + exp(-2u)): same function, no tanhf, no cancellation
  return x / (1.0f + expf(-2.0f * inner));
}

// FP16 pair of an fp32 value (operand format of the f16x3 tensor-core engine): x ~= h + l * 2^-11, both rounded to
// nearest, so |x - (h + l/2048)| <= 2^-24 |x|.  The residual is scaled by 2^11 to stay in fp16's normal range.
constexpr float kF16LoScale = 2048.0f;
__device__ __forceinline__ void split_f16(float x, __half& h, __half& l) {
  x = fminf(fmaxf(x, -65504.0f), 65504.0f);
  h = __float2half_rn(x);
  l = __float2half_rn((x - __half2float(h)) * kF16LoScale);
}
// four consecutive values -> 8-byte stores of the hi and lo halves.  Packed conversions (F2FP, two values per
// instruction, not on the XU pipe that the scalar F2F shares with MUFU); bit-identical to four split_f16 calls.
__device__ __forceinline__ void st4_split_f16(__half* hi, __half* lo, const float* r) {
  float c[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) c[j] = fminf(fmaxf(r[j], -65504.0f), 65504.0f);
  const __half2 h01 = __floats2half2_rn(c[0], c[1]), h23 = __floats2half2_rn(c[2], c[3]);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn((c[0] - f01.x) * kF16LoScale, (c[1] - f01.y) * kF16LoScale);
  const __half2 l23 = __floats2half2_rn((c[2] - f23.x) * kF16LoScale, (c[3] - f23.y) * kF16LoScale);
  uint2 ph, pl;
  ph.x = *reinterpret_cast<const uint32_t*>(&h01), ph.y = *reinterpret_cast<const uint32_t*>(&h23);
  pl.x = *reinterpret_cast<const uint32_t*>(&l01), pl.y = *reinterpret_cast<const uint32_t*>(&l23);
  *reinterpret_cast<uint2*>(hi) = ph;
  *reinterpret_cast<uint2*>(lo) = pl;
}

// q / K operands of the f16 attention kernel: the value is pre-multiplied by 16 and the residual is stored UN-scaled,
// x * 16 ~= h + l, so that q.k = (h_q h_k + h_q l_k + l_q h_k) / 256 accumulates in ONE tensor-core accumulator (the
// standard pair needs a second one for the 2^-11-scaled cross terms).  The factor 16 keeps the residual of an O(1) value
// a normal fp16 number; below that it is quantised to 2^-24 / 16 = 3.7e-9 absolute, far under fp32 resolution of a
// dot product of O(1) terms.  |x| must be < 4094.
constexpr float kQkScale = 16.0f;
__device__ __forceinline__ void split_f16_qk(float x, __half& h, __half& l) {
  x = fminf(fmaxf(x * kQkScale, -65504.0f), 65504.0f);
  h = __float2half_rn(x);
  l = __float2half_rn(x - __half2float(h));
}
__device__ __forceinline__ void st4_split_f16_qk(__half* hi, __half* lo, const float* r) {
  __half h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_f16_qk(r[j], h[j], l[j]);
  uint2 ph, pl;
  ph.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
  ph.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
  pl.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
  pl.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
  *reinterpret_cast<uint2*>(hi) = ph;
  *reinterpret_cast<uint2*>(lo) = pl;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace cvar
