// C-ABI entry points built on the GEMM engines: dense layers, fused QKV + KV-cache append, decoder convolutions.
#include <stdlib.h>
#include "sgemm.cuh"

using namespace cvar;

namespace cvar {
// tcgen05 engine (gemm_tc.cu); returns 1 when it took the problem, 0 when the shape is not supported, <0 on error.
int tc_gemm_try(const cvar_gemm_args* a, cudaStream_t s);
int tc_qkv_try(const float* A, const float* A_lo, const float* Wqkv_hi, const float* Wqkv_lo, const QkvEpilogue& ep, int M,
               int C, cudaStream_t s);
int tc_split(const float* w, float* hi, float* lo, long long n, cudaStream_t s);
// 2-CTA all-TMA engine (gemm_tc2.cu): needs the activation pre-split (A_lo) as well as the weights
int tc2_gemm_try(const cvar_gemm_args* a, cudaStream_t s);
int tc2_qkv_try(const float* A_hi, const float* A_lo, const float* Wqkv_hi, const float* Wqkv_lo, const QkvEpilogue& ep,
                int M, int C, cudaStream_t s);
int tc_conv_try(const cvar_conv_args* a, cudaStream_t s);
// FP16-pair operands (engine 4): 0 ok, < 0 error - never a fall-through
int tc2_gemm_f16(const cvar_gemm_args* a, cudaStream_t s);
int tc2_conv_f16(const cvar_conv_args* a, cudaStream_t s);
int tc2_conv_f16_supported(int H, int W, int Cin, int Cout, int ks);
int tc2_conv_f16_gn_fusable(int H, int W, int Cin, int Cout, int ks, int groups);
int tc2_qkv_f16(const void* A_hi, const void* A_lo, const void* W_hi, const void* W_lo, const QkvEpilogue& ep, int M, int C,
                cudaStream_t s);
}  // namespace cvar

extern "C" int cvar_gemm(const cvar_gemm_args* a, void* stream) {
  CVAR_REQUIRE(a != nullptr, "cvar_gemm: null args");
  CVAR_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->batch > 0, "cvar_gemm: bad shape M=%d N=%d K=%d", a->M, a->N,
               a->K);
  // 128-bit operand loads need 16-byte aligned rows; ragged problems (the 3x3 / 5x5 decoder attention of truncated
  // pyramids) take the scalar-load variant of the same kernel
  const int a_vec = (a->K % 4 == 0) && (a->lda % 4 == 0) && (a->strideA % 4 == 0) && ((((uintptr_t)a->A) & 15) == 0);
  const int w_vec = (a->ldw % 4 == 0) && (a->strideW % 4 == 0) && ((((uintptr_t)a->W) & 15) == 0) &&
                    (a->w_is_kn ? (a->N % 4 == 0) : (a->K % 4 == 0));
  CVAR_REQUIRE(a->epilogue != CVAR_EPI_BIAS_GAMMA_RESID || (a->gamma && a->rows_per_sample > 0),
               "cvar_gemm: gamma epilogue without gamma");
  CVAR_REQUIRE(a->epilogue != CVAR_EPI_BIAS_RESID || a->resid, "cvar_gemm: residual epilogue without resid");
  cudaStream_t s = (cudaStream_t)stream;
  CVAR_REQUIRE(a->out_lo == nullptr || a->epilogue == CVAR_EPI_BIAS || a->epilogue == CVAR_EPI_BIAS_GELU,
               "cvar_gemm: out_lo only goes with the BIAS / BIAS_GELU epilogues");
  CVAR_REQUIRE((a->out16_hi == nullptr) == (a->out16_lo == nullptr), "cvar_gemm: out16_hi/out16_lo must come together");
  CVAR_REQUIRE(a->out16_hi == nullptr || a->epilogue == CVAR_EPI_BIAS || a->epilogue == CVAR_EPI_BIAS_GELU,
               "cvar_gemm: out16 only goes with the BIAS / BIAS_GELU epilogues");
  CVAR_REQUIRE(a->out != nullptr || (a->out16_hi != nullptr && a->out_lo == nullptr), "cvar_gemm: no output");
  if (a->A16_hi != nullptr) {
    CVAR_REQUIRE(g_gemm_engine != 0, "cvar_gemm: FP16-pair operands need a tensor-core engine (engine is 0 = SIMT)");
    return tc2_gemm_f16(a, s);
  }
  CVAR_REQUIRE(a->A != nullptr && a->W != nullptr, "cvar_gemm: null A / W");
  if (g_gemm_engine >= 3 && a->A_lo != nullptr) {
    int took = tc2_gemm_try(a, s);
    if (took < 0) return took;
    if (took == 1) return 0;
  }
  if (g_gemm_engine != 0 && a->W_hi != nullptr && a->W_lo != nullptr) {
    int took = tc_gemm_try(a, s);
    if (took < 0) return took;
    if (took == 1) return 0;
  }
  CVAR_REQUIRE(a->A_lo == nullptr || a_vec, "cvar_gemm: a pre-split activation needs 16-byte aligned rows");
  DenseALoader al{a->A, a->lda, a->strideA, a->M, a->K, a_vec};
  al.A_lo = a->A_lo;
  DenseBLoader bl{a->W, a->ldw, a->strideW, a->N, a->K, a->w_is_kn, w_vec};
  DenseEpilogue ep{a->out, a->ldo, a->strideO, a->bias, a->epilogue, a->alpha, a->gamma, a->gamma_row_stride,
                   a->rows_per_sample, a->resid, a->ldr, a->strideR, a->out_lo};
  return launch_sgemm(al, bl, ep, (long long)a->M, a->N, a->K, a->batch, s, "cvar_gemm");
}

// F.normalize(q).mul(scale_mul), F.normalize(k)                                         basic_var.py:99-104
__global__ void cos_attn_normalize_kernel(float* __restrict__ q, float* __restrict__ k_hi, float* __restrict__ k_lo,
                                          const float* __restrict__ sm, int R, int H, int l, int L_prev, int T_max) {
  // one warp per 64-float head row; rows [0, R*H*l) are q rows, the next R*H*l are the freshly appended k rows
  // (stored split: k = hi + lo exactly; normalise the sum, split again)
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  long long nq = (long long)R * H * l;
  if (row >= 2 * nq) return;
  bool is_q = row < nq;
  long long i = is_q ? row : row - nq;
  int t = (int)(i % l);
  long long rh = i / l;
  int h = (int)(rh % H);
  long long off = is_q ? i * 64 : (rh * T_max + L_prev + t) * 64;
  float2 v;
  if (is_q) {
    v = *reinterpret_cast<float2*>(q + off + lane * 2);
  } else {
    float2 a = *reinterpret_cast<float2*>(k_hi + off + lane * 2), b = *reinterpret_cast<float2*>(k_lo + off + lane * 2);
    v = make_float2(a.x + b.x, a.y + b.y);
  }
  float ss = warp_sum(v.x * v.x + v.y * v.y);
  float denom = fmaxf(sqrtf(ss), 1e-12f);
  v.x = v.x / denom;
  v.y = v.y / denom;
  if (is_q) {
    float mul = expf(fminf(sm[h], 4.605170185988092f));   // clamp_max(log(100)).exp()
    v.x = __fmul_rn(v.x, mul);
    v.y = __fmul_rn(v.y, mul);
    *reinterpret_cast<float2*>(q + off + lane * 2) = v;
  } else {
    float2 hi = make_float2(__uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u),
                            __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
    *reinterpret_cast<float2*>(k_hi + off + lane * 2) = hi;
    *reinterpret_cast<float2*>(k_lo + off + lane * 2) = make_float2(v.x - hi.x, v.y - hi.y);
  }
}

extern "C" int cvar_qkv_project(const float* A, const float* A_lo, const void* A16_hi, const void* A16_lo,
                                const float* Wqkv, const float* Wqkv_hi, const float* Wqkv_lo, const void* W16_hi,
                                const void* W16_lo,
                                const float* q_bias, const float* k_bias, const float* v_bias, float* q_out,
                                float* k_hi, float* k_lo, float* vt_hi, float* vt_lo, int R, int l, int L_prev,
                                int T_max, int H, int cos_attn, const float* scale_mul_H, void* stream) {
  CVAR_REQUIRE(R > 0 && l > 0 && H > 0 && L_prev >= 0 && L_prev + l <= T_max, "cvar_qkv_project: bad shape");
  CVAR_REQUIRE(!cos_attn || scale_mul_H != nullptr, "cvar_qkv_project: cosine attention needs scale_mul");
  cudaStream_t s = (cudaStream_t)stream;
  const int C = H * 64, M = R * l;
  QkvEpilogue ep{q_bias, k_bias, v_bias, q_out, k_hi, k_lo, vt_hi, vt_lo, C, H, l, L_prev, T_max};
  int took = 0;
  if (A16_hi != nullptr) {
    CVAR_REQUIRE(g_gemm_engine != 0, "cvar_qkv_project: FP16-pair operands need a tensor-core engine (engine is 0 = SIMT)");
    CVAR_REQUIRE(C % 64 == 0, "cvar_qkv_project: C must be a multiple of 64");
    int rc = tc2_qkv_f16(A16_hi, A16_lo, W16_hi, W16_lo, ep, M, C, s);
    if (rc) return rc;
    took = 1;
  }
  if (took == 0 && g_gemm_engine >= 3 && A_lo != nullptr && Wqkv_hi != nullptr && Wqkv_lo != nullptr) {
    took = tc2_qkv_try(A, A_lo, Wqkv_hi, Wqkv_lo, ep, M, C, s);
    if (took < 0) return took;
  }
  if (took == 0 && g_gemm_engine != 0 && Wqkv_hi != nullptr && Wqkv_lo != nullptr) {
    took = tc_qkv_try(A, A_lo, Wqkv_hi, Wqkv_lo, ep, M, C, s);
    if (took < 0) return took;
  }
  if (took == 0) {
    DenseALoader al{A, C, 0, M, C, 1};
    al.A_lo = A_lo;
    DenseBLoader bl{Wqkv, C, 0, 3 * C, C, 0, 1};
    int rc = launch_sgemm(al, bl, ep, (long long)M, 3 * C, C, 1, s, "cvar_qkv_project");
    if (rc) return rc;
  }
  if (cos_attn) {
    long long rows = 2LL * R * H * l;
    cos_attn_normalize_kernel<<<cdiv(rows, 8), 256, 0, s>>>(q_out, k_hi, k_lo, scale_mul_H, R, H, l, L_prev, T_max);
    CVAR_CHECK_LAUNCH("cvar_qkv_project/cos_normalize");
  }
  return 0;
}

// F.normalize(q).mul(scale_mul), F.normalize(k) on the FP16-pair form                    basic_var.py:99-104
__global__ void cos_attn_normalize16_kernel(__half* __restrict__ q_hi, __half* __restrict__ q_lo,
                                            __half* __restrict__ k_hi, __half* __restrict__ k_lo,
                                            const float* __restrict__ sm, int R, int H, int l, int L_prev, int T_max) {
  // one warp per 64-element head row; rows [0, R*H*l) are q rows, the next R*H*l are the freshly appended k rows
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  long long nq = (long long)R * H * l;
  if (row >= 2 * nq) return;
  bool is_q = row < nq;
  long long i = is_q ? row : row - nq;
  int t = (int)(i % l);
  long long rh = i / l;
  int h = (int)(rh % H);
  long long off = (is_q ? i * 64 : (rh * T_max + L_prev + t) * 64) + lane * 2;
  __half* ph = (is_q ? q_hi : k_hi) + off;
  __half* pl = (is_q ? q_lo : k_lo) + off;
  const __half2 a = *reinterpret_cast<const __half2*>(ph), b = *reinterpret_cast<const __half2*>(pl);
  // qk pairs: 16 x = hi + lo
  float2 v = make_float2((__half2float(__low2half(a)) + __half2float(__low2half(b))) * (1.0f / kQkScale),
                         (__half2float(__high2half(a)) + __half2float(__high2half(b))) * (1.0f / kQkScale));
  float ss = warp_sum(v.x * v.x + v.y * v.y);
  float denom = fmaxf(sqrtf(ss), 1e-12f);
  v.x = v.x / denom;
  v.y = v.y / denom;
  if (is_q) {
    float mul = expf(fminf(sm[h], 4.605170185988092f));   // clamp_max(log(100)).exp()
    v.x = __fmul_rn(v.x, mul);
    v.y = __fmul_rn(v.y, mul);
  }
  __half h0, l0, h1, l1;
  split_f16_qk(v.x, h0, l0);
  split_f16_qk(v.y, h1, l1);
  *reinterpret_cast<__half2*>(ph) = __halves2half2(h0, h1);
  *reinterpret_cast<__half2*>(pl) = __halves2half2(l0, l1);
}

extern "C" int cvar_qkv_project16(const void* A16_hi, const void* A16_lo, const void* W16_hi, const void* W16_lo,
                                  const float* q_bias, const float* k_bias, const float* v_bias, void* q16_hi,
                                  void* q16_lo, void* k16_hi, void* k16_lo, void* vt16_hi, void* vt16_lo, int R, int l,
                                  int L_prev, int T_max, int H, int cos_attn, const float* scale_mul_H, void* stream) {
  CVAR_REQUIRE(R > 0 && l > 0 && H > 0 && L_prev >= 0 && L_prev + l <= T_max, "cvar_qkv_project16: bad shape");
  CVAR_REQUIRE(T_max % 8 == 0, "cvar_qkv_project16: T_max must be a multiple of 8 (got %d)", T_max);
  CVAR_REQUIRE(!cos_attn || scale_mul_H != nullptr, "cvar_qkv_project16: cosine attention needs scale_mul");
  CVAR_REQUIRE(q16_hi && q16_lo && k16_hi && k16_lo && vt16_hi && vt16_lo, "cvar_qkv_project16: null output");
  CVAR_REQUIRE(g_gemm_engine != 0, "cvar_qkv_project16: FP16-pair operands need a tensor-core engine (engine is 0 = SIMT)");
  cudaStream_t s = (cudaStream_t)stream;
  const int C = H * 64, M = R * l;
  QkvEpilogue ep{q_bias, k_bias, v_bias, nullptr, nullptr, nullptr, nullptr, nullptr, C, H, l, L_prev, T_max};
  ep.q16_hi = reinterpret_cast<__half*>(q16_hi), ep.q16_lo = reinterpret_cast<__half*>(q16_lo);
  ep.k16_hi = reinterpret_cast<__half*>(k16_hi), ep.k16_lo = reinterpret_cast<__half*>(k16_lo);
  ep.vt16_hi = reinterpret_cast<__half*>(vt16_hi), ep.vt16_lo = reinterpret_cast<__half*>(vt16_lo);
  int rc = tc2_qkv_f16(A16_hi, A16_lo, W16_hi, W16_lo, ep, M, C, s);
  if (rc) return rc;
  if (cos_attn) {
    long long rows = 2LL * R * H * l;
    cos_attn_normalize16_kernel<<<cdiv(rows, 8), 256, 0, s>>>(ep.q16_hi, ep.q16_lo, ep.k16_hi, ep.k16_lo, scale_mul_H, R, H,
                                                              l, L_prev, T_max);
    CVAR_CHECK_LAUNCH("cvar_qkv_project16/cos_normalize");
  }
  return 0;
}

// 3x3 'same' convolution to THREE output channels (Decoder.conv_out, vae_modules.py:226): the implicit-GEMM engines pad
// N = 3 to a 32-wide tile and spend 10x the FLOPs.  Direct form: one thread per output pixel, all weights (27 * Cin floats)
// in shared memory (broadcast reads), the pixel's channel vectors read as float4 through L1; epilogue = ConvEpilogue.
template <int COUT>
__global__ void __launch_bounds__(128) conv3x3_small_cout_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 ConvEpilogue ep, int H, int W, int Cin, long long M) {
  extern __shared__ __align__(16) float w_s[];          // [COUT][9 * Cin]
  const int K = 9 * Cin;
  for (int i = threadIdx.x; i < COUT * K; i += blockDim.x) w_s[i] = w[i];
  __syncthreads();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int xw = (int)(m % W);
  const long long q = m / W;
  const int y = (int)(q % H);
  const long long n = q / H;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = xw + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const float* px = x + ((n * H + yy) * W + xx) * Cin;
    const float* wt = w_s + tap * Cin;
#pragma unroll 4
    for (int ci = 0; ci < Cin; ci += 4) {
      const float4 v = ld4(px + ci);
#pragma unroll
      for (int c = 0; c < COUT; ++c) {
        const float4 ww = *reinterpret_cast<const float4*>(wt + c * K + ci);
        acc[c] = fmaf(v.x, ww.x, acc[c]);
        acc[c] = fmaf(v.y, ww.y, acc[c]);
        acc[c] = fmaf(v.z, ww.z, acc[c]);
        acc[c] = fmaf(v.w, ww.w, acc[c]);
      }
    }
  }
  float v4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < COUT; ++c) v4[c] = acc[c];
  ep.store(m, 0, v4, COUT, 0);
}

// The same convolution for wide images (W a multiple of 128): the kernel above reads every input element 9 times through
// L1 with one 640-byte-strided pixel per thread - 32 sectors per request, 12.5 ms for the 128 x 256 x 256 x 160 input of a
// bench step against 0.8 ms of HBM time (ncu launch list, profiles/r02_launches_table.md).  Here a block of 256 threads owns
// a 2 x 128 output tile; per 32-channel chunk the 4 x 130 halo is staged in shared memory with coalesced 16-byte loads
// (pixel pitch 36 floats: conflict-free float4 reads) and read 9 times from there; every input element leaves L2 twice.
// Summation order: channel chunks outermost (the kernel above: taps outermost) - fp32 results differ in the last bits.
constexpr int kCoTH = 2, kCoTW = 128, kCoCh = 32, kCoPitch = 36;
template <int COUT>
__global__ void __launch_bounds__(kCoTH * kCoTW) conv3x3_small_cout_tiled_kernel(const float* __restrict__ x,
                                                                                   const float* __restrict__ w, ConvEpilogue ep,
                                                                                   int H, int W, int Cin) {
  extern __shared__ __align__(16) float smem_co[];
  const int K = 9 * Cin;
  float* w_s = smem_co;                                   // [COUT][9 * Cin]
  float* tile = smem_co + COUT * K;                       // [kCoTH + 2][kCoTW + 2][kCoPitch]
  const int tid = threadIdx.x;
  for (int i = tid; i < COUT * K / 4; i += blockDim.x) reinterpret_cast<float4*>(w_s)[i] = ld4(w + 4 * i);
  const int x0 = blockIdx.x * kCoTW, y0 = blockIdx.y * kCoTH;
  const long long n = blockIdx.z;
  const int ty = tid / kCoTW, px = tid - ty * kCoTW;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
  constexpr int kHaloVec = (kCoTH + 2) * (kCoTW + 2) * (kCoCh / 4);
  for (int c0 = 0; c0 < Cin; c0 += kCoCh) {
    __syncthreads();                                      // previous chunk consumed (and w_s filled, first time round)
    for (int i = tid; i < kHaloVec; i += kCoTH * kCoTW) {
      const int c4 = i & (kCoCh / 4 - 1);
      const int pp = i / (kCoCh / 4);
      const int r = pp / (kCoTW + 2), p = pp - r * (kCoTW + 2);
      const int yy = y0 - 1 + r, xx = x0 - 1 + p;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = ld4(x + ((n * H + yy) * W + xx) * Cin + c0 + c4 * 4);
      *reinterpret_cast<float4*>(tile + (r * (kCoTW + 2) + p) * kCoPitch + c4 * 4) = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap - 3 * ky;
      const float* xs = tile + ((ty + ky) * (kCoTW + 2) + px + kx) * kCoPitch;
      const float* wt = w_s + tap * Cin + c0;
#pragma unroll
      for (int c4 = 0; c4 < kCoCh / 4; ++c4) {
        const float4 v = *reinterpret_cast<const float4*>(xs + c4 * 4);
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          const float4 ww = *reinterpret_cast<const float4*>(wt + c * K + c4 * 4);
          acc[c] = fmaf(v.x, ww.x, acc[c]);
          acc[c] = fmaf(v.y, ww.y, acc[c]);
          acc[c] = fmaf(v.z, ww.z, acc[c]);
          acc[c] = fmaf(v.w, ww.w, acc[c]);
        }
      }
    }
  }
  float v4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < COUT; ++c) v4[c] = acc[c];
  const long long m = (n * H + (y0 + ty)) * W + x0 + px;
  ep.store(m, 0, v4, COUT, 0);
}

// Round 2, second step: the tile kernel above is bound by shared-memory bandwidth, not by its FMAs - per 16-byte slice of a
// pixel it issues one 4-wavefront activation read and three broadcast weight reads for 12 FMAs (7 wavefronts per 12 warp-FMAs; the
// LSU pipe delivers one wavefront per cycle, the four schedulers 12 FMAs in 3).  Here a thread owns FOUR vertically adjacent
// output pixels: the six input rows it needs are read once per (kx, channel slice) and feed 3 ky x 3 channels x 4 rows x 4 = 144
// FMAs against 6 + 9 reads (2.75 wavefronts per 12 FMAs).  Block = 128 threads = a 4 x 128 tile, 16-channel chunks (pixel pitch
// 20 floats: conflict-free float4 reads), the chunk's 432 weights reloaded per chunk: 64 KB of shared memory, three CTAs per SM,
// and the 6 x 130 halo costs 1.5 reads per input element instead of 2.
constexpr int kR4TH = 4, kR4TW = 128, kR4Ch = 16, kR4Pitch = 20;
template <int COUT>
__global__ void __launch_bounds__(kR4TW) conv3x3_small_cout_rows4_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                         ConvEpilogue ep, int H, int W, int Cin) {
  extern __shared__ __align__(16) float smem_r4[];
  float* w_s = smem_r4;                                   // [COUT][9][kR4Ch]
  float* tile = smem_r4 + COUT * 9 * kR4Ch;               // [kR4TH + 2][kR4TW + 2][kR4Pitch]
  const int K = 9 * Cin;
  const int px = threadIdx.x;
  const int x0 = blockIdx.x * kR4TW, y0 = blockIdx.y * kR4TH;
  const long long n = blockIdx.z;
  float acc[kR4TH][COUT];
#pragma unroll
  for (int j = 0; j < kR4TH; ++j)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[j][c] = 0.f;
  constexpr int kHaloVec = (kR4TH + 2) * (kR4TW + 2) * (kR4Ch / 4);
  for (int c0 = 0; c0 < Cin; c0 += kR4Ch) {
    __syncthreads();                                      // previous chunk consumed
    for (int i = px; i < COUT * 9 * (kR4Ch / 4); i += kR4TW) {
      const int c4 = i & (kR4Ch / 4 - 1);
      const int ct = i / (kR4Ch / 4);                     // channel * 9 + tap
      const int c = ct / 9, tap = ct - 9 * c;
      *reinterpret_cast<float4*>(w_s + ct * kR4Ch + c4 * 4) = ld4(w + (long long)c * K + tap * Cin + c0 + c4 * 4);
    }
    for (int i = px; i < kHaloVec; i += kR4TW) {
      const int c4 = i & (kR4Ch / 4 - 1);
      const int pp = i / (kR4Ch / 4);
      const int r = pp / (kR4TW + 2), p = pp - r * (kR4TW + 2);
      const int yy = y0 - 1 + r, xx = x0 - 1 + p;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = ld4(x + ((n * H + yy) * W + xx) * Cin + c0 + c4 * 4);
      *reinterpret_cast<float4*>(tile + (r * (kR4TW + 2) + p) * kR4Pitch + c4 * 4) = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
      for (int c4 = 0; c4 < kR4Ch / 4; ++c4) {
        float4 xv[kR4TH + 2];
#pragma unroll
        for (int r = 0; r < kR4TH + 2; ++r)
          xv[r] = *reinterpret_cast<const float4*>(tile + (r * (kR4TW + 2) + px + kx) * kR4Pitch + c4 * 4);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
          for (int c = 0; c < COUT; ++c) {
            const float4 ww = *reinterpret_cast<const float4*>(w_s + (c * 9 + ky * 3 + kx) * kR4Ch + c4 * 4);
#pragma unroll
            for (int j = 0; j < kR4TH; ++j) {
              acc[j][c] = fmaf(xv[j + ky].x, ww.x, acc[j][c]);
              acc[j][c] = fmaf(xv[j + ky].y, ww.y, acc[j][c]);
              acc[j][c] = fmaf(xv[j + ky].z, ww.z, acc[j][c]);
              acc[j][c] = fmaf(xv[j + ky].w, ww.w, acc[j][c]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kR4TH; ++j) {
    float v4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < COUT; ++c) v4[c] = acc[j][c];
    const long long m = (n * H + (y0 + j)) * W + x0 + px;
    ep.store(m, 0, v4, COUT, 0);
  }
}

extern "C" int cvar_conv2d(const cvar_conv_args* a, void* stream) {
  CVAR_REQUIRE(a != nullptr, "cvar_conv2d: null args");
  CVAR_REQUIRE(a->ks == 1 || a->ks == 3, "cvar_conv2d: ks must be 1 or 3");
  CVAR_REQUIRE(a->Cin % 16 == 0, "cvar_conv2d: Cin must be a multiple of 16 (got %d)", a->Cin);
  CVAR_REQUIRE(a->out_mode >= 0 && a->out_mode <= 3, "cvar_conv2d: bad out_mode");
  CVAR_REQUIRE(!a->downsample2x || (a->ks == 3 && !a->upsample2x && a->Hin % 2 == 0 && a->Win % 2 == 0 &&
                                    a->x16_hi == nullptr && a->in_a == nullptr),
               "cvar_conv2d: downsample2x needs ks = 3, even Hin / Win, fp32 input, no upsampling / fused input transform");
  CVAR_REQUIRE(a->out_mode == 0 || a->resid == nullptr, "cvar_conv2d: image output takes no residual");
  CVAR_REQUIRE(a->out_samples >= 0 && (a->out_samples == 0 || (a->out_mode != 0 && a->B % a->out_samples == 0)),
               "cvar_conv2d: out_samples must divide B and needs an image out_mode (B=%d out_samples=%d)", a->B, a->out_samples);
  CVAR_REQUIRE((a->in_a == nullptr) == (a->in_b == nullptr), "cvar_conv2d: in_a/in_b must come together");
  cudaStream_t s = (cudaStream_t)stream;
  if (a->x16_hi != nullptr) {
    CVAR_REQUIRE(g_gemm_engine != 0, "cvar_conv2d: FP16-pair operands need a tensor-core engine (engine is 0 = SIMT)");
    return tc2_conv_f16(a, s);
  }
  CVAR_REQUIRE(a->x != nullptr && a->w != nullptr, "cvar_conv2d: null x / w");
  CVAR_REQUIRE(a->gn_part == nullptr, "cvar_conv2d: gn_part is produced by the FP16-pair kernel only (x16_* operands)");
  if (g_gemm_engine != 0 && a->engine != 0 && a->w_hi != nullptr && a->w_lo != nullptr) {
    int took = tc_conv_try(a, s);
    if (took < 0) return took;
    if (took == 1) return 0;
  }
  const int up = a->upsample2x ? 1 : 0, down = a->downsample2x ? 1 : 0;
  const int Hout = (a->Hin << up) >> down, Wout = (a->Win << up) >> down;
  const long long M = (long long)a->B * Hout * Wout;
  const int K = a->ks * a->ks * a->Cin;
  if (a->Cout == 3 && a->ks == 3 && !up && !down && a->in_a == nullptr && a->resid == nullptr && a->Cin % 4 == 0 &&
      (size_t)3 * K * sizeof(float) <= 48 * 1024) {
    ConvEpilogue ep3{a->out, a->bias, nullptr, a->Cout, a->out_mode, Hout, Wout, a->out_rows_total, a->row_offset};
    ep3.out_samples = a->out_samples;
    static const bool rows4_off = getenv("CVAR_CONV3_ROWS4") != nullptr && getenv("CVAR_CONV3_ROWS4")[0] == '0';   // A/B (diagnostic)
    if (!rows4_off && Wout % kR4TW == 0 && Hout % kR4TH == 0 && a->Cin % kR4Ch == 0 && a->B <= 65535 &&
        ((((uintptr_t)a->x) | ((uintptr_t)a->w)) & 15) == 0) {
      const size_t smem_r = ((size_t)3 * 9 * kR4Ch + (size_t)(kR4TH + 2) * (kR4TW + 2) * kR4Pitch) * sizeof(float);
      auto kern = conv3x3_small_cout_rows4_kernel<3>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
      CVAR_REQUIRE(e == cudaSuccess, "cvar_conv2d[cout3]: cannot raise shared memory: %s", cudaGetErrorString(e));
      dim3 grid(Wout / kR4TW, Hout / kR4TH, a->B);
      kern<<<grid, kR4TW, smem_r, s>>>(a->x, a->w, ep3, Hout, Wout, a->Cin);
      CVAR_CHECK_LAUNCH("cvar_conv2d[cout3/rows4]");
      return 0;
    }
    const size_t smem_t = ((size_t)3 * K + (size_t)(kCoTH + 2) * (kCoTW + 2) * kCoPitch) * sizeof(float);
    if (Wout % kCoTW == 0 && Hout % kCoTH == 0 && a->Cin % kCoCh == 0 && a->B <= 65535 && smem_t <= 100 * 1024 &&
        ((((uintptr_t)a->x) | ((uintptr_t)a->w)) & 15) == 0) {
      auto kern = conv3x3_small_cout_tiled_kernel<3>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
      CVAR_REQUIRE(e == cudaSuccess, "cvar_conv2d[cout3]: cannot raise shared memory: %s", cudaGetErrorString(e));
      dim3 grid(Wout / kCoTW, Hout / kCoTH, a->B);
      kern<<<grid, kCoTH * kCoTW, smem_t, s>>>(a->x, a->w, ep3, Hout, Wout, a->Cin);
      CVAR_CHECK_LAUNCH("cvar_conv2d[cout3/tiled]");
      return 0;
    }
    conv3x3_small_cout_kernel<3><<<cdiv(M, 128), 128, (size_t)3 * K * sizeof(float), s>>>(a->x, a->w, ep3, Hout, Wout,
                                                                                         a->Cin, M);
    CVAR_CHECK_LAUNCH("cvar_conv2d[cout3]");
    return 0;
  }
  ConvALoader al;
  al.x = a->x, al.in_a = a->in_a, al.in_b = a->in_b, al.in_silu = a->in_silu;
  al.Hin = a->Hin, al.Win = a->Win, al.Cin = a->Cin, al.ks = a->ks, al.up = up;
  al.Hout = Hout, al.Wout = Wout, al.Mtot = M, al.K = K;
  al.Hv = a->Hin << up, al.Wv = a->Win << up, al.stride = down ? 2 : 1, al.pad = down ? 0 : (a->ks >> 1);
  DenseBLoader bl{a->w, K, 0, a->Cout, K, 0, 1};
  ConvEpilogue ep{a->out, a->bias, a->resid, a->Cout, a->out_mode, Hout, Wout, a->out_rows_total, a->row_offset};
  ep.out_samples = a->out_samples;
  return launch_sgemm(al, bl, ep, M, a->Cout, K, 1, s, "cvar_conv2d");
}

extern "C" int cvar_conv2d_f16_supported(int H, int W, int Cin, int Cout, int ks) {
  return tc2_conv_f16_supported(H, W, Cin, Cout, ks);
}
extern "C" int cvar_conv2d_gn_fusable(int H, int W, int Cin, int Cout, int ks, int groups) {
  return tc2_conv_f16_gn_fusable(H, W, Cin, Cout, ks, groups);
}

extern "C" int cvar_split_tf32(const float* w, float* w_hi, float* w_lo, long long n, void* stream) {
  CVAR_REQUIRE(n > 0 && n % 4 == 0, "cvar_split_tf32: n must be a positive multiple of 4");
  return tc_split(w, w_hi, w_lo, n, (cudaStream_t)stream);
}
