// Library state + the small HBM-bound kernels of the path (prologue, AdaLN LayerNorm, GroupNorm statistics, layout).
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include "common.cuh"

namespace cvar {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
// default: the fastest engine that has passed the whole parity suite (tests/test_gpu_*.py); CVAR_GEMM_ENGINE overrides
static int initial_engine() {
  const char* e = getenv("CVAR_GEMM_ENGINE");
  if (e != nullptr && (e[0] == '0' || e[0] == '1' || e[0] == '3' || e[0] == '4') && e[1] == '\0') return e[0] - '0';
  return 4;
}
int g_gemm_engine = initial_engine();

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace cvar

using namespace cvar;

extern "C" int cvar_abi_version(void) { return CVAR_ABI_VERSION; }
extern "C" const char* cvar_last_error(void) { return cvar::g_err; }
extern "C" long long cvar_launch_count(void) { return cvar::g_launches.load(); }
extern "C" int cvar_set_gemm_engine(int e) {
  int old = cvar::g_gemm_engine;
  if (e == 0 || e == 1 || e == 3 || e == 4) cvar::g_gemm_engine = e;
  return old;
}
extern "C" int cvar_get_gemm_engine(void) { return cvar::g_gemm_engine; }
namespace cvar {
extern int g_epi_overlap;
static int initial_fast_mode() {
  const char* e = getenv("CVAR_FAST_MODE");
  return (e != nullptr && e[0] == '1') ? 1 : 0;
}
int g_fast_mode = initial_fast_mode();
}  // namespace cvar
extern "C" int cvar_add_launch_count(long long n) {
  cvar::count_launch((int)n);
  return 0;
}
extern "C" int cvar_set_fast_mode(int on) {
  int old = cvar::g_fast_mode;
  cvar::g_fast_mode = on ? 1 : 0;
  return old;
}
extern "C" int cvar_get_fast_mode(void) { return cvar::g_fast_mode; }
extern "C" int cvar_set_epilogue_overlap(int on) {
  int old = cvar::g_epi_overlap;
  cvar::g_epi_overlap = on ? 1 : 0;
  return old;
}
namespace cvar { namespace tc { extern int g_tc_bk; int set_trace(long long*); } }
extern "C" int cvar_debug_set_trace(long long* dev_buf) { return cvar::tc::set_trace(dev_buf); }
extern "C" int cvar_set_tc_kblock(int bk) {
  int old = cvar::tc::g_tc_bk;
  if (bk == 16 || bk == 32) cvar::tc::g_tc_bk = bk;
  return old;
}

// ------------------------------------------------------------------------------------------------ lvl_pos
__global__ void lvl_pos_kernel(const float* __restrict__ lvl_embed, const int64_t* __restrict__ lvl_1L,
                               const float* __restrict__ pos, float* __restrict__ out, int T, int C) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // float4 index
  long long n4 = (long long)T * C / 4;
  if (i >= n4) return;
  int t = (int)(i / (C / 4));
  int c4 = (int)(i % (C / 4));
  float4 a = ld4(lvl_embed + (long long)lvl_1L[t] * C + c4 * 4);
  float4 b = ld4(pos + (long long)t * C + c4 * 4);
  st4(out + i * 4, make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w));
}

extern "C" int cvar_lvl_pos(const float* lvl_embed, const int64_t* lvl_1L, const float* pos_1LC, float* lvl_pos,
                            int T, int C, void* stream) {
  CVAR_REQUIRE(C % 4 == 0 && T > 0, "cvar_lvl_pos: bad shape T=%d C=%d", T, C);
  long long n4 = (long long)T * C / 4;
  lvl_pos_kernel<<<cdiv(n4, 256), 256, 0, (cudaStream_t)stream>>>(lvl_embed, lvl_1L, pos_1LC, lvl_pos, T, C);
  CVAR_CHECK_LAUNCH("cvar_lvl_pos");
  return 0;
}

// ----------------------------------------------------------------------------------------------- prologue
__global__ void prologue_kernel(const float* __restrict__ class_emb, const float* __restrict__ cond_embed,
                                const float* __restrict__ pos_start, const float* __restrict__ lvl_pos,
                                const int64_t* __restrict__ label, const int64_t* __restrict__ ctype, int B, int C,
                                int num_classes, float* __restrict__ cond_BD, float* __restrict__ silu_cond,
                                float* __restrict__ x0) {
  int r = blockIdx.x;
  long long lab = r < B ? label[r] : (long long)num_classes;
  long long ct = r < B ? ctype[r] : 4;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float ce = class_emb[lab * C + c];
    float te = cond_embed != nullptr ? cond_embed[ct * C + c] : ce;   // multi_cond=False: both start tokens = sos
    cond_BD[(long long)r * C + c] = ce;
    silu_cond[(long long)r * C + c] = silu_f(ce);
    // (next_token_map + pos_start) + lvl_pos[:, :2]                      control_var.py:409
    x0[((long long)r * 2 + 0) * C + c] = __fadd_rn(__fadd_rn(te, pos_start[c]), lvl_pos[c]);
    x0[((long long)r * 2 + 1) * C + c] = __fadd_rn(__fadd_rn(ce, pos_start[C + c]), lvl_pos[C + c]);
  }
}

extern "C" int cvar_prologue(const float* class_emb, const float* cond_embed, const float* pos_start,
                             const float* lvl_pos, const int64_t* label_B, const int64_t* cond_type_B, int B, int C,
                             int num_classes, float* cond_BD, float* silu_cond, float* x0, void* stream) {
  CVAR_REQUIRE(B > 0 && C > 0, "cvar_prologue: bad shape");
  prologue_kernel<<<2 * B, 256, 0, (cudaStream_t)stream>>>(class_emb, cond_embed, pos_start, lvl_pos, label_B,
                                                           cond_type_B, B, C, num_classes, cond_BD, silu_cond, x0);
  CVAR_CHECK_LAUNCH("cvar_prologue");
  return 0;
}

extern "C" int cvar_prologue_rows(const float* class_emb, const float* cond_embed, const float* pos_start,
                                  const float* lvl_pos, const int64_t* label_R, const int64_t* cond_type_R, int R, int C,
                                  float* cond_BD, float* silu_cond, float* x0, void* stream) {
  CVAR_REQUIRE(R > 0 && C > 0 && label_R && cond_type_R, "cvar_prologue_rows: bad arguments");
  // B = R: every row reads its own ids, the implicit "unconditional half" of cvar_prologue is never taken
  prologue_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(class_emb, cond_embed, pos_start, lvl_pos, label_R, cond_type_R,
                                                       R, C, 0, cond_BD, silu_cond, x0);
  CVAR_CHECK_LAUNCH("cvar_prologue_rows");
  return 0;
}

// ------------------------------------------------------------------------------------------- ln_modulate
// One warp per row; the row (C <= 2048 floats) lives in registers between the statistics pass and the write.
template <int MAXV>   // float4 per lane
__global__ void __launch_bounds__(256) ln_modulate_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, long long mod_stride,
                                                          float* __restrict__ y, float* __restrict__ y_lo,
                                                          __half* __restrict__ y16_hi, __half* __restrict__ y16_lo, int M, int C,
                                                          int rows_per_sample, float eps) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long m = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (m >= M) return;
  const float* xr = x + m * C;
  int nv = C >> 2;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int idx = lane + i * 32;
    if (idx < nv) {
      v[i] = ld4(xr + idx * 4);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  float mean = warp_sum(s) / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int idx = lane + i * 32;
    if (idx < nv) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
  }
  float var = warp_sum(ss) / (float)C;
  float rstd = 1.0f / sqrtf(var + eps);
  long long r = m / rows_per_sample;
  const float* sc = scale + r * mod_stride;
  const float* sh = shift + r * mod_stride;
  float* yr = y + m * C;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int idx = lane + i * 32;
    if (idx < nv) {
      float4 a = ld4(sc + idx * 4), b = ld4(sh + idx * 4), o;
      // ln(x).mul(scale.add(1)).add_(shift): three separately rounded steps          basic_var.py:208
      o.x = __fadd_rn(__fmul_rn(__fmul_rn(v[i].x - mean, rstd), __fadd_rn(a.x, 1.f)), b.x);
      o.y = __fadd_rn(__fmul_rn(__fmul_rn(v[i].y - mean, rstd), __fadd_rn(a.y, 1.f)), b.y);
      o.z = __fadd_rn(__fmul_rn(__fmul_rn(v[i].z - mean, rstd), __fadd_rn(a.z, 1.f)), b.z);
      o.w = __fadd_rn(__fmul_rn(__fmul_rn(v[i].w - mean, rstd), __fadd_rn(a.w, 1.f)), b.w);
      if (y16_hi != nullptr) {     // FP16 pair for the f16x3 GEMM
        const float ov[4] = {o.x, o.y, o.z, o.w};
        st4_split_f16(y16_hi + m * C + idx * 4, y16_lo + m * C + idx * 4, ov);
        if (y == nullptr) continue;
      }
      if (y_lo != nullptr) {       // TF32 split for the all-TMA GEMM: hi keeps the top 19 bits, lo is the exact remainder
        float4 hi = make_float4(__uint_as_float(__float_as_uint(o.x) & 0xFFFFE000u), __uint_as_float(__float_as_uint(o.y) & 0xFFFFE000u),
                                __uint_as_float(__float_as_uint(o.z) & 0xFFFFE000u), __uint_as_float(__float_as_uint(o.w) & 0xFFFFE000u));
        st4(y_lo + m * C + idx * 4, make_float4(o.x - hi.x, o.y - hi.y, o.z - hi.z, o.w - hi.w));
        o = hi;
      }
      st4(yr + idx * 4, o);
    }
  }
}

extern "C" int cvar_ln_modulate(const float* x, const float* scale, const float* shift, long long mod_row_stride,
                                float* y, float* y_lo, void* y16_hi_, void* y16_lo_, int M, int C, int rows_per_sample,
                                float eps, void* stream) {
  CVAR_REQUIRE(y != nullptr || y16_hi_ != nullptr, "cvar_ln_modulate: no output");
  CVAR_REQUIRE((y16_hi_ == nullptr) == (y16_lo_ == nullptr), "cvar_ln_modulate: y16_hi/y16_lo must come together");
  CVAR_REQUIRE(y != nullptr || y_lo == nullptr, "cvar_ln_modulate: y_lo without y");
  __half* y16_hi = reinterpret_cast<__half*>(y16_hi_);
  __half* y16_lo = reinterpret_cast<__half*>(y16_lo_);
  CVAR_REQUIRE(C % 4 == 0 && C <= 2048 && M > 0 && rows_per_sample > 0, "cvar_ln_modulate: bad shape M=%d C=%d", M, C);
  CVAR_REQUIRE(mod_row_stride % 4 == 0, "cvar_ln_modulate: modulation stride must be a multiple of 4 floats");
  dim3 grid(cdiv(M, 8));
  cudaStream_t s = (cudaStream_t)stream;
  int nv = C / 4;
  static const bool stream_off = getenv("CVAR_LN_STREAM") != nullptr && getenv("CVAR_LN_STREAM")[0] == '0';   // A/B (diagnostic)
  if (y == nullptr && y16_hi != nullptr && !stream_off && ln_stream_usable(M, C) && ((uintptr_t)x & 15) == 0) {
    // large scales of the f16x3 path: persistent kernel, rows staged in shared memory by bulk copies (ln_stream.cu)
    int rc = launch_ln_stream(x, scale, shift, mod_row_stride, y16_hi, y16_lo, M, C, rows_per_sample, eps, s);
    if (rc) return rc;
    CVAR_CHECK_LAUNCH("cvar_ln_modulate");
    return 0;
  }
  if (nv <= 32 * 4)
    ln_modulate_kernel<4><<<grid, 256, 0, s>>>(x, scale, shift, mod_row_stride, y, y_lo, y16_hi, y16_lo, M, C,
                                               rows_per_sample, eps);
  else if (nv <= 32 * 8)
    ln_modulate_kernel<8><<<grid, 256, 0, s>>>(x, scale, shift, mod_row_stride, y, y_lo, y16_hi, y16_lo, M, C,
                                               rows_per_sample, eps);
  else if (nv <= 32 * 12)      // C = 1536 (depth 24): 12 float4 per lane, not 16 - fewer registers, a third CTA per SM
    ln_modulate_kernel<12><<<grid, 256, 0, s>>>(x, scale, shift, mod_row_stride, y, y_lo, y16_hi, y16_lo, M, C,
                                                rows_per_sample, eps);
  else
    ln_modulate_kernel<16><<<grid, 256, 0, s>>>(x, scale, shift, mod_row_stride, y, y_lo, y16_hi, y16_lo, M, C,
                                               rows_per_sample, eps);
  CVAR_CHECK_LAUNCH("cvar_ln_modulate");
  return 0;
}

// ------------------------------------------------------------------------------------------- layout / GN
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W,
                                    long long in_batch_stride, long long in_chan_stride, long long in_row_stride,
                                    long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  long long p = i / C;
  int xw = (int)(p % W);
  long long q = p / W;
  int yh = (int)(q % H);
  long long n = q / H;
  out[i] = in[n * in_batch_stride + (long long)c * in_chan_stride + (long long)yh * in_row_stride + xw];
}

extern "C" int cvar_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, long long in_batch_stride,
                                 void* stream) {
  // the f_hat halves of control_var.py:525-526 are views of a (B, C, 2H, W) tensor: channel stride 2*H*W
  long long chan_stride = in_batch_stride / C;
  long long total = (long long)B * C * H * W;
  nchw_to_nhwc_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, C, H, W, in_batch_stride,
                                                                          chan_stride, W, total);
  CVAR_CHECK_LAUNCH("cvar_nchw_to_nhwc");
  return 0;
}

__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W,
                                        int Cpad, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % Cpad);
  long long p = i / Cpad;
  int xw = (int)(p % W);
  long long q = p / W;
  int yh = (int)(q % H);
  long long n = q / H;
  out[i] = c < C ? in[((n * C + c) * H + yh) * W + xw] : 0.f;
}

extern "C" int cvar_nchw_to_nhwc_pad(const float* in, float* out, int B, int C, int H, int W, int Cpad, void* stream) {
  CVAR_REQUIRE(B > 0 && C > 0 && Cpad >= C && H > 0 && W > 0, "cvar_nchw_to_nhwc_pad: bad shape");
  long long total = (long long)B * Cpad * H * W;
  nchw_to_nhwc_pad_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, C, H, W, Cpad, total);
  CVAR_CHECK_LAUNCH("cvar_nchw_to_nhwc_pad");
  return 0;
}

static const int kGnChunkPixels = 256;
extern "C" int cvar_gn_chunks(int HW) { return (HW + kGnChunkPixels - 1) / kGnChunkPixels; }

// stage 1: per (chunk of pixels, sample): per-channel double sums, reduced to per-group partials.
__global__ void gn_partial_kernel(const float* __restrict__ x, double* __restrict__ scratch, int HW, int C, int groups,
                                  int chunks) {
  extern __shared__ double sm[];   // [2][C]
  int n = blockIdx.y, chunk = blockIdx.x;
  int p0 = chunk * kGnChunkPixels;
  int p1 = min(HW, p0 + kGnChunkPixels);
  int cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* base = x + ((long long)n * HW) * C + c;
    double s = 0.0, ss = 0.0;
    for (int p = p0; p < p1; ++p) {
      double v = (double)base[(long long)p * C];
      s += v;
      ss += v * v;
    }
    sm[c] = s;
    sm[C + c] = ss;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double s = 0.0, ss = 0.0;
    for (int j = 0; j < cpg; ++j) {
      s += sm[g * cpg + j];
      ss += sm[C + g * cpg + j];
    }
    long long o = (((long long)n * groups + g) * chunks + chunk) * 2;
    scratch[o] = s;
    scratch[o + 1] = ss;
  }
}

// stage 2: fold mean / rstd with the affine into per-(n,c) a, b.  One warp per (image, group): the lanes stride over the partials
// (2 048 per group at 256 x 256 from the conv epilogues), a fixed-order butterfly adds the 32 lane sums, then the lanes write the
// group's channels.  (Round 1: one thread per CHANNEL walked all partials serially - 68 us per launch, 39 launches per call.)
__global__ void gn_finalize_kernel(const double* __restrict__ scratch, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ a_out,
                                   float* __restrict__ b_out, int B, int HW, int C, int groups, int chunks, float eps) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)B * groups) return;
  const int n = (int)(wid / groups), g = (int)(wid % groups);
  const int cpg = C / groups;
  const double2* sp = reinterpret_cast<const double2*>(scratch) + ((long long)n * groups + g) * chunks;
  double s = 0.0, ss = 0.0;
  for (int k = lane; k < chunks; k += 32) {
    const double2 p = sp[k];
    s += p.x;
    ss += p.y;
  }
  s = warp_sum_d(s);
  ss = warp_sum_d(ss);
  const double cnt = (double)HW * cpg;
  const double mean = s / cnt;
  double var = ss / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = 1.0f / sqrtf((float)var + eps);
  for (int j = lane; j < cpg; j += 32) {
    const int c = g * cpg + j;
    const float a = rstd * gamma[c];
    a_out[(long long)n * C + c] = a;
    b_out[(long long)n * C + c] = beta[c] - (float)mean * a;
  }
}

extern "C" int cvar_gn_stats(const float* x_nhwc, const float* gamma, const float* beta, float* a_out, float* b_out,
                             double* scratch, int B, int HW, int C, int groups, float eps, void* stream) {
  CVAR_REQUIRE(C % groups == 0 && C <= 4096, "cvar_gn_stats: bad C=%d groups=%d", C, groups);
  int chunks = cvar_gn_chunks(HW);
  int threads = ((C + 31) / 32) * 32;
  if (threads > 640) threads = 640;
  gn_partial_kernel<<<dim3(chunks, B), threads, 2 * C * sizeof(double), (cudaStream_t)stream>>>(x_nhwc, scratch, HW, C,
                                                                                              groups, chunks);
  CVAR_CHECK_LAUNCH("cvar_gn_stats/partial");
  gn_finalize_kernel<<<cdiv((long long)B * groups, 8), 256, 0, (cudaStream_t)stream>>>(scratch, gamma, beta, a_out, b_out, B, HW,
                                                                                        C, groups, chunks, eps);
  CVAR_CHECK_LAUNCH("cvar_gn_stats/finalize");
  return 0;
}

// GroupNorm coefficients from partials a convolution's epilogue wrote (cvar_conv_args.gn_part): the second stage of
// cvar_gn_stats on (HW / 32) partials per (image, group) instead of cvar_gn_chunks(HW).  Fixed summation order.
extern "C" int cvar_gn_finalize_parts(const double* gn_part, const float* gamma, const float* beta, float* a_out, float* b_out,
                                      int B, int HW, int C, int groups, float eps, void* stream) {
  CVAR_REQUIRE(gn_part != nullptr && C % groups == 0 && HW % 32 == 0 && B > 0, "cvar_gn_finalize_parts: bad shape HW=%d C=%d", HW, C);
  gn_finalize_kernel<<<cdiv((long long)B * groups, 8), 256, 0, (cudaStream_t)stream>>>(gn_part, gamma, beta, a_out, b_out, B, HW,
                                                                                        C, groups, HW / 32, eps);
  CVAR_CHECK_LAUNCH("cvar_gn_finalize_parts");
  return 0;
}

// kFastSilu (FP16-pair output only, i.e. the operand of an f16x3 convolution): x * rcp(1 + ex2(-x log2 e)) instead of the IEEE
// division and the full-range expf - a few ulp, like the GELU of the dense-layer epilogue (gelu_tanh_fast), under the 2^-22 of
// the pair product that consumes it.  With the exact form the kernel was issue-bound (~65 instructions per element, 3.07 of 4
// issue slots, 0.72 of the copy bandwidth: profiles/r02_ln_affine.md).  The fp32-output path keeps the exact form.
template <bool kFastSilu>
__global__ void affine_nc_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                                 float* __restrict__ y, __half* __restrict__ y16_hi, __half* __restrict__ y16_lo,
                                 long long HWC4, int C4, long long total4, int silu) {
  // grid.y = sample, grid.x covers the sample's HW*C/4 float4s: no 64-bit division per thread
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= HWC4) return;
  const long long n = blockIdx.y;
  const long long i = n * HWC4 + j;
  const int c4 = (int)((unsigned)j % (unsigned)C4);
  float4 v = ld4(x + i * 4);
  float4 av = ld4(a + (n * C4 + c4) * 4), bv = ld4(b + (n * C4 + c4) * 4);
  v.x = fmaf(v.x, av.x, bv.x);
  v.y = fmaf(v.y, av.y, bv.y);
  v.z = fmaf(v.z, av.z, bv.z);
  v.w = fmaf(v.w, av.w, bv.w);
  if (silu) {
    if (kFastSilu) {
      v.x = __fdividef(v.x, 1.0f + __expf(-v.x));
      v.y = __fdividef(v.y, 1.0f + __expf(-v.y));
      v.z = __fdividef(v.z, 1.0f + __expf(-v.z));
      v.w = __fdividef(v.w, 1.0f + __expf(-v.w));
    } else {
      v.x = silu_f(v.x);
      v.y = silu_f(v.y);
      v.z = silu_f(v.z);
      v.w = silu_f(v.w);
    }
  }
  if (y16_hi != nullptr) {
    const float vv[4] = {v.x, v.y, v.z, v.w};
    st4_split_f16(y16_hi + i * 4, y16_lo + i * 4, vv);
  }
  if (y != nullptr) st4(y + i * 4, v);
}

// nearest x2 upsample + FP16-pair split: one thread per 4 channels of a SOURCE pixel, four destination pixels
__global__ void upsample2x_split_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo,
                                        int H, int W, int C4, long long total4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = (int)(i % C4);
  long long p = i / C4;
  const int xw = (int)(p % W);
  p /= W;
  const int yh = (int)(p % H);
  const long long n = p / H;
  const float4 v = ld4(x + i * 4);
  const float vv[4] = {v.x, v.y, v.z, v.w};
  __half h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_f16(vv[j], h[j], l[j]);
  uint2 ph, pl;
  ph.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
  ph.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
  pl.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
  pl.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const long long o = (((n * (2 * H) + (2 * yh + dy)) * (2 * W) + (2 * xw + dx)) * C4 + c4) * 4;
      *reinterpret_cast<uint2*>(hi + o) = ph;
      *reinterpret_cast<uint2*>(lo + o) = pl;
    }
}

extern "C" int cvar_upsample2x_split_f16(const float* x_nhwc, void* hi, void* lo, int B, int H, int W, int C, void* stream) {
  CVAR_REQUIRE(C % 4 == 0 && B > 0 && H > 0 && W > 0, "cvar_upsample2x_split_f16: bad shape");
  CVAR_REQUIRE(x_nhwc && hi && lo, "cvar_upsample2x_split_f16: null pointer");
  const long long total4 = (long long)B * H * W * C / 4;
  upsample2x_split_kernel<<<cdiv(total4, 256), 256, 0, (cudaStream_t)stream>>>(
      x_nhwc, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), H, W, C / 4, total4);
  CVAR_CHECK_LAUNCH("cvar_upsample2x_split_f16");
  return 0;
}

extern "C" int cvar_affine_nc(const float* x_nhwc, const float* a, const float* b, float* y, void* y16_hi, void* y16_lo,
                              int B, int HW, int C, int silu, void* stream) {
  CVAR_REQUIRE(C % 4 == 0, "cvar_affine_nc: C %% 4 != 0");
  CVAR_REQUIRE(y != nullptr || y16_hi != nullptr, "cvar_affine_nc: no output");
  CVAR_REQUIRE((y16_hi == nullptr) == (y16_lo == nullptr), "cvar_affine_nc: y16_hi/y16_lo must come together");
  long long total4 = (long long)B * HW * C / 4;
  CVAR_REQUIRE(B <= 65535 && (long long)HW * C / 4 < (1LL << 31), "cvar_affine_nc: shape too large");
  static const bool exact_silu = getenv("CVAR_EXACT_SILU") != nullptr && getenv("CVAR_EXACT_SILU")[0] == '1';   // A/B (diagnostic)
  auto kern = (y == nullptr && !exact_silu) ? affine_nc_kernel<true> : affine_nc_kernel<false>;
  kern<<<dim3(cdiv((long long)HW * C / 4, 256), B), 256, 0, (cudaStream_t)stream>>>(
      x_nhwc, a, b, y, reinterpret_cast<__half*>(y16_hi), reinterpret_cast<__half*>(y16_lo), (long long)HW * C / 4, C / 4,
      total4, silu);
  CVAR_CHECK_LAUNCH("cvar_affine_nc");
  return 0;
}

// row softmax, one warp per row (AttnBlock: 256 columns)
__global__ void softmax_rows_kernel(float* __restrict__ x, int rows, int cols) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long r = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= rows) return;
  float* xr = x + r * cols;
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, xr[c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) {
    float e = expf(xr[c] - mx);
    xr[c] = e;
    s += e;
  }
  s = warp_sum(s);
  for (int c = lane; c < cols; c += 32) xr[c] = xr[c] / s;
}

extern "C" int cvar_softmax_rows(float* x, int rows, int cols, void* stream) {
  softmax_rows_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, cols);
  CVAR_CHECK_LAUNCH("cvar_softmax_rows");
  return 0;
}

__global__ void repack_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ o, int Cout, int Cin,
                                          int ks) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)Cout * Cin * ks * ks;
  if (i >= total) return;
  int ci = (int)(i % Cin);
  long long r = i / Cin;
  int tap = (int)(r % (ks * ks));
  int co = (int)(r / (ks * ks));
  o[i] = w[((long long)co * Cin + ci) * ks * ks + tap];
}

extern "C" int cvar_repack_conv_weight(const float* w_oihw, float* w_out, int Cout, int Cin, int ks, void* stream) {
  long long total = (long long)Cout * Cin * ks * ks;
  repack_conv_weight_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, w_out, Cout, Cin, ks);
  CVAR_CHECK_LAUNCH("cvar_repack_conv_weight");
  return 0;
}

__global__ void repack_conv_weight_pad_kernel(const float* __restrict__ w, float* __restrict__ o, int Cout, int Cin,
                                              int ks, int Cin_pad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)Cout * Cin_pad * ks * ks;
  if (i >= total) return;
  int ci = (int)(i % Cin_pad);
  long long r = i / Cin_pad;
  int tap = (int)(r % (ks * ks));
  int co = (int)(r / (ks * ks));
  o[i] = ci < Cin ? w[((long long)co * Cin + ci) * ks * ks + tap] : 0.f;
}

extern "C" int cvar_repack_conv_weight_pad(const float* w_oihw, float* w_out, int Cout, int Cin, int ks, int Cin_pad,
                                           void* stream) {
  CVAR_REQUIRE(Cout > 0 && Cin > 0 && Cin_pad >= Cin && ks > 0, "cvar_repack_conv_weight_pad: bad shape");
  long long total = (long long)Cout * Cin_pad * ks * ks;
  repack_conv_weight_pad_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, w_out, Cout, Cin, ks,
                                                                                     Cin_pad);
  CVAR_CHECK_LAUNCH("cvar_repack_conv_weight_pad");
  return 0;
}

// ------------------------------------------------------------------------------------------- FP16-pair split
__global__ void split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = cvar::ld4(x + i * 4);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    cvar::st4_split_f16(hi + i * 4, lo + i * 4, vv);
  }
}

extern "C" int cvar_split_f16(const float* x, void* hi, void* lo, long long n, void* stream) {
  CVAR_REQUIRE(n > 0 && n % 4 == 0, "cvar_split_f16: n must be a positive multiple of 4");
  CVAR_REQUIRE(x && hi && lo, "cvar_split_f16: null pointer");
  const long long n4 = n / 4;
  const int blocks = (int)std::min<long long>((n4 + 255) / 256, 148LL * 16);
  split_f16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), n4);
  CVAR_CHECK_LAUNCH("cvar_split_f16");
  return 0;
}
