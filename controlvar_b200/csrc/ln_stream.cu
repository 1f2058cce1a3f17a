// AdaLN LayerNorm + modulation for the large scales, as a persistent streaming kernel (basic_var.py:207-208, 232-233).
//
// Round 1's kernel (misc.cu: one warp per row, the row fetched by 12 vector loads per lane, 8 rows per CTA) reached 3.8 TB/s
// = 58 % of the measured copy bandwidth at the last scale.  What held it there was not the row traffic but the MODULATION
// operands: every row re-read its sample's scale and shift vectors (2 x 6 KB at d24, twice the row itself) through L1 / L2
// (measured with this kernel: 284 us with those loads, 150 us with constants in their place, profiles/r02_ln_affine.md).
// Here one CTA per SM stays resident and walks a CONTIGUOUS range of rows, so it changes sample once or twice in its life:
// the sample's scale / shift sit in shared memory (loaded at a sample change, behind a CTA barrier).  Every warp owns a
// private ring of `slots` row buffers which it fills with ONE bulk copy per row (cp.async.bulk, completion on an mbarrier:
// loads in flight hold no registers) and refills as soon as the row is in registers.  The arithmetic - element-to-lane
// mapping, order of every sum, every rounding - is the round-1 kernel's: the output is bit-identical to it (tested).
#include "tc_ptx.cuh"

namespace cvar {
namespace {
using namespace tc;

__device__ __forceinline__ void bulk_load_row(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// The rows of one warp, in order: within the CTA's range [r0, r1), sample segment by sample segment (a segment = the part of
// one sample's rows inside the range), rows lo + warp, lo + warp + W, ... of each segment.
struct RowWalk {
  long long m, seg_hi, r1;
  int l, w, W;
  __device__ void start(long long r0, long long r1_, int l_, int w_, int W_) {
    r1 = r1_, l = l_, w = w_, W = W_;
    seg_hi = min(r1, (r0 / l + 1) * (long long)l);
    m = r0 + w;
    settle();
  }
  __device__ void settle() {
    while (m >= seg_hi && seg_hi < r1) {
      m = seg_hi + w;
      seg_hi = min(r1, seg_hi + l);
    }
  }
  __device__ bool valid() const { return m < seg_hi; }
  __device__ void next() {
    m += W;
    settle();
  }
};

template <int MAXV>   // float4 per lane: C <= 128 * MAXV
__global__ void __launch_bounds__(512) ln_modulate_stream_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift, long long mod_stride,
                                                                 __half* __restrict__ y16_hi, __half* __restrict__ y16_lo,
                                                                 int M, int C, int rows_per_sample, float eps, int slots,
                                                                 int rows_per_cta) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  const uint32_t row_bytes = (uint32_t)C * 4u;
  const int nv = C >> 2;
  float* mod = reinterpret_cast<float*>(smem);                       // [2][C]: scale, shift of the current sample
  float* ring = mod + 2 * C + (size_t)warp * slots * C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)(2 + W * slots) * row_bytes) + warp * slots;
  if (lane == 0) {
    for (int s = 0; s < slots; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncwarp();
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min((long long)M, r0 + rows_per_cta);
  RowWalk ahead;                                                     // the row the next bulk copy fetches
  ahead.start(r0, r1, rows_per_sample, warp, W);
  if (lane == 0) {
    for (int s = 0; s < slots && ahead.valid(); ++s, ahead.next()) {
      mbar_arrive_expect_tx(&bars[s], row_bytes);
      bulk_load_row(ring + (size_t)s * C, x + ahead.m * C, row_bytes, &bars[s]);
    }
  }
  int slot = 0;
  uint32_t phase = 0;
  for (long long seg_lo = r0; seg_lo < r1;) {
    const long long r = seg_lo / rows_per_sample;
    const long long seg_hi = min(r1, (r + 1) * (long long)rows_per_sample);
    __syncthreads();                                                 // every warp is done with the previous sample's vectors
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      st4(mod + i * 4, ld4(scale + r * mod_stride + i * 4));
      st4(mod + C + i * 4, ld4(shift + r * mod_stride + i * 4));
    }
    __syncthreads();
    for (long long m = seg_lo + warp; m < seg_hi; m += W) {
      mbar_wait(&bars[slot], phase);
      const float* xr = ring + (size_t)slot * C;
      float4 v[MAXV];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int idx = lane + i * 32;
        if (idx < nv) {
          v[i] = ld4(xr + idx * 4);
          s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
      }
      const float mean = warp_sum(s) / (float)C;
      // every lane's reads of the slot have returned (the butterfly sum consumed them, and lane 0's result depends on all of
      // them); __syncwarp orders them before lane 0's refill in the memory model as well.  (compute-sanitizer racecheck still
      // reports the bulk copy against these reads: it does not follow mbarrier / async-proxy completion.)
      __syncwarp();
      if (lane == 0 && ahead.valid()) {
        mbar_arrive_expect_tx(&bars[slot], row_bytes);
        bulk_load_row(ring + (size_t)slot * C, x + ahead.m * C, row_bytes, &bars[slot]);
        ahead.next();
      }
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int idx = lane + i * 32;
        if (idx < nv) {
          const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
          ss += (a * a + b * b) + (c * c + d * d);
        }
      }
      const float var = warp_sum(ss) / (float)C;
      const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int idx = lane + i * 32;
        if (idx < nv) {
          const float4 a = ld4(mod + idx * 4), b = ld4(mod + C + idx * 4);
          // ln(x).mul(scale.add(1)).add_(shift): three separately rounded steps          basic_var.py:208
          const float ov[4] = {__fadd_rn(__fmul_rn(__fmul_rn(v[i].x - mean, rstd), __fadd_rn(a.x, 1.f)), b.x),
                               __fadd_rn(__fmul_rn(__fmul_rn(v[i].y - mean, rstd), __fadd_rn(a.y, 1.f)), b.y),
                               __fadd_rn(__fmul_rn(__fmul_rn(v[i].z - mean, rstd), __fadd_rn(a.z, 1.f)), b.z),
                               __fadd_rn(__fmul_rn(__fmul_rn(v[i].w - mean, rstd), __fadd_rn(a.w, 1.f)), b.w)};
          st4_split_f16(y16_hi + m * C + idx * 4, y16_lo + m * C + idx * 4, ov);
        }
      }
      if (++slot == slots) {
        slot = 0;
        phase ^= 1u;
      }
    }
    seg_lo = seg_hi;
  }
}

constexpr int kLnSmemBudget = 200 * 1024;
constexpr int kLnMinRows = 2 * 148 * 16;
}  // namespace

// Does the streaming kernel take this call?  (FP16-pair output only, enough rows to fill the machine, >= 2 ring slots.)
bool ln_stream_usable(int M, int C) {
  if (C % 4 != 0 || C > 2048 || M < kLnMinRows) return false;
  return (kLnSmemBudget - 2 * C * 4) / (12 * C * 4) >= 2;
}

int launch_ln_stream(const float* x, const float* scale, const float* shift, long long mod_stride, __half* y16_hi,
                     __half* y16_lo, int M, int C, int rows_per_sample, float eps, cudaStream_t s) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int ring_budget = kLnSmemBudget - 2 * C * 4;
  int W = 16;
  int slots = ring_budget / (W * C * 4);
  if (slots < 2) {
    W = 12;
    slots = ring_budget / (W * C * 4);
  }
  if (slots > 4) slots = 4;
  const int smem = (2 + W * slots) * C * 4 + W * slots * 8;
  const int nv = C / 4;
  auto kern = nv <= 32 * 4 ? ln_modulate_stream_kernel<4>
              : nv <= 32 * 8 ? ln_modulate_stream_kernel<8>
              : nv <= 32 * 12 ? ln_modulate_stream_kernel<12>
                              : ln_modulate_stream_kernel<16>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    set_error("cvar_ln_modulate: cannot raise shared memory to %d: %s", smem, cudaGetErrorString(e));
    return -2;
  }
  const int rows_per_cta = cdiv(M, sms);
  const int grid = cdiv(M, rows_per_cta);
  kern<<<grid, W * 32, smem, s>>>(x, scale, shift, mod_stride, y16_hi, y16_lo, M, C, rows_per_sample, eps, slots,
                                  rows_per_cta);
  return 0;
}
}  // namespace cvar
