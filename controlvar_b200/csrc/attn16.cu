// KV-cached attention on FP16-pair operands (engine 4): basic_var.py:106-117, same contract as attn.cu.
//
// q, K and V^T arrive as FP16 pairs written by cvar_qkv_project16 (V^T: x ~= hi + lo * 2^-11, cvar_split_f16; q, K:
// 16 x = hi + lo), i.e. half the bytes of the TF32 hi/lo split the first tensor-core kernel used, and the products run
// as three kind::f16 MMAs (hi*hi, hi*lo, lo*hi) at twice the TF32 tensor rate.  Design of the tensor-core kernel (attn16_tc_kernel):
//   * CTA = 128 queries of one (row, head), 192 threads, and TWO CTAs per SM: everything is sized to half an SM
//     (256 TMEM columns, 97 KB shared memory, <= 168 registers), so the softmax of one CTA overlaps the MMAs of the other.
//     The measured limiter of the TF32 kernel was the softmax warpgroup (~3100 cycles per 64-key tile against 1536 of
//     MMA, profiles/r01_attn_trace.md), and its 448 TMEM columns / 255 registers allowed one CTA per SM only.
//   * Q tile (128 x 64, hi and lo) and the K / V^T tiles (64 x 64) are fetched by TMA straight from the pair arrays:
//     one 128-byte-swizzled block each, no operand conversion anywhere.
//   * q and K are "qk pairs" (16 x = hi + lo, un-scaled residual, common.cuh), so S = Q K^T needs ONE accumulator and
//     the 256 TMEM columns hold TWO S buffers: S(j+1) and S(j+2) are issued ahead and the softmax warpgroup never waits
//     for the tensor core in steady state (v2, with a main + cross S accumulator and a single buffer, spent 30 % of its
//     warp samples in the s_full poll loop: profiles/r01_attn16.md).  The softmax thread (one per query row) reads S in
//     16-column chunks and writes P back as a standard FP16 pair OVER the cells it has just consumed (chunk c: hi in
//     columns [16c, 16c+8), lo in [16c+8, 16c+16), two keys per 32-bit cell); P @ V reads its A operand from there.
//     tcgen05.mma executes in issue order, so S(j+2), issued after P @ V(j), may overwrite those cells.
//   * O accumulates in TMEM for kDrain = 4 tiles (48 truncating MMA steps) and is then added into the running output
//     row in registers with round-to-nearest adds.  The exponentials are taken against a reference that is the first tile's
//     row maximum and moves only when a later tile's maximum exceeds it by 2^8 (then that row rescales what it has
//     accumulated, in registers and in its TMEM row), so the common tile does no rescaling work at all.  Two 16-bit values
//     per TMEM cell, the lower key index in the low half (A operand of kind::f16 from tensor memory).
//   TMEM columns: S/P buffer 0 [0,64)  S/P buffer 1 [64,128)  O_main [128,192)  O_x [192,256)
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "tc_ptx.cuh"

using namespace cvar;

namespace {
constexpr float kInvLo = 1.0f / kF16LoScale;

__device__ __forceinline__ float pair_val(__half h, __half l) { return fmaf(__half2float(l), kInvLo, __half2float(h)); }
// four consecutive elements of a pair array (8-byte aligned)
__device__ __forceinline__ float4 ld4_pair(const __half* hi, const __half* lo, long long off) {
  const uint2 a = *reinterpret_cast<const uint2*>(hi + off), b = *reinterpret_cast<const uint2*>(lo + off);
  const __half2 a0 = *reinterpret_cast<const __half2*>(&a.x), a1 = *reinterpret_cast<const __half2*>(&a.y);
  const __half2 b0 = *reinterpret_cast<const __half2*>(&b.x), b1 = *reinterpret_cast<const __half2*>(&b.y);
  return make_float4(pair_val(__low2half(a0), __low2half(b0)), pair_val(__high2half(a0), __high2half(b0)),
                     pair_val(__low2half(a1), __low2half(b1)), pair_val(__high2half(a1), __high2half(b1)));
}

// four consecutive elements of a qk pair array (16 x = hi + lo, split_f16_qk)
__device__ __forceinline__ float4 ld4_qk(const __half* hi, const __half* lo, long long off) {
  const uint2 a = *reinterpret_cast<const uint2*>(hi + off), b = *reinterpret_cast<const uint2*>(lo + off);
  const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
  const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&b.x)), b1 = __half22float2(*reinterpret_cast<const __half2*>(&b.y));
  constexpr float inv = 1.0f / kQkScale;
  return make_float4((a0.x + b0.x) * inv, (a0.y + b0.y) * inv, (a1.x + b1.x) * inv, (a1.y + b1.y) * inv);
}

// ---------------------------------------------------------------------------------------------------- SIMT kernel
// The fp32 flash-style kernel of attn.cu reading the pair format: serves l < 64 and is the cross-check of the
// tensor-core kernel.  One CTA = 64 queries of one (row, head).
constexpr int BQ = 64, BKV = 64, D = 64;
constexpr int PS = 68;

struct AttnSmem {
  float Qt[D][BQ];
  float Kt[D][BKV];
  float V[BKV][D];
  float Pt[BKV][PS];
};

__global__ void __launch_bounds__(256)
attn16_simt_kernel(const __half* __restrict__ q_hi, const __half* __restrict__ q_lo, const __half* __restrict__ k_hi,
                   const __half* __restrict__ k_lo, const __half* __restrict__ vt_hi, const __half* __restrict__ vt_lo,
                   float* __restrict__ out, __half* __restrict__ o16_hi, __half* __restrict__ o16_lo, int H, int l, int L,
                   int T_max, float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, r = blockIdx.z;
  const long long rh = (long long)r * H + h;
  const long long q_off = rh * l * D;
  const long long kv_off = rh * T_max * D;

  for (int it = 0; it < 4; ++it) {
    int item = it * 256 + tid;
    int i = item & 63, dq = item >> 6;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + i < l) v = ld4_qk(q_hi, q_lo, q_off + (long long)(q0 + i) * D + dq * 4);
    sm.Qt[dq * 4 + 0][i] = v.x * scale;
    sm.Qt[dq * 4 + 1][i] = v.y * scale;
    sm.Qt[dq * 4 + 2][i] = v.z * scale;
    sm.Qt[dq * 4 + 3][i] = v.w * scale;
  }

  float o[4][4];
  float mrow[4], lrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mrow[i] = -INFINITY;
    lrow[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }

  for (int k0 = 0; k0 < L; k0 += BKV) {
    __syncthreads();
    for (int it = 0; it < 4; ++it) {
      int item = it * 256 + tid;
      int j = item & 63, dq = item >> 6;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + j < L) kv = ld4_qk(k_hi, k_lo, kv_off + (long long)(k0 + j) * D + dq * 4);
      sm.Kt[dq * 4 + 0][j] = kv.x;
      sm.Kt[dq * 4 + 1][j] = kv.y;
      sm.Kt[dq * 4 + 2][j] = kv.z;
      sm.Kt[dq * 4 + 3][j] = kv.w;
    }
    for (int it = 0; it < 4; ++it) {
      int item = it * 256 + tid;
      int d = item >> 4, jq = item & 15;
      const long long base = kv_off + (long long)d * T_max + k0 + jq * 4;
      float vv[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + jq * 4 + 3 < T_max) {          // T_max % 8 == 0 and k0 % 64 == 0: the 8-byte load is aligned
        float4 t4 = ld4_pair(vt_hi, vt_lo, base);
        vv[0] = t4.x, vv[1] = t4.y, vv[2] = t4.z, vv[3] = t4.w;
      } else {
        for (int i = 0; i < 4; ++i)
          if (k0 + jq * 4 + i < T_max) vv[i] = pair_val(vt_hi[base + i], vt_lo[base + i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int j = jq * 4 + i;
        sm.V[j][(((d >> 2) ^ (j & 15)) << 2) + (d & 3)] = (k0 + j < L) ? vv[i] : 0.f;
      }
    }
    __syncthreads();

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
      float4 a = *reinterpret_cast<const float4*>(&sm.Qt[d][ty * 4]);
      float b0 = sm.Kt[d][tx], b1 = sm.Kt[d][16 + tx], b2 = sm.Kt[d][32 + tx], b3 = sm.Kt[d][48 + tx];
      s[0][0] = fmaf(a.x, b0, s[0][0]), s[0][1] = fmaf(a.x, b1, s[0][1]), s[0][2] = fmaf(a.x, b2, s[0][2]), s[0][3] = fmaf(a.x, b3, s[0][3]);
      s[1][0] = fmaf(a.y, b0, s[1][0]), s[1][1] = fmaf(a.y, b1, s[1][1]), s[1][2] = fmaf(a.y, b2, s[1][2]), s[1][3] = fmaf(a.y, b3, s[1][3]);
      s[2][0] = fmaf(a.z, b0, s[2][0]), s[2][1] = fmaf(a.z, b1, s[2][1]), s[2][2] = fmaf(a.z, b2, s[2][2]), s[2][3] = fmaf(a.z, b3, s[2][3]);
      s[3][0] = fmaf(a.w, b0, s[3][0]), s[3][1] = fmaf(a.w, b1, s[3][1]), s[3][2] = fmaf(a.w, b2, s[3][2]), s[3][3] = fmaf(a.w, b3, s[3][3]);
    }
    float p[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (k0 + j * 16 + tx >= L) s[i][j] = -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      float mnew = fmaxf(mrow[i], mx);
      float corr = expf(mrow[i] - mnew);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p[i][j] = expf(s[i][j] - mnew);
        rs += p[i][j];
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      lrow[i] = lrow[i] * corr + rs;
      mrow[i] = mnew;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(&sm.Pt[j * 16 + tx][ty * 4]) = make_float4(p[0][j], p[1][j], p[2][j], p[3][j]);
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < BKV; ++j) {
      float4 a = *reinterpret_cast<const float4*>(&sm.Pt[j][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&sm.V[j][(tx ^ (j & 15)) << 2]);
      o[0][0] = fmaf(a.x, b.x, o[0][0]), o[0][1] = fmaf(a.x, b.y, o[0][1]), o[0][2] = fmaf(a.x, b.z, o[0][2]), o[0][3] = fmaf(a.x, b.w, o[0][3]);
      o[1][0] = fmaf(a.y, b.x, o[1][0]), o[1][1] = fmaf(a.y, b.y, o[1][1]), o[1][2] = fmaf(a.y, b.z, o[1][2]), o[1][3] = fmaf(a.y, b.w, o[1][3]);
      o[2][0] = fmaf(a.z, b.x, o[2][0]), o[2][1] = fmaf(a.z, b.y, o[2][1]), o[2][2] = fmaf(a.z, b.z, o[2][2]), o[2][3] = fmaf(a.z, b.w, o[2][3]);
      o[3][0] = fmaf(a.w, b.x, o[3][0]), o[3][1] = fmaf(a.w, b.y, o[3][1]), o[3][2] = fmaf(a.w, b.z, o[3][2]), o[3][3] = fmaf(a.w, b.w, o[3][3]);
    }
  }

  const int C = H * D;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int t = q0 + ty * 4 + i;
    if (t < l) {
      float inv = 1.0f / lrow[i];
      const float vv[4] = {o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv};
      const long long off = ((long long)r * l + t) * C + h * D + tx * 4;
      if (o16_hi != nullptr) st4_split_f16(o16_hi + off, o16_lo + off, vv);
      if (out != nullptr) st4(out + off, make_float4(vv[0], vv[1], vv[2], vv[3]));
    }
  }
}
}  // namespace

// ============================================================================================== tensor-core kernel
namespace tcattn16 {
using namespace cvar::tc;
constexpr int BQ = 128, BKV = 64, D = 64;
constexpr int kThreads = 192;
constexpr int kTile = BKV * 128;                  // 64 rows x 128 B: one K or V^T tile half (hi or lo), 8 KB
constexpr int kQTile = BQ * 128;                  // 128 rows x 128 B: Q hi or lo, 16 KB
constexpr int kStageBytes = 4 * kTile;            // K_hi | K_lo | VT_hi | VT_lo = 32 KB (K and V^T halves have separate barriers)
constexpr int kStages = 2;
constexpr int kSmem = 1024 + 2 * kQTile + kStages * kStageBytes + 256;     // 99,584 B: two CTAs per SM
// S has ONE accumulator (q, k are qk pairs) and two buffers: S(j) lives in buffer j & 1 at column 64 * (j & 1)
constexpr uint32_t kColS = 0, kColO = 128, kColOx = 192, kTmemCols = 256;
// O stays in TMEM for kDrain tiles (kDrain * 12 truncating accumulation steps) before it is added into the fp32 registers
constexpr int kDrain = 4;
// the softmax reference moves when a logit exceeds it by more than 2^kRebase (p <= 256: far inside fp16 / fp32 range)
constexpr float kRebase = 8.0f;
// D fp32 (bit 4), A / B format 0 = F16, both K-major, N = 64, M = 128
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// A and B from shared memory
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// A from tensor memory (two halves per 32-bit cell, 8 cells per 16-wide k-step), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// TMEM loads WITHOUT the wait: the destination registers are valid only after tmem_wait_ld()
__device__ __forceinline__ void tmem_ld_nowait_x16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_nowait_x32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// Compiler-level dependency: values loaded by a *_nowait load may only be used after the wait.  The wait is a volatile asm
// without register operands, so plain arithmetic on the loaded registers could legally be scheduled above it; these empty
// volatile asms (ordered after the wait) re-define the registers and pin every use behind it.
__device__ __forceinline__ void reg_fence16(float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                    "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}
__device__ __forceinline__ void reg_fence32(float* v) {
  reg_fence16(v);
  reg_fence16(v + 16);
}
__device__ __forceinline__ void tmem_alloc_n(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
// Optional phase trace (diagnostics, cvar_debug_set_attn_trace): CTA (0,0,0) stamps clock64() per KV tile j < 32.
// trace[(who * 32 + j) * 8 + ev]:
//   who 0 = softmax thread 0: 0 s_full(j) seen, 1 row maximum known, 2 P(j) stored, 3 p_ready(j) signalled
//   who 1 = MMA thread:       0 v_full + p_ready(j) seen, 1 P @ V(j) issued, 2 k_full(j+2) seen, 3 S(j+2) issued
//   who 2 = TMA thread:       0 k_empty seen for tile j (K load issued), 1 v_empty seen for tile j (V^T load issued)
__device__ long long* g_attn16_trace = nullptr;
__device__ __forceinline__ void astamp(int who, int j, int ev) {
  if (g_attn16_trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j >= 0 && j < 32)
    g_attn16_trace[(who * 32 + j) * 8 + ev] = clock64();
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// kFast (cvar_set_fast_mode; NOT a parity mode): hi halves only - one MMA per product, no P residual, half the K / V^T bytes.
// Block-causal full-sequence pass (ControlVAR.forward, control_var.py:158-198, 622-636: a query of scale s sees the keys of
// scales <= s): the mask is a step function of the query's scale, so it needs no L x L tensor - the launch carries a table
// of query tiles {first query, number of queries, number of visible keys}, one CTA column per tile; a tile never straddles
// two scales.  n == 0: the plain KV-cache launch (tile bx = queries [128 bx, 128 bx + 128), all L keys).
constexpr int kMaxSegTiles = 40;
struct AttnSegs {
  int n;
  int q0[kMaxSegTiles];
  int nq[kMaxSegTiles];
  int L[kMaxSegTiles];
};

// kOnePass: S(j) is read from tensor memory ONCE (64 registers) and the P pass runs out of those registers; the round-1
// kernel read it twice (row maximum, then 16-column chunks).  Tensor-memory reads are ~95 B/clk per SM (measured on the GEMM
// drain, profiles/r02_gemm_epilogue.md): the second read of the 32 KB S tile cost ~340 cycles of that port per CTA and
// tile, plus four exposed load round trips.
template <bool kFast, bool kOnePass>
__global__ void __launch_bounds__(kThreads, 2)
attn16_tc_kernel(const __grid_constant__ CUtensorMap mapQhi, const __grid_constant__ CUtensorMap mapQlo,
                 const __grid_constant__ CUtensorMap mapKhi, const __grid_constant__ CUtensorMap mapKlo,
                 const __grid_constant__ CUtensorMap mapVhi, const __grid_constant__ CUtensorMap mapVlo,
                 float* __restrict__ out, __half* __restrict__ o16_hi, __half* __restrict__ o16_lo, int H, int l, int L_all,
                 float scale, const __grid_constant__ AttnSegs segs) {
  using G = Geo<32>;          // 128-byte rows, SWIZZLE_128B: 64 halves per row
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* q_s = smem;                                        // Q_hi | Q_lo
  auto stage = [&](int s) { return smem + 2 * kQTile + s * kStageBytes; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kQTile + kStages * kStageBytes);
  uint64_t* k_full = bars;                   // [kStages]  K tile landed
  uint64_t* k_empty = bars + kStages;        // [kStages]  S(j) has read it
  uint64_t* v_full = bars + 2 * kStages;     // [kStages]  V^T tile landed
  uint64_t* v_empty = bars + 3 * kStages;    // [kStages]  P @ V(j) has read it
  uint64_t* s_full = bars + 4 * kStages;     // [2]        S(j) accumulated in buffer j & 1
  uint64_t* q_full = bars + 4 * kStages + 2;
  // P(j) in TMEM (and O drained when due), by all 128 softmax threads.  TWO barriers, indexed by the S buffer: with S
  // issued ahead the softmax can finish tile j+1 before the MMA thread has looked at tile j, and a single barrier would
  // then be two phases ahead of its waiter - whose parity test can no longer tell "done" from "not yet".
  uint64_t* p_ready = bars + 4 * kStages + 3;// [2]
  // P @ V(j) accumulated: also two barriers (j & 1).  A softmax thread that waits for P @ V(j-1) only knows that
  // P @ V(j-3) has completed (S(j), which it has consumed, was issued behind P @ V(j-2) and completes in order); on a
  // single barrier that is two phases back and the parity test would pass at once.
  uint64_t* o_full = bars + 4 * kStages + 5; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * kStages + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, r = blockIdx.z;
  const int q0 = segs.n > 0 ? segs.q0[blockIdx.x] : blockIdx.x * BQ;
  const int q_end = segs.n > 0 ? q0 + segs.nq[blockIdx.x] : l;          // queries [q0, q_end) are this tile's to store
  const int L = segs.n > 0 ? segs.L[blockIdx.x] : L_all;                // keys this tile sees
  const int rh = r * H + h;
  const int ntiles = (L + BKV - 1) / BKV;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&mapQhi), tma_prefetch_desc(&mapQlo);
    tma_prefetch_desc(&mapKhi), tma_prefetch_desc(&mapKlo), tma_prefetch_desc(&mapVhi), tma_prefetch_desc(&mapVlo);
    for (int s = 0; s < kStages; ++s)
      mbar_init(&k_full[s], 1), mbar_init(&k_empty[s], 1), mbar_init(&v_full[s], 1), mbar_init(&v_empty[s], 1);
    mbar_init(q_full, 1);
    mbar_init(&s_full[0], 1), mbar_init(&s_full[1], 1);
    mbar_init(&p_ready[0], 128), mbar_init(&p_ready[1], 128);
    mbar_init(&o_full[0], 1), mbar_init(&o_full[1], 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc_n(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ================================================================ softmax + output rows
    // Budget (ncu of the first version, profiles/r01_attn16.md): 1705 instructions per warp and tile, the XU pipe (MUFU
    // and scalar F2F conversions, 8 cycles per warp instruction) 45 % busy, branches from per-element masking.  Hence:
    // packed conversions (F2FP, not XU), masking only in the last tile, a lazily updated reference instead of the exact
    // running maximum (no per-tile rescale), and O left accumulating in TMEM for kDrain tiles.
    const int row = threadIdx.x;
    const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
    // p = exp((s - ref) * scale) = 2^((s - ref) * scale * log2 e): the difference is formed first, so the rounding of the
    // product is a relative error of |t| * 2^-24 on a term of weight e^t
    // S holds 256 q.k (both operands are qk pairs, 16 x = hi + lo): the factor is folded into the exponent scale
    const float sl2 = scale * 1.4426950408889634f * (1.0f / (kQkScale * kQkScale));
    float o_reg[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o_reg[d] = 0.f;
    float m_ref = 0.f, l_run = 0.f;

    // one 16-key chunk: p for the chunk, written back over the S_main cells it came from as an FP16 pair
    // nref = -ref * scale * log2e; t = fma(s, sl2, nref): the product is exact inside the FMA, and the rounding of nref
    // is a factor common to every key of the row (it cancels in the normalisation)
    auto chunk = [&](auto masked, uint32_t sb, int c, const float* a, float nref, int nvalid, float& rs) {
      uint32_t ph[8], pl[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float p0 = ex2_approx(fmaf(a[i], sl2, nref)), p1 = ex2_approx(fmaf(a[i + 1], sl2, nref));
        if constexpr (decltype(masked)::value) {   // last tile only: keys past L contribute nothing
          if (16 * c + i >= nvalid) p0 = 0.f;
          if (16 * c + i + 1 >= nvalid) p1 = 0.f;
        }
        rs += p0 + p1;
        const __half2 h = __floats2half2_rn(p0, p1);                 // low half = lower key index
        ph[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        if (!kFast) {
          const float2 hf = __half22float2(h);
          const __half2 lo = __floats2half2_rn((p0 - hf.x) * kF16LoScale, (p1 - hf.y) * kF16LoScale);
          pl[i >> 1] = *reinterpret_cast<const uint32_t*>(&lo);
        }
      }
      tmem_st_32x32b_x8(tl + sb + 16 * c, ph);
      if (!kFast) tmem_st_32x32b_x8(tl + sb + 16 * c + 8, pl);
    };
    // the whole tile against reference `ref`; chunk loads are prefetched one ahead (one exposed TMEM round trip)
    auto tile_pass = [&](auto masked, uint32_t sb, float nref, int nvalid, float& rs) {
      float a0[16], a1[16];
      tmem_ld_nowait_x16(tl + sb, a0);
      tmem_wait_ld();
      reg_fence16(a0);
      tmem_ld_nowait_x16(tl + sb + 16, a1);
      chunk(masked, sb, 0, a0, nref, nvalid, rs);
      tmem_wait_ld();
      reg_fence16(a1);
      tmem_ld_nowait_x16(tl + sb + 32, a0);
      chunk(masked, sb, 1, a1, nref, nvalid, rs);
      tmem_wait_ld();
      reg_fence16(a0);
      tmem_ld_nowait_x16(tl + sb + 48, a1);
      chunk(masked, sb, 2, a0, nref, nvalid, rs);
      tmem_wait_ld();
      reg_fence16(a1);
      chunk(masked, sb, 3, a1, nref, nvalid, rs);
    };
    // the same pass out of registers (kOnePass): a = columns [0,32), b = columns [32,64) of this row of S(j)
    auto tile_pass_regs = [&](auto masked, uint32_t sb, const float* a, const float* b, float nref, int nvalid, float& rs) {
      chunk(masked, sb, 0, a, nref, nvalid, rs);
      chunk(masked, sb, 1, a + 16, nref, nvalid, rs);
      chunk(masked, sb, 2, b, nref, nvalid, rs);
      chunk(masked, sb, 3, b + 16, nref, nvalid, rs);
    };
    // o_reg += O accumulated in TMEM (round-to-nearest adds)
    auto drain_O = [&]() {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float a[32], b[32];
        tmem_ld_nowait_x32(tl + kColO + 32 * hf, a);
        if (!kFast) tmem_ld_nowait_x32(tl + kColOx + 32 * hf, b);
        tmem_wait_ld();
        reg_fence32(a);
        if (!kFast) reg_fence32(b);
#pragma unroll
        for (int i = 0; i < 32; ++i) o_reg[32 * hf + i] += kFast ? a[i] : fmaf(b[i], kInvLo, a[i]);
      }
    };

    for (int j = 0; j < ntiles; ++j) {
      const uint32_t sb = kColS + 64u * (uint32_t)(j & 1);        // S(j) / P(j) buffer
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const int nvalid = min(BKV, L - j * BKV);
      if (row == 0) astamp(0, j, 0);
      // pass 1: row maximum of the tile
      float mx = -INFINITY;
      float a[32], b[32];
      tmem_ld_nowait_x32(tl + sb, a);
      tmem_ld_nowait_x32(tl + sb + 32, b);
      tmem_wait_ld();
      reg_fence32(a), reg_fence32(b);
      if (nvalid < 64) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < nvalid) mx = fmaxf(mx, a[i]);
          if (32 + i < nvalid) mx = fmaxf(mx, b[i]);
        }
      } else {
        // four independent chains (a single 64-long chain of dependent FMNMX is ~300 cycles of exposed latency)
        float m0 = fmaxf(a[0], b[0]), m1 = fmaxf(a[1], b[1]), m2 = fmaxf(a[2], b[2]), m3 = fmaxf(a[3], b[3]);
#pragma unroll
        for (int i = 4; i < 32; i += 4) {
          m0 = fmaxf(m0, fmaxf(a[i], b[i]));
          m1 = fmaxf(m1, fmaxf(a[i + 1], b[i + 1]));
          m2 = fmaxf(m2, fmaxf(a[i + 2], b[i + 2]));
          m3 = fmaxf(m3, fmaxf(a[i + 3], b[i + 3]));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      if (row == 0) astamp(0, j, 1);
      const bool need = j > 0 && (mx - m_ref) * sl2 > kRebase;
      bool drained = false;
      if (j == 0) {
        m_ref = mx;
      } else if (__any_sync(0xffffffffu, need)) {
        // A key far above the reference (p would exceed 2^kRebase): move the reference to it.  Rare after the first
        // tiles - the reference only has to stay within a factor 2^kRebase of the true maximum.  Everything accumulated
        // so far is rescaled: registers, the running sum, and this row of the O accumulators in TMEM (P @ V(j-1) may
        // still be running: wait for it).  The TMEM instructions are warp-collective, so the whole warp takes the
        // path; rows that keep their reference use corr = 1.
        const float corr = need ? ex2_approx((m_ref - mx) * sl2) : 1.0f;
        if (need) m_ref = mx;
        mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        if ((j % kDrain) == 0) {       // a drain tile: what the accumulators hold is under the OLD reference - take it first
          drain_O();
          drained = true;
        }
        l_run *= corr;
#pragma unroll
        for (int d = 0; d < D; ++d) o_reg[d] *= corr;
        if ((j % kDrain) != 0) {
#pragma unroll 1
          for (int c = 0; c < (kFast ? 4 : 8); ++c) {
            float a[16];
            tmem_ld_nowait_x16(tl + kColO + 16 * c, a);        // O_main [128,192) and O_x [192,256) are contiguous
            tmem_wait_ld();
            reg_fence16(a);
            uint32_t u[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = __float_as_uint(a[i] * corr);
            tmem_st_32x32b_x8(tl + kColO + 16 * c, u);
            tmem_st_32x32b_x8(tl + kColO + 16 * c + 8, u + 8);
          }
        }
      }
      // pass 2: p against the reference, written back as the A operand of P @ V
      float rs = 0.f;
      const float nref = -m_ref * sl2;
      if (kOnePass) {
        if (nvalid < BKV)
          tile_pass_regs(std::true_type{}, sb, a, b, nref, nvalid, rs);
        else
          tile_pass_regs(std::false_type{}, sb, a, b, nref, nvalid, rs);
      } else {
        if (nvalid < BKV)
          tile_pass(std::true_type{}, sb, nref, nvalid, rs);
        else
          tile_pass(std::false_type{}, sb, nref, nvalid, rs);
      }
      l_run += rs;
      if (row == 0) astamp(0, j, 2);
      // every kDrain tiles the O accumulators move into the registers; P @ V(j) then starts fresh.  Done at the END of
      // the tile's softmax: P @ V(j-1), issued when this warpgroup finished tile j-1, has long completed by now.
      if (j > 0 && (j % kDrain) == 0 && !drained) {
        mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        drain_O();
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_ready[j & 1]);
      if (row == 0) astamp(0, j, 3);
    }
    mbar_wait(&o_full[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
    tc_fence_after();
    const int t = q0 + row;
    const float inv = 1.0f / l_run;
    const long long off = ((long long)r * l + t) * (H * D) + h * D;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float a[16], b[16];
      tmem_ld_32x32b_x16(tl + kColO + 16 * c, a);
      if (!kFast) tmem_ld_32x32b_x16(tl + kColOx + 16 * c, b);
      if (t < q_end) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float vv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            vv[e] = (o_reg[16 * c + i + e] + (kFast ? a[i + e] : fmaf(b[i + e], kInvLo, a[i + e]))) * inv;
          if (o16_hi != nullptr) st4_split_f16(o16_hi + off + 16 * c + i, o16_lo + off + 16 * c + i, vv);
          if (out != nullptr) st4(out + off + 16 * c + i, make_float4(vv[0], vv[1], vv[2], vv[3]));
        }
      }
    }
  } else if (warp == 4) {
    // ================================================================ TMA: the Q tile once, then K / V^T tiles
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, (uint32_t)((kFast ? 1 : 2) * kQTile));
      tma_load_3d(&mapQhi, q_full, q_s, 0, q0, rh);
      if (!kFast) tma_load_3d(&mapQlo, q_full, q_s + kQTile, 0, q0, rh);
      // Issue order = the order the MMA thread CONSUMES the tiles: S(0), S(1), then P @ V(j), S(j + 2), ... - K runs two tiles
      // ahead of V (round 1 issued K(j), V(j) pairwise, so K(j + 2) queued behind V(j + 1), whose stage is released a whole
      // softmax later).  Measured: K tiles now land ~1 000 cycles earlier and the kernel time does not move (1.892 ms) - the
      // ~400 cycles the MMA thread spends between P @ V(j) and S(j + 2) are not a wait for data but the tensor core working
      // through 12 dependent MMAs (profiles/r02_attn16.md section 2e).
      auto load_K = [&](int j) {
        const int s = j % kStages;
        unsigned char* st = stage(s);
        mbar_wait(&k_empty[s], ((j / kStages) & 1) ^ 1);   // released by S(j - kStages): early
        astamp(2, j, 0);
        mbar_arrive_expect_tx(&k_full[s], (uint32_t)((kFast ? 1 : 2) * kTile));
        tma_load_3d(&mapKhi, &k_full[s], st + 0 * kTile, 0, j * BKV, rh);
        if (!kFast) tma_load_3d(&mapKlo, &k_full[s], st + 1 * kTile, 0, j * BKV, rh);
      };
      auto load_V = [&](int j) {
        const int s = j % kStages;
        unsigned char* st = stage(s);
        mbar_wait(&v_empty[s], ((j / kStages) & 1) ^ 1);   // released by P @ V(j - kStages): a softmax later
        astamp(2, j, 1);
        mbar_arrive_expect_tx(&v_full[s], (uint32_t)((kFast ? 1 : 2) * kTile));
        tma_load_3d(&mapVhi, &v_full[s], st + 2 * kTile, j * BKV, 0, rh);
        if (!kFast) tma_load_3d(&mapVlo, &v_full[s], st + 3 * kTile, j * BKV, 0, rh);
      };
      static_assert(kStages == 2, "the K-ahead issue order below assumes S runs kStages tiles ahead of P @ V");
      load_K(0);
      if (ntiles > 1) load_K(1);
      for (int j = 0; j < ntiles; ++j) {
        load_V(j);
        if (j + 2 < ntiles) load_K(j + 2);
      }
    }
  } else {
    // ================================================================ MMA issue
    if (elect_one()) {
      const uint64_t dqh = G::desc(smem_u32(q_s)), dql = G::desc(smem_u32(q_s + kQTile));
      // S(j) = Q K(j)^T into buffer j & 1: ONE accumulator.  The cross terms (~2^-11 of the result) go first, while the
      // accumulator is small, so that the tensor core's truncation after every MMA acts on the big sum only during the
      // last four steps (the hi * hi products), as with a separate cross accumulator.
      auto issue_S = [&](int j) {
        const int s = j % kStages;
        mbar_wait(&k_full[s], (j / kStages) & 1);
        tc_fence_after();
        astamp(1, j - 2, 2);
        unsigned char* st = stage(s);
        const uint32_t d = tmem_base + kColS + 64u * (uint32_t)(j & 1);
        const uint64_t dkh = G::desc(smem_u32(st)), dkl = G::desc(smem_u32(st + kTile));
        if (!kFast) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {                     // 16 head dims (32 bytes of a row) per MMA
            const uint64_t adv = (uint64_t)(2 * k);
            umma_f16_ss(d, dql + adv, dkh + adv, kIdesc, k != 0);
            umma_f16_ss(d, dqh + adv, dkl + adv, kIdesc, 1u);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adv = (uint64_t)(2 * k);
          umma_f16_ss(d, dqh + adv, dkh + adv, kIdesc, (kFast && k == 0) ? 0u : 1u);
        }
        umma_commit(&s_full[j & 1]);
        umma_commit(&k_empty[s]);
        astamp(1, j - 2, 3);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_S(0);
      if (ntiles > 1) issue_S(1);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % kStages;
        mbar_wait(&v_full[s], (j / kStages) & 1);
        mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
        tc_fence_after();
        astamp(1, j, 0);
        unsigned char* st = stage(s);
        const uint64_t dvh = G::desc(smem_u32(st + 2 * kTile)), dvl = G::desc(smem_u32(st + 3 * kTile));
        const uint32_t pb = tmem_base + kColS + 64u * (uint32_t)(j & 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {                       // 16 keys per MMA: P cells [16k, 16k+8) hi, [16k+8, 16k+16) lo
          const uint64_t adv = (uint64_t)(2 * k);
          const uint32_t p_hi = pb + 16 * k, p_lo = p_hi + 8;
          const uint32_t acc = (k != 0 || (j % kDrain) != 0) ? 1u : 0u;   // fresh accumulators every kDrain tiles
          if (!kFast) {
            umma_f16_ts(tmem_base + kColOx, p_lo, dvh + adv, kIdesc, acc);
            umma_f16_ts(tmem_base + kColOx, p_hi, dvl + adv, kIdesc, 1u);
          }
          umma_f16_ts(tmem_base + kColO, p_hi, dvh + adv, kIdesc, acc);
        }
        umma_commit(&o_full[j & 1]);
        umma_commit(&v_empty[s]);
        astamp(1, j, 1);
        // in-order execution of tcgen05.mma: S(j+2) overwrites the P(j) cells only after P @ V(j) has read them
        if (j + 2 < ntiles) issue_S(j + 2);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, kTmemCols);
}

// CVAR_ATTN_ONE_PASS = 1 selects the one-read softmax variant (A/B timing).  Default 0: measured SLOWER (2.31 vs 2.06 ms at
// the last scale, profiles/r02_attn16.md) - with the running output row (64 registers) the 64 S registers spill.
static int initial_one_pass() {
  const char* e = getenv("CVAR_ATTN_ONE_PASS");
  return (e != nullptr && e[0] == '1') ? 1 : 0;
}
int g_one_pass = initial_one_pass();
int set_trace(long long* p) { return cudaMemcpyToSymbol(g_attn16_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -1; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// 3-D map over a pair array of halves: dims (inner, mid, RH), box (64, box_mid, 1), 128-byte swizzle
static int make_map3(CUtensorMap* map, const void* base, long long inner, long long mid, long long rh, int box_mid) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cvar_attn_kvcache16: cuTensorMapEncodeTiled is not available from the driver");
    return -3;
  }
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)mid, (cuuint64_t)rh};
  cuuint64_t strides[2] = {(cuuint64_t)inner * 2, (cuuint64_t)inner * mid * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_mid, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cvar_attn_kvcache16: cuTensorMapEncodeTiled failed with %d", (int)rc);
    return -3;
  }
  return 0;
}
}  // namespace tcattn16

static int launch_attn16_tc(const __half* qh, const __half* ql, const __half* kh, const __half* kl, const __half* vh,
                            const __half* vl, float* out, __half* o16h, __half* o16l, int R, int H, int l, int L, int T_max,
                            float scale, const tcattn16::AttnSegs& segs, cudaStream_t stream, const char* name) {
  CUtensorMap mqh, mql, mkh, mkl, mvh, mvl;
  const long long RH = (long long)R * H;
  int rc = tcattn16::make_map3(&mqh, qh, 64, l, RH, tcattn16::BQ);
  if (!rc) rc = tcattn16::make_map3(&mql, ql, 64, l, RH, tcattn16::BQ);
  if (!rc) rc = tcattn16::make_map3(&mkh, kh, 64, T_max, RH, tcattn16::BKV);
  if (!rc) rc = tcattn16::make_map3(&mkl, kl, 64, T_max, RH, tcattn16::BKV);
  if (!rc) rc = tcattn16::make_map3(&mvh, vh, T_max, 64, RH, 64);
  if (!rc) rc = tcattn16::make_map3(&mvl, vl, T_max, 64, RH, 64);
  if (rc) return rc;
  auto kern = g_fast_mode ? (tcattn16::g_one_pass ? tcattn16::attn16_tc_kernel<true, true> : tcattn16::attn16_tc_kernel<true, false>)
                          : (tcattn16::g_one_pass ? tcattn16::attn16_tc_kernel<false, true> : tcattn16::attn16_tc_kernel<false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tcattn16::kSmem);
  CVAR_REQUIRE(e == cudaSuccess, "%s: cannot raise shared memory: %s", name, cudaGetErrorString(e));
  dim3 grid(segs.n > 0 ? segs.n : cdiv(l, tcattn16::BQ), H, R);
  kern<<<grid, tcattn16::kThreads, tcattn16::kSmem, stream>>>(mqh, mql, mkh, mkl, mvh, mvl, out, o16h, o16l, H, l, L, scale,
                                                               segs);
  CVAR_CHECK_LAUNCH(name);
  return 0;
}

extern "C" int cvar_attn_kvcache16(const void* q16_hi, const void* q16_lo, const void* k16_hi, const void* k16_lo,
                                   const void* vt16_hi, const void* vt16_lo, float* out, void* out16_hi, void* out16_lo,
                                   int R, int H, int l, int L, int T_max, float scale, int engine, void* stream) {
  CVAR_REQUIRE(q16_hi && q16_lo && k16_hi && k16_lo && vt16_hi && vt16_lo, "cvar_attn_kvcache16: null operand");
  CVAR_REQUIRE(out != nullptr || out16_hi != nullptr, "cvar_attn_kvcache16: no output");
  CVAR_REQUIRE((out16_hi == nullptr) == (out16_lo == nullptr), "cvar_attn_kvcache16: out16_hi/out16_lo must come together");
  CVAR_REQUIRE(R > 0 && H > 0 && l > 0 && L >= l && L <= T_max, "cvar_attn_kvcache16: bad shape l=%d L=%d T=%d", l, L,
               T_max);
  CVAR_REQUIRE(R <= 65535 && H <= 65535, "cvar_attn_kvcache16: grid too large");
  CVAR_REQUIRE(T_max % 8 == 0, "cvar_attn_kvcache16: T_max must be a multiple of 8 (got %d)", T_max);
  const __half* qh = reinterpret_cast<const __half*>(q16_hi);
  const __half* ql = reinterpret_cast<const __half*>(q16_lo);
  const __half* kh = reinterpret_cast<const __half*>(k16_hi);
  const __half* kl = reinterpret_cast<const __half*>(k16_lo);
  const __half* vh = reinterpret_cast<const __half*>(vt16_hi);
  const __half* vl = reinterpret_cast<const __half*>(vt16_lo);
  __half* o16h = reinterpret_cast<__half*>(out16_hi);
  __half* o16l = reinterpret_cast<__half*>(out16_lo);
  // default: tensor cores for every scale.  A 128-query tile that is mostly empty still beats the SIMT kernel (l = 32:
  // 0.076 vs 0.17 ms per layer, l = 18 / 8 / 2: ~0.07 vs ~0.13 ms, profiles/r02_attn16.md), and the block-causal pass
  // runs the short scales on the same kernel anyway.
  if (engine < 0) engine = (g_gemm_engine != 0) ? 1 : 0;
  if (engine == 1) {
    tcattn16::AttnSegs segs;
    segs.n = 0;
    return launch_attn16_tc(qh, ql, kh, kl, vh, vl, out, o16h, o16l, R, H, l, L, T_max, scale, segs, (cudaStream_t)stream,
                            "cvar_attn_kvcache16[tc]");
  }
  cudaError_t e = cudaFuncSetAttribute(attn16_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(AttnSmem));
  CVAR_REQUIRE(e == cudaSuccess, "cvar_attn_kvcache16: cannot raise shared memory: %s", cudaGetErrorString(e));
  dim3 grid(cdiv(l, BQ), H, R);
  attn16_simt_kernel<<<grid, 256, sizeof(AttnSmem), (cudaStream_t)stream>>>(qh, ql, kh, kl, vh, vl, out, o16h, o16l, H, l,
                                                                           L, T_max, scale);
  CVAR_CHECK_LAUNCH("cvar_attn_kvcache16");
  return 0;
}

// Block-causal attention over a whole token pyramid in ONE launch (ControlVAR.forward): scale s holds host_scale_lens[s]
// consecutive queries and sees the keys of scales 0..s.  q16 (R, H, l_total, 64) and the caches as cvar_qkv_project16 wrote
// them with L_prev = 0, l = l_total.
extern "C" int cvar_attn_blockcausal16(const void* q16_hi, const void* q16_lo, const void* k16_hi, const void* k16_lo,
                                       const void* vt16_hi, const void* vt16_lo, float* out, void* out16_hi, void* out16_lo,
                                       int R, int H, int l_total, int T_max, float scale, int n_scales,
                                       const int* host_scale_lens, void* stream) {
  CVAR_REQUIRE(q16_hi && q16_lo && k16_hi && k16_lo && vt16_hi && vt16_lo, "cvar_attn_blockcausal16: null operand");
  CVAR_REQUIRE(out != nullptr || out16_hi != nullptr, "cvar_attn_blockcausal16: no output");
  CVAR_REQUIRE((out16_hi == nullptr) == (out16_lo == nullptr), "cvar_attn_blockcausal16: out16_hi/out16_lo must come together");
  CVAR_REQUIRE(g_gemm_engine != 0, "cvar_attn_blockcausal16: needs a tensor-core engine (engine is 0 = SIMT)");
  CVAR_REQUIRE(R > 0 && H > 0 && R <= 65535 && H <= 65535 && n_scales > 0 && host_scale_lens != nullptr && l_total <= T_max &&
                   T_max % 8 == 0,
               "cvar_attn_blockcausal16: bad shape R=%d H=%d l=%d T=%d scales=%d", R, H, l_total, T_max, n_scales);
  tcattn16::AttnSegs segs;
  segs.n = 0;
  int start = 0;
  for (int s = 0; s < n_scales; ++s) {
    const int ls = host_scale_lens[s];
    CVAR_REQUIRE(ls > 0, "cvar_attn_blockcausal16: scale %d has no tokens", s);
    for (int q = 0; q < ls; q += tcattn16::BQ) {
      CVAR_REQUIRE(segs.n < tcattn16::kMaxSegTiles, "cvar_attn_blockcausal16: more than %d query tiles", tcattn16::kMaxSegTiles);
      segs.q0[segs.n] = start + q;
      segs.nq[segs.n] = ls - q < tcattn16::BQ ? ls - q : tcattn16::BQ;
      segs.L[segs.n] = start + ls;
      ++segs.n;
    }
    start += ls;
  }
  CVAR_REQUIRE(start == l_total, "cvar_attn_blockcausal16: scale lengths sum to %d, not l_total = %d", start, l_total);
  return launch_attn16_tc(reinterpret_cast<const __half*>(q16_hi), reinterpret_cast<const __half*>(q16_lo),
                          reinterpret_cast<const __half*>(k16_hi), reinterpret_cast<const __half*>(k16_lo),
                          reinterpret_cast<const __half*>(vt16_hi), reinterpret_cast<const __half*>(vt16_lo), out,
                          reinterpret_cast<__half*>(out16_hi), reinterpret_cast<__half*>(out16_lo), R, H, l_total, l_total, T_max,
                          scale, segs, (cudaStream_t)stream, "cvar_attn_blockcausal16");
}
