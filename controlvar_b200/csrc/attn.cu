// KV-cached attention for the next-scale sampler (basic_var.py:106-117): l new queries of one scale attend to the
// L = L_prev + l cached keys of all scales so far; no mask (the cache IS the block-causal prefix).
// fp32 flash-style kernel: one CTA = 64 queries of one (row, head); K/V streamed through shared memory in 64-key
// tiles, online softmax, accumulators in registers.
#include "common.cuh"
#include "tc_ptx.cuh"

using namespace cvar;

namespace {
constexpr int BQ = 64, BKV = 64, D = 64;
constexpr int PS = 68;   // row stride of the transposed P tile (floats), keeps float4 alignment

struct AttnSmem {
  float Qt[D][BQ];      // Qt[d][i]  = scale * q[i][d]
  float Kt[D][BKV];     // Kt[d][j]  = k[j][d]
  float V[BKV][D];      // V[j][d]
  float Pt[BKV][PS];    // Pt[j][i]  = exp(s[i][j] - m[i])
};

__global__ void __launch_bounds__(256) attn_kvcache_kernel(const float* __restrict__ q, const float* __restrict__ k_hi,
                                                           const float* __restrict__ k_lo,
                                                           const float* __restrict__ vt_hi,
                                                           const float* __restrict__ vt_lo, float* __restrict__ out,
                                                           float* __restrict__ out_lo, __half* __restrict__ o16_hi,
                                                           __half* __restrict__ o16_lo, int H, int l, int L, int T_max,
                                                           float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, r = blockIdx.z;
  const float* qb = q + (((long long)r * H + h) * l) * D;
  // cache format of cvar_qkv_project: K split hi/lo [rh][T][64] (hi + lo == k exactly), V^T split [rh][64][T]
  const long long kv_off = (((long long)r * H + h) * T_max) * D;
  const float* kbh = k_hi + kv_off;
  const float* kbl = k_lo + kv_off;
  const float* vbh = vt_hi + kv_off;
  const float* vbl = vt_lo + kv_off;

  // Q tile, transposed into shared memory (lanes walk the query index so the transposing store is conflict-free)
  for (int it = 0; it < 4; ++it) {
    int item = it * 256 + tid;
    int i = item & 63, dq = item >> 6;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + i < l) v = ld4(qb + (long long)(q0 + i) * D + dq * 4);
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
    __syncthreads();   // previous tile fully consumed (also orders the Q stores before the first S)
    for (int it = 0; it < 4; ++it) {
      int item = it * 256 + tid;
      int j = item & 63, dq = item >> 6;          // lanes walk the key index: conflict-free transposing store
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + j < L) {
        float4 a = ld4(kbh + (long long)(k0 + j) * D + dq * 4), b = ld4(kbl + (long long)(k0 + j) * D + dq * 4);
        kv = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      }
      sm.Kt[dq * 4 + 0][j] = kv.x;
      sm.Kt[dq * 4 + 1][j] = kv.y;
      sm.Kt[dq * 4 + 2][j] = kv.z;
      sm.Kt[dq * 4 + 3][j] = kv.w;
    }
    for (int it = 0; it < 4; ++it) {
      int item = it * 256 + tid;
      int d = item >> 4, jq = item & 15;          // V^T rows are keys-contiguous: 16 lanes read 64 consecutive keys
      const long long base = (long long)d * T_max + k0 + jq * 4;
      float vv[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + jq * 4 + 3 < T_max) {
        float4 a = ld4(vbh + base), b = ld4(vbl + base);
        vv[0] = a.x + b.x, vv[1] = a.y + b.y, vv[2] = a.z + b.z, vv[3] = a.w + b.w;
      } else {
        for (int i = 0; i < 4; ++i)
          if (k0 + jq * 4 + i < T_max) vv[i] = vbh[base + i] + vbl[base + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int j = jq * 4 + i;
        // 16-byte chunk index XOR-swizzled with the key so the transposing store spreads over banks
        sm.V[j][(((d >> 2) ^ (j & 15)) << 2) + (d & 3)] = (k0 + j < L) ? vv[i] : 0.f;
      }
    }
    __syncthreads();

    // S[i][j] for i = ty*4+ii, j = jj*16+tx
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
    // mask the tail of the last tile, online softmax per query row (a row is spread over the 16 tx lanes)
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
      float corr = expf(mrow[i] - mnew);   // exp(-inf) = 0 on the first tile
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
    // O[i][d] += sum_j P[i][j] V[j][d], d = tx*4 + dd
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
      float4 v = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
      const long long off = ((long long)r * l + t) * C + h * D + tx * 4;
      if (o16_hi != nullptr) {     // FP16 pair for the f16x3 proj GEMM
        const float vv[4] = {v.x, v.y, v.z, v.w};
        st4_split_f16(o16_hi + off, o16_lo + off, vv);
        if (out == nullptr) continue;
      }
      if (out_lo != nullptr) {     // TF32 split for the all-TMA proj GEMM
        float4 hi = make_float4(tc::trunc_tf32(v.x), tc::trunc_tf32(v.y), tc::trunc_tf32(v.z), tc::trunc_tf32(v.w));
        st4(out_lo + off, make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w));
        v = hi;
      }
      st4(out + off, v);
    }
  }
}
}  // namespace


// =====================================================================================================================
// Tensor-core path (tcgen05 + TMEM + TMA), fp32-class accuracy through the same 3xTF32 split as the GEMM engine.
//   CTA = 128 queries of one (row, head); KV streamed in 64-key tiles through a 2-stage TMA ring.  K tiles and V^T tiles
//   arrive already split hi/lo (cvar_qkv_project wrote them that way), so the kernel does no operand conversion for K/V.
//   warps 0-3  softmax: thread i owns query row i.  S = Q K^T is read from TMEM (tcgen05.ld), exp'ed, split hi/lo and
//              written back to TMEM as the A operand of P @ V (tcgen05.st).  The running output row lives in REGISTERS:
//              every 64-key tile produces a fresh O_tile in TMEM that is added with round-to-nearest fp32 adds, so the
//              tensor core's round-toward-zero accumulation never runs longer than 8 steps.
//   warp 4     TMA issue (8 boxes per tile: K_hi, K_lo, V^T_hi, V^T_lo, two 128-byte-wide blocks each)
//   warp 5     TMEM allocation + MMA issue; S(j+1) is issued before P@V(j) so QK^T of the next tile overlaps softmax(j)
//   TMEM columns: S_main [0,64) S_lo [64,128) P_hi [128,192) P_lo [192,256) O_tile [256,320) Q_hi [320,384) Q_lo [384,448)
namespace tcattn {
using namespace cvar::tc;
// Optional phase trace (diagnostics): CTA (0,0,0) stamps clock64() per KV tile.  trace[(who * 32 + j) * 8 + ev], j < 32.
//   who 0 = softmax thread 0: ev 0 s_full seen, 1 S loaded, 2 p computed, 3 o_full(j-1) seen, 4 O updated, 5 P stored+signalled
//   who 1 = MMA thread:       ev 0 S(j+1) issued, 1 p_ready(j) seen, 2 PV(j) issued
__device__ long long* g_attn_trace = nullptr;
__device__ __forceinline__ void astamp(int who, int j, int ev) {
  if (g_attn_trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j < 32)
    g_attn_trace[(who * 32 + j) * 8 + ev] = clock64();
}
constexpr int BQ = 128, BKV = 64, D = 64;
constexpr int kThreads = 192;
constexpr int kKVBlock = BKV * 128;       // bytes of one 32-wide K-block of a K / V^T tile (64 rows x 128 B)
constexpr int kStageBytes = 4 * 2 * kKVBlock;            // K_hi K_lo VT_hi VT_lo, two blocks = 64 KiB
// Q lives in TMEM (the A operand of S = Q K^T is read from there, like P for P @ V), so shared memory holds nothing but
// the K/V ring and THREE stages fit.  The phase trace of the 2-stage version (profiles/r01_attn_trace.md) showed the
// ~2700-cycle TMA round trip of a tile on the critical path of every iteration: the softmax threads waited for S half
// of the time.
constexpr int kStages = 3;
constexpr int kSmem = kStages * kStageBytes + 1024 + 1024;
constexpr uint32_t kColSmain = 0, kColSlo = 64, kColPhi = 128, kColPlo = 192, kColO = 256, kColQhi = 320, kColQlo = 384;
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap mapKhi, const __grid_constant__ CUtensorMap mapKlo,
               const __grid_constant__ CUtensorMap mapVhi, const __grid_constant__ CUtensorMap mapVlo,
               const float* __restrict__ q, float* __restrict__ out, float* __restrict__ out_lo,
               __half* __restrict__ o16_hi, __half* __restrict__ o16_lo, int H, int l, int L,
               float scale) {
  using G = Geo<32>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  auto stage = [&](int s) { return smem + s * kStageBytes; };   // K_hi | K_lo | VT_hi | VT_lo (2 blocks each)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* kv_full = bars;                  // [kStages]
  uint64_t* kv_empty = bars + kStages;       // [kStages]
  uint64_t* s_full = bars + 2 * kStages;     // S(j) accumulated
  uint64_t* s_free = bars + 2 * kStages + 1; // S(j) read by all 128 softmax threads
  uint64_t* p_ready = bars + 2 * kStages + 2;// P(j) written to TMEM by all 128 softmax threads
  uint64_t* o_full = bars + 2 * kStages + 3; // O_tile(j) accumulated
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, r = blockIdx.z;
  const long long rh = (long long)r * H + h;
  const int ntiles = (L + BKV - 1) / BKV;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&mapKhi), tma_prefetch_desc(&mapKlo), tma_prefetch_desc(&mapVhi), tma_prefetch_desc(&mapVlo);
    for (int s = 0; s < kStages; ++s) mbar_init(&kv_full[s], 1), mbar_init(&kv_empty[s], 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {
    // Q row of this thread: load, pre-scale, split hi/lo, park in TMEM (lane = query row, column = d) as the A operand
    const int qrow = q0 + (int)threadIdx.x;
    const float* qp = q + (rh * l + qrow) * D;
    const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float hi[32], lo[32];
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qrow < l) x = ld4(qp + half * 32 + c4 * 4);
        const float xs[4] = {x.x * scale, x.y * scale, x.z * scale, x.w * scale};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          hi[c4 * 4 + e] = trunc_tf32(xs[e]);
          lo[c4 * 4 + e] = xs[e] - hi[c4 * 4 + e];
        }
      }
      tmem_st_32x32b_x32(tl + kColQhi + half * 32, hi);
      tmem_st_32x32b_x32(tl + kColQlo + half * 32, lo);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();      // Q is in TMEM before the first S = Q K^T is issued
  tc_fence_after();

  if (warp < 4) {
    // ================================================================ softmax + output rows
    const int row = threadIdx.x;
    const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
    float o_reg[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o_reg[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(s_full, j & 1);
      if (row == 0) astamp(0, j, 0);
      tc_fence_after();
      float s[BKV];
      {
        float a[32], b[32];
        tmem_ld_32x32b_x32(tlane + kColSmain, a);
        tmem_ld_32x32b_x32(tlane + kColSlo, b);
#pragma unroll
        for (int c = 0; c < 32; ++c) s[c] = a[c] + b[c];
        tmem_ld_32x32b_x32(tlane + kColSmain + 32, a);
        tmem_ld_32x32b_x32(tlane + kColSlo + 32, b);
#pragma unroll
        for (int c = 0; c < 32; ++c) s[32 + c] = a[c] + b[c];
      }
      tc_fence_before();
      mbar_arrive(s_free);
      if (row == 0) astamp(0, j, 1);
      const int kbase = j * BKV;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < BKV; ++c) {
        if (kbase + c >= L) s[c] = -INFINITY;
        mx = fmaxf(mx, s[c]);
      }
      const float m_new = fmaxf(m_run, mx);
      // expf, not exp2f(x * log2e): the rounding of that product is an ABSOLUTE exponent error of |x| * 2^-24, i.e. a
      // relative error in p that grows with the logit range (cosine attention multiplies q by up to 100)
      const float corr = expf(m_run - m_new);                 // 0 on the first tile
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < BKV; ++c) {
        s[c] = expf(s[c] - m_new);
        rs += s[c];
      }
      l_run = l_run * corr + rs;
      m_run = m_new;
      if (row == 0) astamp(0, j, 2);
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        if (row == 0) astamp(0, j, 3);
        tc_fence_after();
        float a[32];
        tmem_ld_32x32b_x32(tlane + kColO, a);
#pragma unroll
        for (int c = 0; c < 32; ++c) o_reg[c] = (o_reg[c] + a[c]) * corr;
        tmem_ld_32x32b_x32(tlane + kColO + 32, a);
#pragma unroll
        for (int c = 0; c < 32; ++c) o_reg[32 + c] = (o_reg[32 + c] + a[c]) * corr;
      }
      if (row == 0) astamp(0, j, 4);
      // P(j) -> TMEM as the A operand of P @ V, split hi/lo (P(j-1) was consumed: o_full(j-1) has been observed)
      {
        float hi[32], lo[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            hi[c] = trunc_tf32(s[half * 32 + c]);
            lo[c] = s[half * 32 + c] - hi[c];
          }
          tmem_st_32x32b_x32(tlane + kColPhi + half * 32, hi);
          tmem_st_32x32b_x32(tlane + kColPlo + half * 32, lo);
        }
        tmem_wait_st();
      }
      tc_fence_before();
      mbar_arrive(p_ready);
      if (row == 0) astamp(0, j, 5);
    }
    mbar_wait(o_full, (ntiles - 1) & 1);
    tc_fence_after();
    {
      float a[32];
      tmem_ld_32x32b_x32(tlane + kColO, a);
#pragma unroll
      for (int c = 0; c < 32; ++c) o_reg[c] += a[c];
      tmem_ld_32x32b_x32(tlane + kColO + 32, a);
#pragma unroll
      for (int c = 0; c < 32; ++c) o_reg[32 + c] += a[c];
    }
    const int t = q0 + row;
    if (t < l) {
      const float inv = 1.0f / l_run;
      const long long off = ((long long)r * l + t) * (H * D) + h * D;
#pragma unroll
      for (int d = 0; d < D; d += 4) {
        float4 v = make_float4(o_reg[d] * inv, o_reg[d + 1] * inv, o_reg[d + 2] * inv, o_reg[d + 3] * inv);
        if (o16_hi != nullptr) {   // FP16 pair for the f16x3 proj GEMM
          const float vv[4] = {v.x, v.y, v.z, v.w};
          st4_split_f16(o16_hi + off + d, o16_lo + off + d, vv);
          if (out == nullptr) continue;
        }
        if (out_lo != nullptr) {   // TF32 split for the all-TMA proj GEMM
          float4 hi = make_float4(trunc_tf32(v.x), trunc_tf32(v.y), trunc_tf32(v.z), trunc_tf32(v.w));
          st4(out_lo + off + d, make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w));
          v = hi;
        }
        st4(out + off + d, v);
      }
    }
  } else if (warp == 4) {
    // ================================================================ TMA: K / V^T tiles, pre-split
    if (lane == 0) {
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % kStages;
        mbar_wait(&kv_empty[s], ((j / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], (uint32_t)kStageBytes);
        unsigned char* st = stage(s);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          tma_load_3d(&mapKhi, &kv_full[s], st + (0 + b) * kKVBlock, b * 32, j * BKV, (int)rh);
          tma_load_3d(&mapKlo, &kv_full[s], st + (2 + b) * kKVBlock, b * 32, j * BKV, (int)rh);
          tma_load_3d(&mapVhi, &kv_full[s], st + (4 + b) * kKVBlock, j * BKV + b * 32, 0, (int)rh);
          tma_load_3d(&mapVlo, &kv_full[s], st + (6 + b) * kKVBlock, j * BKV + b * 32, 0, (int)rh);
        }
      }
    }
  } else {
    // ================================================================ MMA issue
    if (lane == 0) {
      auto issue_S = [&](int j) {
        unsigned char* st = stage(j % kStages);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t dkh = G::desc(smem_u32(st + (0 + kb) * kKVBlock)), dkl = G::desc(smem_u32(st + (2 + kb) * kKVBlock));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(2 * k);
            const uint32_t qc = (uint32_t)((kb * 4 + k) * 8);       // 8 d-columns of Q per k-step
            umma_tf32_ts(tmem_base + kColSlo, tmem_base + kColQlo + qc, dkh + adv, kIdesc, (kb | k) != 0);
            umma_tf32_ts(tmem_base + kColSlo, tmem_base + kColQhi + qc, dkl + adv, kIdesc, 1u);
            umma_tf32_ts(tmem_base + kColSmain, tmem_base + kColQhi + qc, dkh + adv, kIdesc, (kb | k) != 0);
          }
        }
        umma_commit(s_full);
      };
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_S(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) {
          mbar_wait(&kv_full[(j + 1) % kStages], ((j + 1) / kStages) & 1);
          mbar_wait(s_free, j & 1);
          tc_fence_after();
          issue_S(j + 1);
          astamp(1, j, 0);
        }
        mbar_wait(p_ready, j & 1);
        astamp(1, j, 1);
        tc_fence_after();
        unsigned char* st = stage(j % kStages);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t adv = (uint64_t)(2 * (k & 3));
          const uint64_t dvh = G::desc(smem_u32(st + (4 + (k >> 2)) * kKVBlock)) + adv;
          const uint64_t dvl = G::desc(smem_u32(st + (6 + (k >> 2)) * kKVBlock)) + adv;
          umma_tf32_ts(tmem_base + kColO, tmem_base + kColPlo + 8 * k, dvh, kIdesc, k != 0);
          umma_tf32_ts(tmem_base + kColO, tmem_base + kColPhi + 8 * k, dvl, kIdesc, 1u);
          umma_tf32_ts(tmem_base + kColO, tmem_base + kColPhi + 8 * k, dvh, kIdesc, 1u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[j % kStages]);
        astamp(1, j, 2);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

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
// 3-D map over a cache array: dims (inner, mid, RH) with a (32, 64, 1) box and 128-byte swizzle
static int make_map3(CUtensorMap* map, const float* base, long long inner, long long mid, long long rh) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cvar_attn_kvcache: cuTensorMapEncodeTiled is not available from the driver");
    return -3;
  }
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)mid, (cuuint64_t)rh};
  cuuint64_t strides[2] = {(cuuint64_t)inner * 4, (cuuint64_t)inner * mid * 4};
  cuuint32_t box[3] = {32, 64, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cvar_attn_kvcache: cuTensorMapEncodeTiled failed with %d", (int)rc);
    return -3;
  }
  return 0;
}
int set_trace(long long* p) { return cudaMemcpyToSymbol(g_attn_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -1; }
}  // namespace tcattn

namespace tcattn16 { int set_trace(long long* p); }
extern "C" int cvar_debug_set_attn_trace(long long* dev_buf) {
  if (tcattn16::set_trace(dev_buf) != 0) return -1;      // the FP16-pair kernel shares the switch (3*32*8 int64, own layout)
  return tcattn::set_trace(dev_buf);
}

extern "C" int cvar_attn_kvcache(const float* q, const float* k_hi, const float* k_lo, const float* vt_hi,
                                 const float* vt_lo, float* out, float* out_lo, void* out16_hi, void* out16_lo, int R,
                                 int H, int l, int L, int T_max, float scale, int engine, void* stream) {
  CVAR_REQUIRE(out != nullptr || out16_hi != nullptr, "cvar_attn_kvcache: no output");
  CVAR_REQUIRE((out16_hi == nullptr) == (out16_lo == nullptr), "cvar_attn_kvcache: out16_hi/out16_lo must come together");
  CVAR_REQUIRE(out != nullptr || out_lo == nullptr, "cvar_attn_kvcache: out_lo without out");
  __half* o16h = reinterpret_cast<__half*>(out16_hi);
  __half* o16l = reinterpret_cast<__half*>(out16_lo);
  CVAR_REQUIRE(R > 0 && H > 0 && l > 0 && L >= l && L <= T_max, "cvar_attn_kvcache: bad shape l=%d L=%d T=%d", l, L,
               T_max);
  CVAR_REQUIRE(R <= 65535 && H <= 65535, "cvar_attn_kvcache: grid too large");
  CVAR_REQUIRE(T_max % 4 == 0, "cvar_attn_kvcache: T_max must be a multiple of 4 (got %d)", T_max);
  if (engine < 0) engine = (g_gemm_engine != 0 && l >= 64) ? 1 : 0;
  if (engine == 1) {
    CUtensorMap mkh, mkl, mvh, mvl;
    const long long RH = (long long)R * H;
    int rc = tcattn::make_map3(&mkh, k_hi, 64, T_max, RH);
    if (!rc) rc = tcattn::make_map3(&mkl, k_lo, 64, T_max, RH);
    if (!rc) rc = tcattn::make_map3(&mvh, vt_hi, T_max, 64, RH);
    if (!rc) rc = tcattn::make_map3(&mvl, vt_lo, T_max, 64, RH);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(tcattn::attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         tcattn::kSmem);
    CVAR_REQUIRE(e == cudaSuccess, "cvar_attn_kvcache: cannot raise shared memory: %s", cudaGetErrorString(e));
    dim3 grid(cdiv(l, tcattn::BQ), H, R);
    tcattn::attn_tc_kernel<<<grid, tcattn::kThreads, tcattn::kSmem, (cudaStream_t)stream>>>(mkh, mkl, mvh, mvl, q, out,
                                                                                          out_lo, o16h, o16l, H, l, L, scale);
    CVAR_CHECK_LAUNCH("cvar_attn_kvcache[tc]");
    return 0;
  }
  cudaError_t e = cudaFuncSetAttribute(attn_kvcache_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(AttnSmem));
  CVAR_REQUIRE(e == cudaSuccess, "cvar_attn_kvcache: cannot raise shared memory: %s", cudaGetErrorString(e));
  dim3 grid(cdiv(l, BQ), H, R);
  attn_kvcache_kernel<<<grid, 256, sizeof(AttnSmem), (cudaStream_t)stream>>>(q, k_hi, k_lo, vt_hi, vt_lo, out, out_lo, o16h,
                                                                            o16l, H, l, L, T_max, scale);
  CVAR_CHECK_LAUNCH("cvar_attn_kvcache");
  return 0;
}
