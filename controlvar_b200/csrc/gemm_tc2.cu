// 2-CTA (cta_group::2) tcgen05 GEMM for the dense transformer layers: out = epilogue(A[M,K] * W[N,K]^T), three
// error-compensated MMAs per product in one of two operand formats:
//   TF32 pairs (F16 = false): x = hi + lo, fp32 storage, kind::tf32                       (engine 3)
//   FP16 pairs (F16 = true) : x = h + l * 2^-11, h = half_rn(x), l = half_rn((x - h) * 2^11), kind::f16   (engine 4)
//     - the residual is SCALED so that it stays in fp16's normal range; it accumulates in the 'lo' TMEM accumulator,
//       which the epilogue folds in with one fma(lo, 2^-11, main).  Representation error 2^-24 |x| (round-to-nearest
//       twice) against 2^-21 for the truncating TF32 split, at TWICE the tensor-core rate and half the operand bytes.
//       Operands must be below 65504 in magnitude: LayerNorm-modulated rows, GELU outputs, attention outputs and
//       weights are (cvar_split_f16 saturates anything else).
// The shared-memory row is 128 bytes in both formats (32 floats or 64 halves), so tiles, swizzle and UMMA descriptors
// are identical; only the instruction kind, the K extent of a block and the tensor-map element type differ.
//
// Why a second kernel: tools/trace_gemm.py showed the 1-CTA engine (gemm_tc.cu) is bound by the TMA round trip
// (~2080 cycles for the 64 KB of weight tiles of a K-block) against 1460 cycles of MMA, with room for only TWO 96 KB
// stages.  A CTA pair computes a 256 x 256 tile: each CTA keeps its own 128 rows of A and only HALF of the weight rows,
// so a stage is 64 KB, THREE stages fit, and the MMA time per K-block is unchanged (M = 256 across the pair).
//
// All four operand tiles arrive by TMA: activations are split hi/lo by the kernel that PRODUCES them
// (cvar_ln_modulate, the GELU epilogue, cvar_attn_kvcache) exactly as the weights are split once at pack time.  There
// is therefore no generic-proxy writer of operand tiles in this kernel: every shared-memory fill is an async-proxy
// write that completes on an mbarrier, which is what makes the cross-CTA hand-over simple (and standard).
//
// Per CTA: warps 0-7 epilogue, warp 8 TMA, warp 9 TMEM alloc (+ MMA issue in the leader CTA, rank 0).
// Barriers: full[s]  - LEADER's copy only; tx bytes from both CTAs (cta_group::2 TMA signals the leader's barrier)
//           empty[s] - both copies; released by the leader's tcgen05.commit ... multicast::cluster 0b11
//           tm_full  - both copies (multicast commit); tm_empty - leader's copy, 2 x 256 epilogue arrivals (peer: remote)
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include "sgemm.cuh"
#include "tc_ptx.cuh"

namespace cvar {
// 1 (default): the tcgen05 epilogues drain tensor memory into registers and overlap their global-memory work with the
// next tile's main loop; 0: the round-1 epilogue (tensor memory held until the tile is stored).  Same values either way.
static int initial_epi_overlap() {
  const char* e = getenv("CVAR_EPI_OVERLAP");
  return (e != nullptr && e[0] == '0') ? 0 : 1;
}
int g_epi_overlap = initial_epi_overlap();
namespace tc2 {
using namespace cvar::tc;

constexpr int BM = 128;            // rows per CTA (256 per pair)
constexpr int BN = 256;            // columns per pair; each CTA stages BN/2 weight rows
constexpr int BK = 32;            // K-block in 4-byte units: 32 floats or 64 halves = one 128-byte swizzled row
constexpr int kEpiWarps = 8;
constexpr int kTmaWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
constexpr int kThreads = (kEpiWarps + 2) * 32;       // 320
constexpr int kEpiCols = 16, kStagePitch = kEpiCols + 4;
constexpr int kABytes = BM * BK * 4;                 // 16 KiB
constexpr int kBBytes = (BN / 2) * BK * 4;           // 16 KiB
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;   // 64 KiB
constexpr int kStages = 3;
constexpr int kStagingBytes = kEpiWarps * 32 * kStagePitch * 4;
constexpr int kSmem = kStages * kStageBytes + kStagingBytes + 1024 + 1024;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;          // clears the CTA-rank bit of a shared::cluster address (leader's copy)
static int initial_group_m() {
  const char* e = getenv("CVAR_GROUP_M");
  int v = e ? atoi(e) : 1;
  return (v >= 1 && v <= 64) || (v <= -1 && v >= -64) ? v : 1;   // < 0: column groups of -v weight tiles
}
// CVAR_GROUP_M: rasterisation of the pair-tiles (diagnostic knob).  g >= 1: groups of g row tiles, weight tile outer;
// g <= -1: groups of -g weight tiles, row tile outer.  Default 1 = plain row-major: the 24 (fc1) column tiles of a row tile
// run side by side, so A is read from DRAM exactly once (measured: profiles/r02_gemm_traffic.md).  What is re-read is the
// WEIGHT pair: 38 MB at d24 fc1 / fc2 is swept once per wave of 74 pair-tiles and the part's L2 keeps only ~25-30 MB of such
// a cyclic working set next to the A and result streams (each die caches its own copy of data every SM touches).  Larger
// row groups only add A footprint (2.4 GB of reads at 8, 3.1 GB at 32); column groups trade weight re-reads for A re-reads.
int g_group_m = initial_group_m();

// Optional tile trace (diagnostics, cvar_debug_set_trace): CTA 0 stamps clock64() for its first 64 tiles.
// trace[tile * 8 + ev]: 0 MMA thread has tensor memory (tm_empty seen), 1 last MMA of the tile committed,
// 2 epilogue warp 0 sees tm_full, 3 it has released tensor memory, 4 it has stored its slice; 5 TMA thread issued the
// tile's first stage, 6 its last stage.
__device__ long long* g_trace2 = nullptr;
__device__ __forceinline__ void trace2(int tile_i, int ev) {
  if (g_trace2 != nullptr && blockIdx.x == 0 && tile_i < 64) g_trace2[tile_i * 8 + ev] = clock64();
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA into THIS CTA's shared memory, completion counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (F16)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// completion of all prior MMAs of the pair -> the same-offset barrier in BOTH CTAs
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
// arrive on the LEADER's copy of a barrier from either CTA of the pair (cluster-scope release)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}

__device__ __forceinline__ void tile_coords(int tile, int m_tiles, int n_tiles, int kGroupM, int& mt, int& nt) {
  if (kGroupM < 0) {   // column groups: -kGroupM weight tiles stay hot while every row tile passes them (row tile outer, weight tile inner)
    const int per_group = -kGroupM * m_tiles;
    const int g = tile / per_group;
    const int first_n = g * -kGroupM;
    const int gn = min(-kGroupM, n_tiles - first_n);
    const int in_g = tile - g * per_group;
    mt = in_g / gn;
    nt = first_n + (in_g - mt * gn);
    return;
  }
  const int per_group = kGroupM * n_tiles;
  const int g = tile / per_group;
  const int first_m = g * kGroupM;
  const int gm = min(kGroupM, m_tiles - first_m);
  const int in_g = tile - g * per_group;
  nt = in_g / gm;
  mt = first_m + (in_g - nt * gm);
}


// ------------------------------------------------------------------------------------------------ epilogue of one tile
// Each epilogue warp owns 32 accumulator rows (its TMEM lane quarter) and `ncols` (multiple of 16, <= 128) columns
// starting at `col0`.  The accumulators fill all 512 TMEM columns (main + cross), so the next tile's MMAs cannot start
// until this tile has left tensor memory.
//
// Round 1 held tensor memory for the whole epilogue.  The tile trace (tools/gemm_ab.py, profiles/r02_gemm_epilogue.md)
// showed what that costs: the main loop of a K = 1536 tile runs at the full MMA rate (35.3k cycles for 24 K-blocks)
// and is then followed by 27-30k cycles of epilogue with the tensor core idle - NOT the operand-ingest limit round 1
// blamed.  Two epilogues now exist:
//   * "staged" (round 1; TF32 operands, ragged shapes): 16-column chunks, transposed through shared memory.
//   * "row" (FP16-pair operands, N % 16 == 0, 32-byte aligned outputs): the thread pulls its row slice into REGISTERS
//     (main and cross folded: fma(cross, 2^-11, main), as before), releases tensor memory, and then streams along its own
//     output row with 32-byte global accesses - no shared memory (the MMAs read operands from it at ~100 B/clk while the
//     epilogue runs), the epilogue kind a compile-time parameter (the first overlapped version, a fully unrolled copy of
//     the staged epilogue with run-time modes, was 400 KB of SASS and ran at the speed of the instruction cache: 74k
//     cycles per tile).  Values and the order of every floating-point operation are those of the staged epilogue:
//     results are bit-identical (tools/gemm_ab.py checks).
template <class EP>
__device__ __forceinline__ void epi_chunk(const EP& ep, const float* a16, float* stage, int lane, long long m_base, int n0,
                                          long long M, int N) {
#pragma unroll
  for (int q = 0; q < kEpiCols / 4; ++q)
    *reinterpret_cast<float4*>(stage + lane * kStagePitch + q * 4) =
        make_float4(a16[4 * q], a16[4 * q + 1], a16[4 * q + 2], a16[4 * q + 3]);
  __syncwarp();
  const int n = n0 + (lane & 3) * 4;
  const int nvalid = min(4, N - n);
  EpiAux aux[4];
#pragma unroll
  for (int r8 = 0; r8 < 4; ++r8) {
    const long long m = m_base + r8 * 8 + (lane >> 2);
    if (m < M && n < N) aux[r8] = epi_load_aux(ep, m, n, nvalid, 0);
  }
#pragma unroll
  for (int r8 = 0; r8 < 4; ++r8) {
    const int rr = r8 * 8 + (lane >> 2);
    const float4 x = *reinterpret_cast<const float4*>(stage + rr * kStagePitch + (lane & 3) * 4);
    const long long m = m_base + rr;
    if (m < M && n < N) epi_store_aux(ep, m, n, &x.x, nvalid, 0, aux[r8]);
  }
  __syncwarp();
}

template <class EP, bool kCross>
__device__ __forceinline__ void epilogue_tile_staged(const EP& ep, uint32_t tcol, int col0, int ncols, float lo_scale,
                                                     long long m_base, int n_base, long long M, int N, float* stage,
                                                     int lane, uint64_t* tm_empty, int trace_tile) {
  constexpr int kAccStride = 256;
#pragma unroll 1
  for (int c = 0; c < ncols; c += kEpiCols) {
    float v[kEpiCols], w[kEpiCols];
    tmem_ld_32x32b_x16(tcol + (uint32_t)(col0 + c), v);
    if (kCross) {
      tmem_ld_32x32b_x16(tcol + (uint32_t)(kAccStride + col0 + c), w);
#pragma unroll
      for (int i = 0; i < kEpiCols; ++i) v[i] = fmaf(w[i], lo_scale, v[i]);
    }
    epi_chunk(ep, v, stage, lane, m_base, n_base + col0 + c, M, N);
  }
  tc_fence_before();
  mbar_arrive_leader(tm_empty);
  if (threadIdx.x == 0) trace2(trace_tile, 3);
}

// ---- 32-byte global accesses (LDG / STG .256, sm_100): one full sector per thread and instruction
__device__ __forceinline__ void ld8(const float* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
// outputs are streamed (.cs = evict-first): a GEMM output is far larger than L2 and is next read by a different kernel; left
// at normal priority it evicts the activation / weight tiles the other CTAs of the rasterisation group are re-reading
// (ncu on fc1 with plain stores: 2.24 GB of DRAM reads against 0.44 GB of operands)
__device__ __forceinline__ void st8(float* p, const float* v) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void st8_b32(void* p, const uint32_t* v) {
  asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld8_stream(const float* p, float* v) {     // read once (residual / read-modify-write rows)
  asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void ld16f(const float* p, float* v) {
  ld8(p, v);
  ld8(p + 8, v + 8);
}
__device__ __forceinline__ void ld16f_stream(const float* p, float* v) {
  ld8_stream(p, v);
  ld8_stream(p + 8, v + 8);
}
__device__ __forceinline__ void st16f(float* p, const float* v) {
  st8(p, v);
  st8(p + 8, v + 8);
}
// 16 consecutive values -> 32-byte stores of the hi and lo halves (standard pair; bit-identical to split_f16 per element)
__device__ __forceinline__ void st16_split_f16(__half* hi, __half* lo, const float* r) {
  uint32_t ph[8], pl[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float c0 = fminf(fmaxf(r[2 * j], -65504.0f), 65504.0f), c1 = fminf(fmaxf(r[2 * j + 1], -65504.0f), 65504.0f);
    const __half2 h = __floats2half2_rn(c0, c1);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn((c0 - f.x) * kF16LoScale, (c1 - f.y) * kF16LoScale);
    ph[j] = *reinterpret_cast<const uint32_t*>(&h);
    pl[j] = *reinterpret_cast<const uint32_t*>(&l);
  }
  st8_b32(hi, ph);
  st8_b32(lo, pl);
}
// the same for qk pairs (16 x = hi + lo, residual not scaled; bit-identical to split_f16_qk per element)
__device__ __forceinline__ void st16_split_f16_qk(__half* hi, __half* lo, const float* r) {
  uint32_t ph[8], pl[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float c0 = fminf(fmaxf(r[2 * j] * kQkScale, -65504.0f), 65504.0f);
    const float c1 = fminf(fmaxf(r[2 * j + 1] * kQkScale, -65504.0f), 65504.0f);
    const __half2 h = __floats2half2_rn(c0, c1);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(c0 - f.x, c1 - f.y);
    ph[j] = *reinterpret_cast<const uint32_t*>(&h);
    pl[j] = *reinterpret_cast<const uint32_t*>(&l);
  }
  st8_b32(hi, ph);
  st8_b32(lo, pl);
}

// ---- row epilogues: begin(m) once per tile and thread, row16(ctx, n, v) per 16-column chunk (n % 16 == 0, n + 16 <= N)
template <int MODE>
struct DenseRow {
  DenseEpilogue e;
  struct Ctx {
    long long off;
    const float* g;
    const float* rs;
  };
  __device__ __forceinline__ Ctx begin(long long m) const {
    Ctx c;
    c.off = m * e.ldo;
    c.g = (MODE == CVAR_EPI_BIAS_GAMMA_RESID) ? e.gamma + (m / e.rows_per_sample) * e.gamma_row_stride : nullptr;
    c.rs = (MODE == CVAR_EPI_BIAS_RESID) ? e.resid + m * e.ldr : nullptr;
    return c;
  }
  __device__ __forceinline__ void row16(const Ctx& c, int n, const float* v, int /*lc*/) const {
    float r[16], b[16];
    if (e.bias != nullptr) {
      ld16f(e.bias + n, b);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) b[j] = 0.f;
    }
    if (MODE == CVAR_EPI_BIAS_GAMMA_RESID) {
      float x[16], g[16];
      ld16f_stream(e.out + c.off + n, x);
      ld16f(c.g + n, g);
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = __fadd_rn(x[j], __fmul_rn(__fadd_rn(v[j], b[j]), g[j]));   // x + branch.mul(gamma)
    } else if (MODE == CVAR_EPI_BIAS_RESID) {
      float x[16];
      ld16f_stream(c.rs + n, x);
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = __fadd_rn(x[j], __fadd_rn(v[j], b[j]));                   // shortcut + h
    } else if (MODE == CVAR_EPI_BIAS_GELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = gelu_tanh_fast(__fadd_rn(v[j], b[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = __fadd_rn(__fmul_rn(v[j], e.alpha), b[j]);
    }
    if (e.out16_hi != nullptr) {
      st16_split_f16(e.out16_hi + c.off + n, e.out16_lo + c.off + n, r);
      if (e.out == nullptr) return;
    }
    st16f(e.out + c.off + n, r);
  }
};

struct QkvRow {     // FP16-pair outputs only (cvar_qkv_project16)
  QkvEpilogue e;
  struct Ctx {
    int r, t;
  };
  __device__ __forceinline__ Ctx begin(long long m) const {
    Ctx c;
    c.r = (int)(m / e.l);
    c.t = (int)(m - (long long)c.r * e.l);
    return c;
  }
  __device__ __forceinline__ void row16(const Ctx& c, int n, const float* v, int /*lc*/) const {
    const int which = n / e.C;                 // a 16-column chunk stays inside one of q / k / v and inside one head
    const int cc = n - which * e.C;
    const int h = cc >> 6, d = cc & 63;
    const float* bias = which == 0 ? e.q_bias : (which == 1 ? e.k_bias : e.v_bias);
    float o[16], b[16];
    ld16f(bias + cc, b);
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = __fadd_rn(v[j], b[j]);
    const long long rh = (long long)c.r * e.H + h;
    if (which == 0) {
      const long long off = ((rh * e.l + c.t) << 6) + d;
      st16_split_f16_qk(e.q16_hi + off, e.q16_lo + off, o);
    } else if (which == 1) {
      const long long off = ((rh * e.T_max + e.L_prev + c.t) << 6) + d;
      st16_split_f16_qk(e.k16_hi + off, e.k16_lo + off, o);
    } else {
      // V^T: the contiguous index is the token = this thread's row, so the warp's 32 rows write 64 contiguous bytes per
      // head dim (round 1 scattered 2-byte stores from a transposed layout)
      const long long off = (rh * 64 + d) * e.T_max + e.L_prev + c.t;
#pragma unroll
      for (int j = 0; j < 16; ++j) split_f16(o[j], e.vt16_hi[off + (long long)j * e.T_max], e.vt16_lo[off + (long long)j * e.T_max]);
    }
  }
};

// one GroupNorm group of this warp's 32 pixels is complete: reduce over the lanes, lane 0 stores the partial
__device__ __forceinline__ void gn_flush(double* dst, float s, float ss) {
  s = warp_sum(s);
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) {
    dst[0] = (double)s;
    dst[1] = (double)ss;
  }
}

// out_mode 0 (NHWC fp32): bias + optional residual, or the K-split continuation out += acc.
// CPG > 0: also the GroupNorm statistics of the stored values (cvar_conv_args.gn_part), CPG = channels per group, a
// compile-time constant so that the group boundaries inside the thread's row are static (a run-time counter version cost
// the convolution 6 %: 128 compare-and-branch sites and their spills, also when the statistics were off).
template <int CPG>
struct ConvRow {
  ConvEpilogue e;
  struct Ctx {
    long long off;
    double* part;      // this warp's slot of group 0 of its image (CPG > 0)
    float s, ss;       // running sums of the group being crossed
  };
  __device__ __forceinline__ Ctx begin(long long m) const {
    Ctx c;
    c.off = m * e.Cout;
    c.s = c.ss = 0.f;
    c.part = nullptr;
    if (CPG > 0) {
      const long long img = m / e.gn_HW;
      const int slot = (int)(m - img * e.gn_HW) >> 5;
      c.part = e.gn_part + ((img * e.gn_groups) * e.gn_slots + slot) * 2;
    }
    return c;
  }
  // lc: column of v[0] inside the thread's slice (a literal after unrolling); the slice starts on a group boundary
  __device__ __forceinline__ void row16(Ctx& c, int n, const float* v, int lc) const {
    float r[16], x[16];
    if (e.accumulate) {
      ld16f_stream(e.out + c.off + n, x);
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = __fadd_rn(x[j], v[j]);
    } else {
      ld16f(e.bias + n, x);
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = __fadd_rn(v[j], x[j]);
      if (e.resid != nullptr) {
        ld16f_stream(e.resid + c.off + n, x);
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = __fadd_rn(x[j], r[j]);
      }
    }
    st16f(e.out + c.off + n, r);
    if (CPG > 0) {
      // GroupNorm statistics of what was just stored (the consumer's Normalize, vae_modules.py:18-19); every lane is at
      // the same channel, so the flush (a warp reduction) is warp-uniform
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        c.s += r[j];
        c.ss = fmaf(r[j], r[j], c.ss);
        if ((lc + j + 1) % (CPG > 0 ? CPG : 1) == 0) {
          gn_flush(c.part + (long long)((n + j) / (CPG > 0 ? CPG : 1)) * e.gn_slots * 2, c.s, c.ss);
          c.s = c.ss = 0.f;
        }
      }
    }
  }
};

template <class EP> struct IsDenseRow { static constexpr bool value = false; };
template <int MODE> struct IsDenseRow<DenseRow<MODE>> { static constexpr bool value = true; };
template <> struct IsDenseRow<QkvRow> { static constexpr bool value = true; };    // QkvRow works per 16-column chunk on absolute columns: any tile width
template <class EP> struct IsRowEpilogue { static constexpr bool value = false; };
template <int MODE> struct IsRowEpilogue<DenseRow<MODE>> { static constexpr bool value = true; };
template <> struct IsRowEpilogue<QkvRow> { static constexpr bool value = true; };
template <int CPG> struct IsRowEpilogue<ConvRow<CPG>> { static constexpr bool value = true; };

// kChunks: 16-column chunks the thread's slice can have (8 = 128 columns: a 256-wide tile; 5 = 80 columns: the 160-wide
// tiles of the decoder).  A compile-time bound: with a run-time one the register array is sized for 128 columns whatever
// the tile, and the 160-wide convolution kernel spilled its accumulators during the drain.
template <class ROW, bool kCross, int kChunks>
__device__ __forceinline__ void epilogue_tile_rows(const ROW& ep, uint32_t tcol, int col0, int ncols, float lo_scale,
                                                   long long m, long long M, int n_base, int N, uint64_t* tm_empty,
                                                   int trace_tile) {
  constexpr int kAccStride = 256;
  float acc[kChunks * kEpiCols];
  // drain: every main chunk is requested up front and lands in its final registers; the cross chunks follow two at a time
  // through a 32-register temporary and are folded in (4 wait rounds instead of 8: a tcgen05.ld + wait round trip
  // measured ~330 cycles, and the tensor core idles for the whole drain)
#pragma unroll
  for (int c = 0; c < kChunks; ++c)
    if (c * kEpiCols < ncols) tmem_ld16_nowait(tcol + (uint32_t)(col0 + c * kEpiCols), &acc[c * kEpiCols]);
  if (kCross) {
#pragma unroll
    for (int c = 0; c < kChunks; c += 2) {
      if (c * kEpiCols < ncols) {
        float w[2 * kEpiCols];
        tmem_ld16_nowait(tcol + (uint32_t)(kAccStride + col0 + c * kEpiCols), w);
        if (c + 1 < kChunks && (c + 1) * kEpiCols < ncols)
          tmem_ld16_nowait(tcol + (uint32_t)(kAccStride + col0 + (c + 1) * kEpiCols), w + kEpiCols);
        tmem_ld_wait();
        if (c == 0) {
#pragma unroll
          for (int cc = 0; cc < kChunks; ++cc)
            if (cc * kEpiCols < ncols) reg_fence_16(&acc[cc * kEpiCols]);
        }
        reg_fence_16(w);
        reg_fence_16(w + kEpiCols);
#pragma unroll
        for (int i = 0; i < kEpiCols; ++i) acc[c * kEpiCols + i] = fmaf(w[i], lo_scale, acc[c * kEpiCols + i]);
        if (c + 1 < kChunks && (c + 1) * kEpiCols < ncols) {
#pragma unroll
          for (int i = 0; i < kEpiCols; ++i)
            acc[(c + 1) * kEpiCols + i] = fmaf(w[kEpiCols + i], lo_scale, acc[(c + 1) * kEpiCols + i]);
        }
      }
    }
  } else {
    tmem_ld_wait();
#pragma unroll
    for (int cc = 0; cc < kChunks; ++cc)
      if (cc * kEpiCols < ncols) reg_fence_16(&acc[cc * kEpiCols]);
  }
  tc_fence_before();
  mbar_arrive_leader(tm_empty);                         // tensor memory is free: the next tile's MMAs may start
  if (threadIdx.x == 0) trace2(trace_tile, 3);
  if (m >= M) return;                                   // plain loads / stores below: no warp-collective operation
  typename ROW::Ctx ctx = ep.begin(m);
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int n = n_base + col0 + c * kEpiCols;
    if (c * kEpiCols < ncols && n < N) ep.row16(ctx, n, &acc[c * kEpiCols], c * kEpiCols);
  }
}

template <class EP, bool kCross, int kChunks = 8>
__device__ __forceinline__ void epilogue_tile(const EP& ep, uint32_t tcol, int col0, int ncols, float lo_scale,
                                              long long m_base, int n_base, long long M, int N, float* stage, int lane,
                                              uint64_t* tm_empty, int trace_tile) {
  if constexpr (IsRowEpilogue<EP>::value)
    epilogue_tile_rows<EP, kCross, kChunks>(ep, tcol, col0, ncols, lo_scale, m_base + lane, M, n_base, N, tm_empty,
                                            trace_tile);
  else
    epilogue_tile_staged<EP, kCross>(ep, tcol, col0, ncols, lo_scale, m_base, n_base, M, N, stage, lane, tm_empty,
                                     trace_tile);
}

// kFast (cvar_set_fast_mode; NOT a parity mode): the hi halves only - one MMA per product, half the operand bytes per
// stage, twice the stages.
// kBN: columns of the pair tile.  256 everywhere except the NARROW variant (128) that launch() picks for FP16-pair row-epilogue
// GEMMs whose 256-wide tiling would leave more than half of the machine idle (small scales, small batches): the same rows
// of A against half the weight rows - less efficient per tile (the operand ingest per MMA doubles), but twice the tiles.
template <class EP, bool F16, bool kFast = false, int kBN = BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, EP ep,
                long long M, int N, int K, int m_tiles, int n_tiles, int group_m, int dbg_traffic) {
  using G = Geo<BK>;
  // instruction descriptor: D fp32; A/B format 2 = TF32 (kind::tf32) or 0 = FP16 (kind::f16); N, M of the pair tile
  constexpr uint32_t kFmt = F16 ? 0u : 2u;
  constexpr uint32_t kIdesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  constexpr int kBKe = F16 ? 2 * BK : BK;            // K elements per block
  static_assert(kBN == BN || (kBN == 128 && F16 && !kFast), "narrow tiles: FP16 pairs, parity mode");
  constexpr int kBBytesT = (kBN / 2) * BK * 4;       // bytes of the weight tile this CTA stages per operand half
  constexpr float kLoScale = F16 ? (1.0f / 2048.0f) : 1.0f;
  constexpr int kAccStride = 256;                    // main [0,256), lo [256,512)

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int kNS = kFast ? 2 * kStages : kStages;                // pipeline stages
  constexpr int kSB = kFast ? kStageBytes / 2 : kStageBytes;        // bytes per stage (same total)
  auto a_hi = [&](int s) { return smem + s * kSB; };
  auto a_lo = [&](int s) { return smem + s * kSB + kABytes; };                       // !kFast only
  auto b_hi = [&](int s) { return smem + s * kSB + (kFast ? 1 : 2) * kABytes; };
  auto b_lo = [&](int s) { return smem + s * kSB + 2 * kABytes + kBBytes; };         // !kFast only
  float* stage_base = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kStagingBytes);
  uint64_t* full = bars;                  // [S]
  uint64_t* empty = bars + kNS;           // [S]
  uint64_t* tm_full = bars + 2 * kNS;
  uint64_t* tm_empty = bars + 2 * kNS + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kNS + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();           // 0 = leader
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nkb = K / kBKe;
  const int total_tiles = m_tiles * n_tiles;         // m_tiles counts 256-row pair tiles

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&mapAhi), tma_prefetch_desc(&mapAlo), tma_prefetch_desc(&mapBhi), tma_prefetch_desc(&mapBlo);
    for (int s = 0; s < kNS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tm_full, 1);
    mbar_init(tm_empty, 2 * kEpiWarps * 32);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc2(tmem_slot, 512);   // both CTAs, same warp index, same slot offset
  tc_fence_before();
  cluster_sync_all();                                   // barriers of both CTAs initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // ================================================================ epilogue (each CTA drains its own 128 rows)
    float* stage = stage_base + warp * (32 * kStagePitch);
    const int quarter = warp & 3, half = warp >> 2;
    int tcount = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ++tcount) {
      int mt, nt;
      tile_coords(tile, m_tiles, n_tiles, group_m, mt, nt);
      mbar_wait(tm_full, tcount & 1);
      tc_fence_after();
      if (threadIdx.x == 0) trace2(tcount, 2);
      const long long m_base = (long long)mt * 256 + rank * BM + quarter * 32;
      const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16);
      epilogue_tile<EP, !kFast, kBN / 32>(ep, tcol, half * (kBN / 2), kBN / 2, kLoScale, m_base, nt * kBN, M, N, stage, lane,
                                          tm_empty, tcount);
      if (threadIdx.x == 0) trace2(tcount, 4);
    }
  } else if (warp == kTmaWarp) {
    // ================================================================ TMA: own A rows, own half of the weight rows
    if (elect_one()) {
      int it = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        int mt, nt;
        tile_coords(tile, m_tiles, n_tiles, group_m, mt, nt);
        // dbg_traffic (CVAR_DEBUG_TRAFFIC, diagnostics only - WRONG results): 1 = every tile reads the A rows of row tile 0,
        // 2 = every tile reads the weight rows of column tile 0: attributes the DRAM reads of a launch to one operand
        const int arow = ((dbg_traffic & 1) ? 0 : mt * 256) + (int)rank * BM;
        const int brow = ((dbg_traffic & 2) ? 0 : nt * kBN) + (int)rank * (kBN / 2);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kNS;
          const uint32_t ph = (it / kNS) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          if (kb == 0) trace2(it / nkb, 5);
          if (kb == nkb - 1) trace2(it / nkb, 6);
          if (rank == 0)      // bytes of BOTH CTAs (the narrow variant stages half the weight rows in the same stage layout)
            mbar_arrive_expect_tx(&full[s], 2u * (uint32_t)(kBN == BN ? kSB : 2 * kABytes + 2 * kBBytesT));
          tma_load_2d_2sm(&mapAhi, &full[s], a_hi(s), kb * kBKe, arow);
          if (!kFast) tma_load_2d_2sm(&mapAlo, &full[s], a_lo(s), kb * kBKe, arow);
          tma_load_2d_2sm(&mapBhi, &full[s], b_hi(s), kb * kBKe, brow);
          if (!kFast) tma_load_2d_2sm(&mapBlo, &full[s], b_lo(s), kb * kBKe, brow);
        }
      }
    }
  } else if (rank == 0) {
    // ================================================================ MMA issue (leader CTA only)
    if (elect_one()) {
      int it = 0, tcount = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs, ++tcount) {
        mbar_wait(tm_empty, (tcount & 1) ^ 1);
        tc_fence_after();
        trace2(tcount, 0);
        const uint32_t d = tmem_base, dl = tmem_base + (uint32_t)kAccStride;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kNS;
          const uint32_t ph = (it / kNS) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t dah = G::desc(smem_u32(a_hi(s))), dal = G::desc(smem_u32(a_lo(s)));
          const uint64_t dbh = G::desc(smem_u32(b_hi(s))), dbl = G::desc(smem_u32(b_lo(s)));
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {        // one MMA consumes 32 bytes of K: 8 TF32 or 16 FP16 elements
            const uint64_t adv = (uint64_t)(k * 2);
            if (!kFast) {
              umma_2sm<F16>(dl, dal + adv, dbh + adv, kIdesc, (kb | k) != 0);
              umma_2sm<F16>(dl, dah + adv, dbl + adv, kIdesc, 1u);
            }
            umma_2sm<F16>(d, dah + adv, dbh + adv, kIdesc, (kb | k) != 0);
          }
          umma_commit_2sm(&empty[s]);
        }
        umma_commit_2sm(tm_full);
        trace2(tcount, 1);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();                                   // no CTA frees TMEM / exits while its peer may still signal it
  if (warp == kMmaWarp) tmem_dealloc2(tmem_base, 512);
}


// ================================================================================================ conv: implicit GEMM
// 3x3 / 1x1 stride-1 'same' convolution on FP16-pair NHWC activations, all operands by TMA.  The A tile of a K-block is
// a shifted WINDOW of the activation fetched with a 4-D tensor map (C, W, H, B): box = 32 channels x Wb x Hb pixels
// (Wb * Hb = 128 output pixels of one image), coordinates offset by the tap; pixels outside the image are zero-filled
// by the TMA unit, which IS the convolution's zero padding - no im2col buffer and no gather warps.  K runs tap-major
// over (ky, kx, 32-channel chunk), matching cvar_repack_conv_weight.  Decoder channel counts are multiples of 160, not
// of 64, so the shared-memory row is 64 bytes (32 halves, SWIZZLE_64B) and a stage is 4 x 8 KiB; six stages.
constexpr int kCvStages = 6;
constexpr int kCvTile = 128 * 64;                  // 8 KiB: 128 rows of 32 halves (A: pixels; B: up to 128 weight rows)
constexpr int kCvStageBytes = 4 * kCvTile;         // A_hi, A_lo, B_hi, B_lo
constexpr int kCvSmem = kCvStages * kCvStageBytes + kStagingBytes + 1024 + 1024;

struct ConvGeo {
  int H, W, Cin, ks;      // resolution (output == input), input channels, 1 or 3
  int Wb;                 // box width min(W, 128); box height 128 / Wb
  int BN;                 // output channels per pair tile: multiple of 32, <= 256
  int tap0, ntaps;        // this launch covers taps [tap0, tap0 + ntaps) of the ks*ks (K-split: cvar_conv_args.ksplit)
};

__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <class EP, bool kFast = false, int kChunks = 8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tc_conv2_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, EP ep, ConvGeo g,
                long long M, int N, int m_tiles, int n_tiles, int group_m) {
  using G = Geo<16>;                                 // 64-byte rows, SWIZZLE_64B
  const int BNr = g.BN;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(BNr >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // F16 x F16 -> F32
  constexpr int kAccStride = 256;
  constexpr float kLoScale = 1.0f / 2048.0f;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int kNS = kFast ? 2 * kCvStages : kCvStages;
  constexpr int kSB = kFast ? kCvStageBytes / 2 : kCvStageBytes;
  auto a_hi = [&](int s) { return smem + s * kSB; };
  auto a_lo = [&](int s) { return smem + s * kSB + kCvTile; };                        // !kFast only
  auto b_hi = [&](int s) { return smem + s * kSB + (kFast ? 1 : 2) * kCvTile; };
  auto b_lo = [&](int s) { return smem + s * kSB + 3 * kCvTile; };                    // !kFast only
  float* stage_base = reinterpret_cast<float*>(smem + kCvStages * kCvStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCvStages * kCvStageBytes + kStagingBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kNS;
  uint64_t* tm_full = bars + 2 * kNS;
  uint64_t* tm_empty = bars + 2 * kNS + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kNS + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int cpb = g.Cin >> 5;                        // 32-channel chunks per tap
  const int nkb = g.ntaps * cpb;
  const int total_tiles = m_tiles * n_tiles;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&mapAhi), tma_prefetch_desc(&mapAlo), tma_prefetch_desc(&mapBhi), tma_prefetch_desc(&mapBlo);
    for (int s = 0; s < kNS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tm_full, 1);
    mbar_init(tm_empty, 2 * kEpiWarps * 32);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // ================================================================ epilogue: row r of the tile is output pixel m0 + r
    float* stage = stage_base + warp * (32 * kStagePitch);
    const int quarter = warp & 3, half = warp >> 2;
    int tcount = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ++tcount) {
      int mt, nt;
      tile_coords(tile, m_tiles, n_tiles, group_m, mt, nt);
      mbar_wait(tm_full, tcount & 1);
      tc_fence_after();
      if (threadIdx.x == 0) trace2(tcount, 2);
      const long long m_base = (long long)mt * 256 + rank * BM + quarter * 32;
      const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16);
      epilogue_tile<EP, !kFast, kChunks>(ep, tcol, half * (BNr / 2), BNr / 2, kLoScale, m_base, nt * BNr, M, N, stage, lane,
                                         tm_empty, tcount);
      if (threadIdx.x == 0) trace2(tcount, 4);
    }
  } else if (warp == kTmaWarp) {
    // ================================================================ TMA: this CTA's 128 pixels (shifted window per tap)
    if (elect_one()) {
      const int pad = g.ks >> 1;
      const long long HW = (long long)g.H * g.W;
      const uint32_t stage_tx = (kFast ? 1u : 2u) * (2u * (uint32_t)kCvTile + 2u * (uint32_t)(BNr / 2) * 64u);   // both CTAs
      int it = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        int mt, nt;
        tile_coords(tile, m_tiles, n_tiles, group_m, mt, nt);
        const long long m0 = (long long)mt * 256 + (long long)rank * BM;
        const int img = (int)(m0 / HW);
        const int rem = (int)(m0 - (long long)img * HW);
        const int y0 = rem / g.W, x0 = rem - y0 * g.W;
        const int brow = nt * BNr + (int)rank * (BNr / 2);
        for (int tap = g.tap0; tap < g.tap0 + g.ntaps; ++tap) {
          const int ky = tap / g.ks, kx = tap - ky * g.ks;
          int kb = tap * cpb;                      // K-block index in the full tap-major weight matrix
          for (int cc = 0; cc < cpb; ++cc, ++kb, ++it) {
            const int s = it % kNS;
            const uint32_t ph = (it / kNS) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&full[s], stage_tx);
            tma_load_4d_2sm(&mapAhi, &full[s], a_hi(s), cc * 32, x0 + kx - pad, y0 + ky - pad, img);
            if (!kFast) tma_load_4d_2sm(&mapAlo, &full[s], a_lo(s), cc * 32, x0 + kx - pad, y0 + ky - pad, img);
            tma_load_2d_2sm(&mapBhi, &full[s], b_hi(s), kb * 32, brow);
            if (!kFast) tma_load_2d_2sm(&mapBlo, &full[s], b_lo(s), kb * 32, brow);
          }
        }
      }
    }
  } else if (rank == 0) {
    // ================================================================ MMA issue (leader CTA only)
    if (elect_one()) {
      int it = 0, tcount = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs, ++tcount) {
        mbar_wait(tm_empty, (tcount & 1) ^ 1);
        tc_fence_after();
        trace2(tcount, 0);
        const uint32_t d = tmem_base, dl = tmem_base + (uint32_t)kAccStride;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kNS;
          const uint32_t ph = (it / kNS) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t dah = G::desc(smem_u32(a_hi(s))), dal = G::desc(smem_u32(a_lo(s)));
          const uint64_t dbh = G::desc(smem_u32(b_hi(s))), dbl = G::desc(smem_u32(b_lo(s)));
#pragma unroll
          for (int k = 0; k < 2; ++k) {               // 32 halves = two K=16 steps of 32 bytes
            const uint64_t adv = (uint64_t)(k * 2);
            if (!kFast) {
              umma_2sm<true>(dl, dal + adv, dbh + adv, idesc, (kb | k) != 0);
              umma_2sm<true>(dl, dah + adv, dbl + adv, idesc, 1u);
            }
            umma_2sm<true>(d, dah + adv, dbh + adv, idesc, (kb | k) != 0);
          }
          umma_commit_2sm(&empty[s]);
        }
        umma_commit_2sm(tm_full);
        trace2(tcount, 1);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc2(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}
// row-major [rows, K] fp32 (or fp16) matrix -> (128 bytes x 128 rows) box, 128-byte swizzle, out-of-range rows zero-filled
static int make_map(CUtensorMap* map, const void* base, long long rows, int K, long long ld, bool f16, int box_rows = 128) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("tc_gemm2: cuTensorMapEncodeTiled is not available from the driver");
    return -3;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (f16 ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)(f16 ? 2 * BK : BK), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tc_gemm2: cuTensorMapEncodeTiled failed with %d (rows=%lld K=%d ld=%lld)", (int)r, rows, K, ld);
    return -3;
  }
  return 0;
}

// NHWC half activation (B, H, W, C) -> 4-D map (C, W, H, B), box 32 channels x Wb x Hb pixels, 64-byte swizzle
static int make_map_act(CUtensorMap* map, const void* base, int B, int H, int W, int C, int Wb, int Hb) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("tc_conv2: cuTensorMapEncodeTiled is not available from the driver");
    return -3;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)Wb, (cuuint32_t)Hb, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tc_conv2: cuTensorMapEncodeTiled(activation) failed with %d (B=%d H=%d W=%d C=%d box %dx%d)", (int)r, B, H, W,
              C, Wb, Hb);
    return -3;
  }
  return 0;
}
// repacked weight [Cout, K] halves -> (32 x rows) box, 64-byte swizzle
static int make_map_w64(CUtensorMap* map, const void* base, int Cout, int K, int rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("tc_conv2: cuTensorMapEncodeTiled is not available from the driver");
    return -3;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {32, (cuuint32_t)rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tc_conv2: cuTensorMapEncodeTiled(weight) failed with %d (Cout=%d K=%d rows=%d)", (int)r, Cout, K, rows);
    return -3;
  }
  return 0;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <class EP, bool F16>
int launch(const EP& ep, const void* A_hi, const void* A_lo, long long lda, const void* W_hi, const void* W_lo,
           long long ldw, long long M, int N, int K, cudaStream_t s, const char* name) {
  // Narrow tiles (256 x 128) when the 256 x 256 tiling fills less than half of the machine: dense layers of the small scales /
  // small batches.  Same rows of A, same K order per output element: bit-identical results (tests/test_gpu_f16x3.py).
  static const bool narrow_off = getenv("CVAR_NARROW_TILES") != nullptr && getenv("CVAR_NARROW_TILES")[0] == '0';   // A/B
  constexpr bool kCanNarrow = F16 && IsDenseRow<EP>::value;
  const bool narrow = kCanNarrow && !narrow_off && !g_fast_mode && N % 128 == 0 &&
                      2 * (long long)cdiv(M, 256) * cdiv(N, BN) <= num_sms() / 2;
  CUtensorMap mah, mal, mbh, mbl;
  int rc = make_map(&mah, A_hi, M, K, lda, F16);
  if (!rc) rc = make_map(&mal, A_lo, M, K, lda, F16);
  if (!rc) rc = make_map(&mbh, W_hi, N, K, ldw, F16, narrow ? 64 : 128);
  if (!rc) rc = make_map(&mbl, W_lo, N, K, ldw, F16, narrow ? 64 : 128);
  if (rc) return rc;
  // fast mode (NOT a parity mode) exists for the FP16-pair row epilogues only
  auto kern = (F16 && IsRowEpilogue<EP>::value && g_fast_mode) ? tc_gemm2_kernel<EP, F16, F16 && IsRowEpilogue<EP>::value>
                                                               : tc_gemm2_kernel<EP, F16, false>;
  if (narrow) kern = tc_gemm2_kernel<EP, F16, false, kCanNarrow ? 128 : BN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  if (e != cudaSuccess) {
    set_error("%s: cannot raise shared memory to %d: %s", name, kSmem, cudaGetErrorString(e));
    return -2;
  }
  const int m_tiles = cdiv(M, 256), n_tiles = cdiv(N, narrow ? 128 : BN);
  const int pairs = min(num_sms() / 2, m_tiles * n_tiles);
  static const int dbg_traffic = getenv("CVAR_DEBUG_TRAFFIC") ? atoi(getenv("CVAR_DEBUG_TRAFFIC")) : 0;
  kern<<<2 * pairs, kThreads, kSmem, s>>>(mah, mal, mbh, mbl, ep, M, N, K, m_tiles, n_tiles, g_group_m, dbg_traffic);
  CVAR_CHECK_LAUNCH(name);
  return 0;
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }
static bool aligned32(const void* p) { return (((uintptr_t)p) & 31) == 0; }     // nullptr counts as aligned
}  // namespace tc2

int tc2_set_trace(long long* dev_ptr) {
  return cudaMemcpyToSymbol(tc2::g_trace2, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? 0 : -1;
}

// returns 1 when taken, 0 when the shape / operands do not qualify, < 0 on error
static DenseEpilogue dense_epilogue(const cvar_gemm_args* a) {
  DenseEpilogue ep{a->out, a->ldo, a->strideO, a->bias, a->epilogue, a->alpha, a->gamma, a->gamma_row_stride,
                   a->rows_per_sample, a->resid, a->ldr, a->strideR, a->out_lo};
  ep.out16_hi = reinterpret_cast<__half*>(a->out16_hi), ep.out16_lo = reinterpret_cast<__half*>(a->out16_lo);
  return ep;
}

// FP16-pair operands (engine 4): the caller chose the format, so a shape this kernel cannot run is an error, not a
// fall-through.  Small M / N are fine (TMA zero-fills the rows beyond the matrix).
int tc2_gemm_f16(const cvar_gemm_args* a, cudaStream_t s) {
  CVAR_REQUIRE(a->A16_hi && a->A16_lo && a->W16_hi && a->W16_lo, "cvar_gemm[f16x3]: A16_hi/A16_lo/W16_hi/W16_lo must all be set");
  CVAR_REQUIRE(a->batch == 1 && !a->w_is_kn, "cvar_gemm[f16x3]: batched / K-by-N weights are not supported");
  CVAR_REQUIRE(a->K % 64 == 0 && a->N % 4 == 0 && a->lda % 8 == 0 && a->ldw % 8 == 0,
               "cvar_gemm[f16x3]: need K %% 64 == 0, N %% 4 == 0, lda/ldw %% 8 == 0 (K=%d N=%d lda=%lld ldw=%lld)", a->K, a->N,
               a->lda, a->ldw);
  CVAR_REQUIRE(tc2::aligned16(a->A16_hi) && tc2::aligned16(a->A16_lo) && tc2::aligned16(a->W16_hi) && tc2::aligned16(a->W16_lo),
               "cvar_gemm[f16x3]: operands must be 16-byte aligned");
  const DenseEpilogue ep = dense_epilogue(a);
  // row epilogue (tensor memory released before the stores, 32-byte accesses): whole 16-column chunks, 32-byte aligned rows
  const bool rows = g_epi_overlap && a->N % 16 == 0 && a->out_lo == nullptr && tc2::aligned32(a->out) &&
                    tc2::aligned32(a->out16_hi) && tc2::aligned32(a->out16_lo) && tc2::aligned32(a->bias) &&
                    a->ldo % (a->out16_hi != nullptr ? 16 : 8) == 0 &&
                    (a->epilogue != CVAR_EPI_BIAS_GAMMA_RESID || (tc2::aligned32(a->gamma) && a->gamma_row_stride % 8 == 0)) &&
                    (a->epilogue != CVAR_EPI_BIAS_RESID || (tc2::aligned32(a->resid) && a->ldr % 8 == 0));
#define CVAR_TC2_ROWS(MODE)                                                                                              \
  return tc2::launch<tc2::DenseRow<MODE>, true>(tc2::DenseRow<MODE>{ep}, a->A16_hi, a->A16_lo, a->lda, a->W16_hi,         \
                                                a->W16_lo, a->ldw, (long long)a->M, a->N, a->K, s, "cvar_gemm[tc2/f16x3/rows]")
  if (rows) {
    switch (a->epilogue) {
      case CVAR_EPI_BIAS: CVAR_TC2_ROWS(CVAR_EPI_BIAS);
      case CVAR_EPI_BIAS_GELU: CVAR_TC2_ROWS(CVAR_EPI_BIAS_GELU);
      case CVAR_EPI_BIAS_GAMMA_RESID: CVAR_TC2_ROWS(CVAR_EPI_BIAS_GAMMA_RESID);
      case CVAR_EPI_BIAS_RESID: CVAR_TC2_ROWS(CVAR_EPI_BIAS_RESID);
      default: break;
    }
  }
#undef CVAR_TC2_ROWS
  return tc2::launch<DenseEpilogue, true>(ep, a->A16_hi, a->A16_lo, a->lda, a->W16_hi, a->W16_lo, a->ldw,
                                          (long long)a->M, a->N, a->K, s, "cvar_gemm[tc2/f16x3]");
}

int tc2_qkv_f16(const void* A_hi, const void* A_lo, const void* W_hi, const void* W_lo, const QkvEpilogue& ep, int M, int C,
                cudaStream_t s) {
  CVAR_REQUIRE(A_hi && A_lo && W_hi && W_lo, "cvar_qkv_project[f16x3]: A16_hi/A16_lo/W16_hi/W16_lo must all be set");
  CVAR_REQUIRE(tc2::aligned16(A_hi) && tc2::aligned16(A_lo) && tc2::aligned16(W_hi) && tc2::aligned16(W_lo),
               "cvar_qkv_project[f16x3]: operands must be 16-byte aligned");
  const bool rows = g_epi_overlap && ep.q16_hi != nullptr && C % 64 == 0 && tc2::aligned32(ep.q16_hi) &&
                    tc2::aligned32(ep.q16_lo) && tc2::aligned32(ep.k16_hi) && tc2::aligned32(ep.k16_lo) &&
                    tc2::aligned32(ep.q_bias) && tc2::aligned32(ep.k_bias) && tc2::aligned32(ep.v_bias);
  if (rows)
    return tc2::launch<tc2::QkvRow, true>(tc2::QkvRow{ep}, A_hi, A_lo, C, W_hi, W_lo, C, (long long)M, 3 * C, C, s,
                                          "cvar_qkv_project[tc2/f16x3/rows]");
  return tc2::launch<QkvEpilogue, true>(ep, A_hi, A_lo, C, W_hi, W_lo, C, (long long)M, 3 * C, C, s,
                                        "cvar_qkv_project[tc2/f16x3]");
}

// FP16-pair convolution (engine 4): 0 ok, < 0 error.  cvar_conv2d_f16_supported() tells the host in advance.
static int conv_f16_bn(int Cout) {
  const int cand[] = {256, 160, 128, 224, 192, 96, 64, 32};
  for (int bn : cand)
    if (Cout % bn == 0) return bn;
  return 0;
}
int tc2_conv_f16_supported(int H, int W, int Cin, int Cout, int ks) {
  if (ks != 1 && ks != 3) return 0;
  if (Cin % 32 != 0 || conv_f16_bn(Cout) == 0) return 0;
  const int Wb = W < 128 ? W : 128;
  if (W <= 0 || (W < 128 ? (128 % W != 0) : (W % 128 != 0))) return 0;
  const int Hb = 128 / Wb;
  if (H % Hb != 0) return 0;
  return 1;
}
// can the row epilogue of this layer emit GroupNorm partials?  whole groups inside each warp's column half, whole warps
// inside an image
int tc2_conv_f16_gn_fusable(int H, int W, int Cin, int Cout, int ks, int groups) {
  if (!g_epi_overlap || !tc2_conv_f16_supported(H, W, Cin, Cout, ks)) return 0;
  if (groups <= 0 || Cout % groups != 0 || Cout % 16 != 0 || (H * W) % 32 != 0) return 0;
  const int bn = conv_f16_bn(Cout), cpg = Cout / groups;
  if ((cpg != 5 && cpg != 10 && cpg != 20) || bn != 160) return 0;   // the instantiated kernels: 160 / 320 / 640 channels, 160-wide tiles
  return (bn / 2) % cpg == 0 && (bn / 2) % 16 == 0 ? 1 : 0;
}
int tc2_conv_f16(const cvar_conv_args* a, cudaStream_t s) {
  CVAR_REQUIRE(a->x16_hi && a->x16_lo && a->w16_hi && a->w16_lo, "cvar_conv2d[f16x3]: x16_hi/x16_lo/w16_hi/w16_lo must all be set");
  CVAR_REQUIRE(!a->upsample2x && a->in_a == nullptr,
               "cvar_conv2d[f16x3]: the FP16-pair input is taken as is (upsample / normalise it when producing the pair)");
  const int H = a->Hin, W = a->Win;
  CVAR_REQUIRE(tc2_conv_f16_supported(H, W, a->Cin, a->Cout, a->ks),
               "cvar_conv2d[f16x3]: unsupported shape H=%d W=%d Cin=%d Cout=%d ks=%d", H, W, a->Cin, a->Cout, a->ks);
  CVAR_REQUIRE(tc2::aligned16(a->x16_hi) && tc2::aligned16(a->x16_lo) && tc2::aligned16(a->w16_hi) && tc2::aligned16(a->w16_lo),
               "cvar_conv2d[f16x3]: operands must be 16-byte aligned");
  tc2::ConvGeo g;
  g.H = H, g.W = W, g.Cin = a->Cin, g.ks = a->ks;
  g.Wb = W < 128 ? W : 128;
  g.BN = conv_f16_bn(a->Cout);
  const int Hb = 128 / g.Wb;
  const int K = a->ks * a->ks * a->Cin;
  CUtensorMap mah, mal, mbh, mbl;
  int rc = tc2::make_map_act(&mah, a->x16_hi, a->B, H, W, a->Cin, g.Wb, Hb);
  if (!rc) rc = tc2::make_map_act(&mal, a->x16_lo, a->B, H, W, a->Cin, g.Wb, Hb);
  if (!rc) rc = tc2::make_map_w64(&mbh, a->w16_hi, a->Cout, K, g.BN / 2);
  if (!rc) rc = tc2::make_map_w64(&mbl, a->w16_lo, a->Cout, K, g.BN / 2);
  if (rc) return rc;
  ConvEpilogue ep{a->out, a->bias, a->resid, a->Cout, a->out_mode, H, W, a->out_rows_total, a->row_offset};
  ep.out_samples = a->out_samples;
  CVAR_REQUIRE(a->gn_part == nullptr || tc2_conv_f16_gn_fusable(H, W, a->Cin, a->Cout, a->ks, a->gn_groups),
               "cvar_conv2d[f16x3]: gn_part set but this layer's epilogue cannot produce GroupNorm statistics "
               "(cvar_conv2d_gn_fusable: H=%d W=%d Cout=%d groups=%d)", H, W, a->Cout, a->gn_groups);
  const bool rows = g_epi_overlap && a->out_mode == 0 && a->Cout % 16 == 0 && tc2::aligned32(a->out) &&
                    tc2::aligned32(a->bias) && tc2::aligned32(a->resid) && a->bias != nullptr;
  auto kern_staged = tc2::tc_conv2_kernel<ConvEpilogue>;
  // the thread's slice is BN / 2 columns: five 16-column chunks for the decoder's 160-wide tiles, eight for 256
  const bool narrow = g.BN <= 160;
  auto kern_rows = narrow ? (g_fast_mode ? tc2::tc_conv2_kernel<tc2::ConvRow<0>, true, 5> : tc2::tc_conv2_kernel<tc2::ConvRow<0>, false, 5>)
                          : (g_fast_mode ? tc2::tc_conv2_kernel<tc2::ConvRow<0>, true, 8> : tc2::tc_conv2_kernel<tc2::ConvRow<0>, false, 8>);
  // GroupNorm partials exist for 160 / 320 / 640 channels only: always 160-wide tiles
  auto kern_gn5 = g_fast_mode ? tc2::tc_conv2_kernel<tc2::ConvRow<5>, true, 5> : tc2::tc_conv2_kernel<tc2::ConvRow<5>, false, 5>;
  auto kern_gn10 = g_fast_mode ? tc2::tc_conv2_kernel<tc2::ConvRow<10>, true, 5> : tc2::tc_conv2_kernel<tc2::ConvRow<10>, false, 5>;
  auto kern_gn20 = g_fast_mode ? tc2::tc_conv2_kernel<tc2::ConvRow<20>, true, 5> : tc2::tc_conv2_kernel<tc2::ConvRow<20>, false, 5>;
  const int cpg = a->gn_part != nullptr ? a->Cout / a->gn_groups : 0;
  cudaError_t e = cudaFuncSetAttribute(kern_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kCvSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kern_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kCvSmem);
  if (e == cudaSuccess && cpg == 5) e = cudaFuncSetAttribute(kern_gn5, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kCvSmem);
  if (e == cudaSuccess && cpg == 10) e = cudaFuncSetAttribute(kern_gn10, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kCvSmem);
  if (e == cudaSuccess && cpg == 20) e = cudaFuncSetAttribute(kern_gn20, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kCvSmem);
  CVAR_REQUIRE(e == cudaSuccess, "cvar_conv2d[f16x3]: cannot raise shared memory to %d: %s", tc2::kCvSmem, cudaGetErrorString(e));
  const long long M = (long long)a->B * H * W;
  const int m_tiles = cdiv(M, 256), n_tiles = a->Cout / g.BN;
  const int pairs = min(tc2::num_sms() / 2, m_tiles * n_tiles);
  // K-split: the tensor core truncates its fp32 accumulator after every MMA, so the error of one accumulation grows with
  // its length (K = 9 * 640 is 360 steps).  ksplit = 3 runs the three kernel rows as three launches whose partial sums
  // meet in the fp32 output with round-to-nearest adds (the epilogue of launches 2, 3 adds to what is there).
  const int taps = a->ks * a->ks;
  const int nsplit = (a->ksplit == 3 && a->ks == 3 && a->out_mode == 0) ? 3 : 1;
  CVAR_REQUIRE(a->ksplit == 0 || a->ksplit == 1 || nsplit == 3, "cvar_conv2d[f16x3]: ksplit = 3 needs ks = 3 and out_mode 0");
  for (int part = 0; part < nsplit; ++part) {
    g.ntaps = taps / nsplit;
    g.tap0 = part * g.ntaps;
    ConvEpilogue epp = ep;
    epp.accumulate = part > 0 ? 1 : 0;
    if (a->gn_part != nullptr && part == nsplit - 1) {       // the launch that stores the final values
      CVAR_REQUIRE(rows, "cvar_conv2d[f16x3]: gn_part needs the row epilogue (32-byte aligned out / bias / resid)");
      epp.gn_part = a->gn_part, epp.gn_groups = a->gn_groups, epp.gn_cpg = a->Cout / a->gn_groups;
      epp.gn_HW = H * W, epp.gn_slots = (H * W) / 32;
    }
    if (rows && epp.gn_part != nullptr && cpg == 5)
      kern_gn5<<<2 * pairs, tc2::kThreads, tc2::kCvSmem, s>>>(mah, mal, mbh, mbl, tc2::ConvRow<5>{epp}, g, M, a->Cout, m_tiles,
                                                               n_tiles, tc2::g_group_m);
    else if (rows && epp.gn_part != nullptr && cpg == 10)
      kern_gn10<<<2 * pairs, tc2::kThreads, tc2::kCvSmem, s>>>(mah, mal, mbh, mbl, tc2::ConvRow<10>{epp}, g, M, a->Cout, m_tiles,
                                                                n_tiles, tc2::g_group_m);
    else if (rows && epp.gn_part != nullptr && cpg == 20)
      kern_gn20<<<2 * pairs, tc2::kThreads, tc2::kCvSmem, s>>>(mah, mal, mbh, mbl, tc2::ConvRow<20>{epp}, g, M, a->Cout, m_tiles,
                                                                n_tiles, tc2::g_group_m);
    else if (rows)
      kern_rows<<<2 * pairs, tc2::kThreads, tc2::kCvSmem, s>>>(mah, mal, mbh, mbl, tc2::ConvRow<0>{epp}, g, M, a->Cout, m_tiles,
                                                                n_tiles, tc2::g_group_m);
    else
      kern_staged<<<2 * pairs, tc2::kThreads, tc2::kCvSmem, s>>>(mah, mal, mbh, mbl, epp, g, M, a->Cout, m_tiles, n_tiles, tc2::g_group_m);
    CVAR_CHECK_LAUNCH("cvar_conv2d[tc2/f16x3]");
  }
  return 0;
}

int tc2_gemm_try(const cvar_gemm_args* a, cudaStream_t s) {
  if (a->A_lo == nullptr || a->W_hi == nullptr || a->W_lo == nullptr) return 0;
  if (a->batch != 1 || a->w_is_kn || a->M < 256 || a->N < 256 || a->K % 32 != 0 || a->N % 4 != 0) return 0;
  if (a->lda % 4 != 0 || a->ldw % 4 != 0 || !tc2::aligned16(a->A) || !tc2::aligned16(a->A_lo) ||
      !tc2::aligned16(a->W_hi) || !tc2::aligned16(a->W_lo))
    return 0;
  int rc = tc2::launch<DenseEpilogue, false>(dense_epilogue(a), a->A, a->A_lo, a->lda, a->W_hi, a->W_lo, a->ldw,
                                             (long long)a->M, a->N, a->K, s, "cvar_gemm[tc2]");
  return rc ? rc : 1;
}

int tc2_qkv_try(const float* A_hi, const float* A_lo, const float* Wqkv_hi, const float* Wqkv_lo, const QkvEpilogue& ep,
                int M, int C, cudaStream_t s) {
  if (A_lo == nullptr || M < 256 || C % 32 != 0) return 0;
  if (!tc2::aligned16(A_hi) || !tc2::aligned16(A_lo) || !tc2::aligned16(Wqkv_hi) || !tc2::aligned16(Wqkv_lo)) return 0;
  int rc = tc2::launch<QkvEpilogue, false>(ep, A_hi, A_lo, C, Wqkv_hi, Wqkv_lo, C, (long long)M, 3 * C, C, s,
                                           "cvar_qkv_project[tc2]");
  return rc ? rc : 1;
}
}  // namespace cvar
