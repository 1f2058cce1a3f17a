// tcgen05 / TMEM / TMA GEMM engine for sm_100a: out = epilogue(A[M,K] * W[N,K]^T) at fp32-class accuracy.
//
// Arithmetic: error-compensated 3xTF32.  Every fp32 operand x is split into hi = x with the 13 low mantissa bits
// cleared (exactly what a TF32 tensor-core input keeps) and lo = x - hi (exact in fp32); the product is accumulated as
//   a_lo*b_hi + a_hi*b_lo + a_hi*b_hi      (three tcgen05.mma.kind::tf32 per 8-wide k-step, fp32 accumulate in TMEM).
// The dropped a_lo*b_lo term is below 2^-20 relative.  Reference inference is fp32, so bf16
// operands would break token parity (section 7.2); this keeps the tensor cores AND the parity contract.
//
// Structure (one persistent CTA per SM, 14 warps, warp-specialised, everything handed over through mbarriers):
//   warps 0-3   epilogue: tcgen05.ld the 128 x BN fp32 accumulator (row = TMEM lane = thread), apply the epilogue
//               functor (bias / GELU / gamma-residual / QKV scatter into the KV cache / conv NHWC+residual / image)
//   warps 4-11  A producers: LDG the activation tile (dense rows, or the implicit-GEMM gather of a 3x3 convolution with
//               optional nearest-x2 upsampling) through a register prefetch ring, split hi/lo, store both into
//               128B/64B-swizzled K-major shared-memory tiles (software swizzle = the pattern TMA / UMMA expect)
//   warp 12     TMA issue (one elected lane): the weights are split hi/lo ONCE at pack time (cvar_split_tf32), so both
//               weight tiles arrive by cp.async.bulk.tensor with hardware swizzle and need no SM work at all
//   warp 13     TMEM allocation + MMA issue (one elected lane): tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = BN
// (A first version converted the weight tile in shared memory with 4 more warps: the pipeline trace in
//  profiles/r01_gemm_trace.md showed that conversion - 1.6k cycles per K-block, fighting the MMA operand reads for
//  shared-memory bandwidth - was the critical path, tensor pipe 47 % busy.)
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include "sgemm.cuh"
#include "tc_ptx.cuh"

namespace cvar {
int tc2_set_trace(long long* dev_ptr);
namespace tc {

constexpr int BM = 128;
constexpr int kEpiWarps = 8, kProdWarps = 8;
constexpr int kTmaWarp = kEpiWarps + kProdWarps, kMmaWarp = kTmaWarp + 1;
constexpr int kThreads = (kEpiWarps + kProdWarps + 2) * 32;   // 576
constexpr int kEpiCols = 16;                                  // accumulator columns per epilogue step
constexpr int kStagePitch = kEpiCols + 4;                     // floats per row of an epilogue transposition tile
constexpr int kMaxSmem = 227 * 1024;

template <int BN, int BK>
struct SmemPlan {
  static constexpr int kABytes = BM * BK * 4;
  static constexpr int kBBytes = BN * BK * 4;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kStagingBytes = kEpiWarps * 32 * kStagePitch * 4;        // 18 KiB
  static constexpr int kFixed = kBarrierBytes + kStagingBytes + 1024;           // +1024: manual 1 KiB alignment
  static constexpr int kStages = (kMaxSmem - kFixed) / kStageBytes > 6 ? 6 : (kMaxSmem - kFixed) / kStageBytes;
  static constexpr int kTotal = kStages * kStageBytes + kFixed;
  static_assert(kStages >= 2, "tile too large for shared memory");
};

// Optional pipeline trace (diagnostics only): CTA 0 stamps clock64() at the hand-over points of its first 64 K-blocks.
// Layout: trace[(role * 64 + kb) * 2 + event]; roles: 0 TMA, 2 A producer, 3 MMA issue.
__device__ long long* g_trace = nullptr;
__device__ __forceinline__ void trace_stamp(int role, int it, int ev) {
  if (g_trace != nullptr && blockIdx.x == 0 && it < 64) g_trace[(role * 64 + it) * 2 + ev] = clock64();
}

// L2-aware tile order.  The persistent loop hands consecutive tile indices to the CTAs that run concurrently, so the
// index -> (m, n) map decides the cache footprint of a wave.  n-fastest (the first version) makes every wave sweep ALL
// weight strips: ncu measured 6.3 GB of DRAM reads for 0.48 GB of operands on fc1 (profiles/r01_ncu_summary.md).
// Grouped order (kGroupM m-tiles per group, m fastest inside a group): a wave touches kGroupM activation strips, which
// stay resident while the group walks across n, and each weight strip is fetched once per group.
constexpr int kGroupM = 32;
__device__ __forceinline__ void tile_coords(int tile, int m_tiles, int n_tiles, int& mt, int& nt) {
  const int per_group = kGroupM * n_tiles;
  const int g = tile / per_group;
  const int first_m = g * kGroupM;
  const int gm = min(kGroupM, m_tiles - first_m);
  const int in_g = tile - g * per_group;
  nt = in_g / gm;
  mt = first_m + (in_g - nt * gm);
}

// -------------------------------------------------------------------------------------------------- the kernel
template <int BN, int BK, class AL, class EP>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmapBhi, const __grid_constant__ CUtensorMap tmapBlo, AL al, EP ep,
               long long M, int N, int K, int m_tiles, int n_tiles) {
  using G = Geo<BK>;
  using P = SmemPlan<BN, BK>;
  constexpr int S = P::kStages;
  // TMEM plan (512 columns): every tile owns TWO fp32 accumulators - 'main' takes a_hi*b_hi, 'lo' takes the two small
  // cross terms.  The tensor core truncates (round-toward-zero) the accumulator after every instruction, a bias that
  // grows with the number of accumulation steps and with |accumulator|; keeping the 2^-11-sized terms out of the main
  // accumulator cuts its steps by 3x (measured on B200, profiles/r01_tc_accuracy.md: K=1536 error 1.1e-5 -> 3.5e-6).
  // BN <= 128: two such pairs (epilogue of tile i overlaps MMAs of tile i+1); BN > 128: one pair.
  constexpr int kAccStride = (BN <= 128) ? 128 : 256;      // columns between main and lo
  constexpr int kAccBufs = (BN <= 128) ? 2 : 1;
  constexpr int kTmemCols = 512;
  static_assert(BN <= 256 && BN % 16 == 0 && BN >= 16, "BN");
  constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  auto a_hi = [&](int s) { return smem + s * P::kStageBytes; };
  auto a_lo = [&](int s) { return smem + s * P::kStageBytes + P::kABytes; };
  auto b_hi = [&](int s) { return smem + s * P::kStageBytes + 2 * P::kABytes; };
  auto b_lo = [&](int s) { return smem + s * P::kStageBytes + 2 * P::kABytes + P::kBBytes; };
  float* stage_base = reinterpret_cast<float*>(smem + S * P::kStageBytes);           // epilogue transposition tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * P::kStageBytes + P::kStagingBytes);
  uint64_t* b_full = bars;                 // [S]  TMA bytes of the two weight tiles (hi, lo) landed
  uint64_t* a_ready = bars + S;            // [S]  hi/lo tiles of A written by the producers (256 arrivals)
  uint64_t* empty = bars + 2 * S;          // [S]  MMAs reading the stage retired (tcgen05.commit)
  uint64_t* tm_full = bars + 3 * S;        // [2]  accumulator complete
  uint64_t* tm_empty = bars + 3 * S + 2;   // [2]  accumulator drained by the epilogue (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = K / BK;
  const int total_tiles = m_tiles * n_tiles;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmapBhi);
    tma_prefetch_desc(&tmapBlo);
    for (int s = 0; s < S; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&a_ready[s], kProdWarps * 32);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < kAccBufs; ++a) {
      mbar_init(&tm_full[a], 1);
      mbar_init(&tm_empty[a], kEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // ================================================================ epilogue: TMEM -> registers -> smem -> global
    // tcgen05.ld hands every thread one accumulator ROW.  Writing rows straight from that layout would touch 32
    // different lines per store instruction, so each warp transposes a 32 x 16 chunk through a private shared-memory
    // tile and walks it 8 rows at a time (4 lanes x float4 = 64 contiguous bytes per row).  Warps w and w+4 share a TMEM
    // lane quarter and split the tile's columns.  All global reads of a pass (old x, gamma / residual) are issued
    // before its math and stores, so they overlap instead of serialising.
    static_assert((BN / 2) % kEpiCols == 0, "column split");
    float* stage = stage_base + warp * (32 * kStagePitch);
    const int quarter = warp & 3, half = warp >> 2;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      int mt, nt;
      tile_coords(tile, m_tiles, n_tiles, mt, nt);
      const int acc = tcount % kAccBufs;
      const uint32_t acc_phase = (tcount / kAccBufs) & 1;
      mbar_wait(&tm_full[acc], acc_phase);
      tc_fence_after();
      const long long m_base = (long long)mt * BM + quarter * 32;
      const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 2 * kAccStride);
#pragma unroll 1
      for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += kEpiCols) {
        float v[kEpiCols], w[kEpiCols];
        tmem_ld_32x32b_x16(tcol + (uint32_t)c, v);
        tmem_ld_32x32b_x16(tcol + (uint32_t)(kAccStride + c), w);
#pragma unroll
        for (int q = 0; q < kEpiCols / 4; ++q)       // main + small terms: one rounded fp32 add per element
          *reinterpret_cast<float4*>(stage + lane * kStagePitch + q * 4) =
              make_float4(v[4 * q] + w[4 * q], v[4 * q + 1] + w[4 * q + 1], v[4 * q + 2] + w[4 * q + 2],
                          v[4 * q + 3] + w[4 * q + 3]);
        __syncwarp();
        const int n = nt * BN + c + (lane & 3) * 4;
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
      tc_fence_before();
      mbar_arrive(&tm_empty[acc]);
    }
  } else if (warp < kEpiWarps + kProdWarps) {
    // ================================================================ A producers: LDG (+gather) -> hi/lo tiles
    constexpr int kPT = kProdWarps * 32;                       // 128 producer threads
    constexpr int kVec = BM * BK / 4 / kPT;                    // float4 per thread per k-block: 8 (BK=32) or 4 (BK=16)
    constexpr int kCh = G::kChunks;
    // Register prefetch ring: the loads of K-block kb + PD are issued before K-block kb is converted, so PD blocks of
    // MMA time cover the HBM latency of the activation stream.
    constexpr int PD = (BK == 32) ? 2 : 4;
    const int t = threadIdx.x - kEpiWarps * 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int mt, nt_unused;
      tile_coords(tile, m_tiles, n_tiles, mt, nt_unused);
#pragma unroll
      for (int i = 0; i < kVec; ++i) al.prep(i, (long long)mt * BM + (i * kPT + t) / kCh, 0);
      float4 ring[PD + 1][kVec];
#pragma unroll
      for (int d = 0; d < PD; ++d) {
        if (d < nkb) {
          al.begin_block(d * BK);
#pragma unroll
          for (int i = 0; i < kVec; ++i) ring[d][i] = al.fetch_blk(i, d * BK, ((i * kPT + t) % kCh) * 4);
        }
      }
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        if (kb + PD < nkb) {
          al.begin_block((kb + PD) * BK);
#pragma unroll
          for (int i = 0; i < kVec; ++i) ring[PD][i] = al.fetch_blk(i, (kb + PD) * BK, ((i * kPT + t) % kCh) * 4);
        }
        mbar_wait(&empty[s], ph ^ 1);
        if (t == 0) trace_stamp(2, it, 0);
        unsigned char* hi = a_hi(s);
        unsigned char* lo = a_lo(s);
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
          const int item = i * kPT + t;
          const uint32_t off = G::offset(item / kCh, item % kCh);
          float4 x = ring[0][i], h, l;
          h.x = trunc_tf32(x.x), h.y = trunc_tf32(x.y), h.z = trunc_tf32(x.z), h.w = trunc_tf32(x.w);
          l.x = x.x - h.x, l.y = x.y - h.y, l.z = x.z - h.z, l.w = x.w - h.w;
          *reinterpret_cast<float4*>(hi + off) = h;
          *reinterpret_cast<float4*>(lo + off) = l;
        }
        fence_proxy_async();
        mbar_arrive(&a_ready[s]);
        if (t == 0) trace_stamp(2, it, 1);
#pragma unroll
        for (int d = 0; d < PD; ++d)
#pragma unroll
          for (int i = 0; i < kVec; ++i) ring[d][i] = ring[d + 1][i];
      }
    }
  } else if (warp == kTmaWarp) {
    // ================================================================ TMA issue (pre-split weights: hi and lo tiles)
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int mt, nt;
        tile_coords(tile, m_tiles, n_tiles, mt, nt);
        (void)mt;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          trace_stamp(0, it, 0);
          mbar_arrive_expect_tx(&b_full[s], 2u * (uint32_t)P::kBBytes);
          tma_load_2d(&tmapBhi, &b_full[s], b_hi(s), kb * BK, nt * BN);
          tma_load_2d(&tmapBlo, &b_full[s], b_lo(s), kb * BK, nt * BN);
        }
      }
    }
  } else {
    // ================================================================ MMA issue
    if (lane == 0) {
      int it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const int acc = tcount % kAccBufs;
        const uint32_t acc_phase = (tcount / kAccBufs) & 1;
        mbar_wait(&tm_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * 2 * kAccStride);
        const uint32_t dl = d + (uint32_t)kAccStride;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&b_full[s], ph);
          trace_stamp(3, it, 0);
          mbar_wait(&a_ready[s], ph);
          trace_stamp(1, it, 0);
          tc_fence_after();
          const uint64_t dah = G::desc(smem_u32(a_hi(s))), dal = G::desc(smem_u32(a_lo(s)));
          const uint64_t dbh = G::desc(smem_u32(b_hi(s))), dbl = G::desc(smem_u32(b_lo(s)));
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);          // 8 fp32 = 32 bytes = 2 x 16 B along K inside the atom
            umma_tf32(dl, dal + adv, dbh + adv, kIdesc, (kb | k) != 0);
            umma_tf32(dl, dah + adv, dbl + adv, kIdesc, 1u);
            umma_tf32(d, dah + adv, dbh + adv, kIdesc, (kb | k) != 0);
          }
          umma_commit(&empty[s]);
          trace_stamp(3, it, 1);
        }
        umma_commit(&tm_full[acc]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

// elementwise hi/lo split of a weight tensor (done once, when weights are packed)
__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 x = ld4(w + i * 4), h, l;
  h.x = trunc_tf32(x.x), h.y = trunc_tf32(x.y), h.z = trunc_tf32(x.z), h.w = trunc_tf32(x.w);
  l.x = x.x - h.x, l.y = x.y - h.y, l.z = x.z - h.z, l.w = x.w - h.w;
  st4(hi + i * 4, h);
  st4(lo + i * 4, l);
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

// weight matrix W[N, K] fp32 row-major -> 2-D tensor map with a (BK x BN) box, hardware swizzle matching Geo<BK>
static int make_weight_map(CUtensorMap* map, const float* W, int N, int K, long long ldw, int BN, int BK) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("tc_gemm: cuTensorMapEncodeTiled is not available from the driver");
    return -3;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)ldw * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(W), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tc_gemm: cuTensorMapEncodeTiled failed with %d (N=%d K=%d ldw=%lld)", (int)r, N, K, ldw);
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

int g_tc_bk = 32;

int set_trace(long long* dev_ptr) {
  if (cvar::tc2_set_trace(dev_ptr) != 0) return -1;           // the 2-CTA kernels share the switch (own layout)
  return cudaMemcpyToSymbol(g_trace, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? 0 : -1;
}   // K-block of the engine: 32 (SWIZZLE_128B) or 16 (SWIZZLE_64B, deeper pipeline)

template <int BN, int BK, class AL, class EP>
static int launch_tc(const AL& al, const EP& ep, const float* W_hi, const float* W_lo, long long ldw, long long M, int N,
                     int K, cudaStream_t s, const char* name) {
  CUtensorMap map_hi, map_lo;
  int rc = make_weight_map(&map_hi, W_hi, N, K, ldw, BN, BK);
  if (rc) return rc;
  rc = make_weight_map(&map_lo, W_lo, N, K, ldw, BN, BK);
  if (rc) return rc;
  using P = SmemPlan<BN, BK>;
  auto kern = tc_gemm_kernel<BN, BK, AL, EP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::kTotal);
  if (e != cudaSuccess) {
    set_error("%s: cannot raise shared memory to %d: %s", name, P::kTotal, cudaGetErrorString(e));
    return -2;
  }
  const int m_tiles = cdiv(M, BM), n_tiles = cdiv(N, BN);
  const int grid = min(num_sms(), m_tiles * n_tiles);
  kern<<<grid, kThreads, P::kTotal, s>>>(map_hi, map_lo, al, ep, M, N, K, m_tiles, n_tiles);
  CVAR_CHECK_LAUNCH(name);
  return 0;
}

template <class AL, class EP>
static int dispatch_tc(const AL& al, const EP& ep, const float* W_hi, const float* W_lo, long long ldw, long long M, int N,
                       int K, cudaStream_t s, const char* name) {
  // BN: 256 for the wide transformer layers, 160 for the decoder's 160/320/640-channel convolutions, 128 otherwise
  const bool bk16 = (g_tc_bk == 16);
  if (N % 160 == 0 && N % 256 != 0 && N <= 640) {
    return bk16 ? launch_tc<160, 16>(al, ep, W_hi, W_lo, ldw, M, N, K, s, name) : launch_tc<160, 32>(al, ep, W_hi, W_lo, ldw, M, N, K, s, name);
  }
  if (N >= 256) {
    return bk16 ? launch_tc<256, 16>(al, ep, W_hi, W_lo, ldw, M, N, K, s, name) : launch_tc<256, 32>(al, ep, W_hi, W_lo, ldw, M, N, K, s, name);
  }
  return bk16 ? launch_tc<128, 16>(al, ep, W_hi, W_lo, ldw, M, N, K, s, name) : launch_tc<128, 32>(al, ep, W_hi, W_lo, ldw, M, N, K, s, name);
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

int tc_gemm_try(const cvar_gemm_args* a, cudaStream_t s) {
  if (g_gemm_engine == 0) return 0;
  if (a->batch != 1 || a->w_is_kn || a->M < 64 || a->N < 64 || a->K % 32 != 0 || a->N % 4 != 0) return 0;
  if (a->lda % 4 != 0 || a->ldw % 4 != 0 || !aligned16(a->A) || !aligned16(a->W_hi) || !aligned16(a->W_lo)) return 0;
  DenseALoader al{a->A, a->lda, a->strideA, a->M, a->K, 1};
  al.A_lo = a->A_lo;
  DenseEpilogue ep{a->out, a->ldo, a->strideO, a->bias, a->epilogue, a->alpha, a->gamma, a->gamma_row_stride,
                   a->rows_per_sample, a->resid, a->ldr, a->strideR, a->out_lo};
  int rc = dispatch_tc(al, ep, a->W_hi, a->W_lo, a->ldw, (long long)a->M, a->N, a->K, s, "cvar_gemm[tc]");
  return rc ? rc : 1;
}

int tc_qkv_try(const float* A, const float* A_lo, const float* Wqkv_hi, const float* Wqkv_lo, const QkvEpilogue& ep, int M,
               int C, cudaStream_t s) {
  if (g_gemm_engine == 0) return 0;
  if (M < 64 || C % 32 != 0 || !aligned16(A) || !aligned16(Wqkv_hi) || !aligned16(Wqkv_lo)) return 0;
  DenseALoader al{A, C, 0, M, C, 1};
  al.A_lo = A_lo;
  int rc = dispatch_tc(al, ep, Wqkv_hi, Wqkv_lo, C, (long long)M, 3 * C, C, s, "cvar_qkv_project[tc]");
  return rc ? rc : 1;
}

int tc_split(const float* w, float* hi, float* lo, long long n, cudaStream_t s) {
  long long n4 = n / 4;
  split_tf32_kernel<<<cdiv(n4, 256), 256, 0, s>>>(w, hi, lo, n4);
  CVAR_CHECK_LAUNCH("cvar_split_tf32");
  return 0;
}

int tc_conv_try(const cvar_conv_args* a, cudaStream_t s) {
  if (g_gemm_engine == 0) return 0;
  if (a->Cin % 32 != 0 || a->Cout % 16 != 0 || a->Cout < 32 || a->downsample2x) return 0;
  const int up = a->upsample2x ? 1 : 0;
  const int Hout = a->Hin << up, Wout = a->Win << up;
  const long long M = (long long)a->B * Hout * Wout;
  if (M < 64) return 0;
  const int K = a->ks * a->ks * a->Cin;
  ConvALoader al;
  al.x = a->x, al.in_a = a->in_a, al.in_b = a->in_b, al.in_silu = a->in_silu;
  al.Hin = a->Hin, al.Win = a->Win, al.Cin = a->Cin, al.ks = a->ks, al.up = up;
  al.Hout = Hout, al.Wout = Wout, al.Mtot = M, al.K = K;
  al.Hv = Hout, al.Wv = Wout, al.stride = 1, al.pad = a->ks >> 1;
  ConvEpilogue ep{a->out, a->bias, a->resid, a->Cout, a->out_mode, Hout, Wout, a->out_rows_total, a->row_offset};
  ep.out_samples = a->out_samples;
  int rc = dispatch_tc(al, ep, a->w_hi, a->w_lo, K, M, a->Cout, K, s, "cvar_conv2d[tc]");
  return rc ? rc : 1;
}

}  // namespace tc

int tc_gemm_try(const cvar_gemm_args* a, cudaStream_t s) { return tc::tc_gemm_try(a, s); }
int tc_qkv_try(const float* A, const float* A_lo, const float* Wqkv_hi, const float* Wqkv_lo, const QkvEpilogue& ep, int M,
               int C, cudaStream_t s) {
  return tc::tc_qkv_try(A, A_lo, Wqkv_hi, Wqkv_lo, ep, M, C, s);
}
int tc_split(const float* w, float* hi, float* lo, long long n, cudaStream_t s) { return tc::tc_split(w, hi, lo, n, s); }
int tc_conv_try(const cvar_conv_args* a, cudaStream_t s) { return tc::tc_conv_try(a, s); }
}  // namespace cvar
