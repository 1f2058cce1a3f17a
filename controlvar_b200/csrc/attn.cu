// KV-cached attention for the next-scale sampler (basic_var.py:106-117): l new queries of one scale attend to the
// L = L_prev + l cached keys of all scales so far; no mask (the cache IS the block-causal prefix).
// fp32 flash-style kernel: one CTA = 64 queries of one (row, head); K/V streamed through shared memory in 64-key
// tiles, online softmax, accumulators in registers.
#include "common.cuh"

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

__global__ void __launch_bounds__(256) attn_kvcache_kernel(const float* __restrict__ q, const float* __restrict__ kc,
                                                           const float* __restrict__ vc, float* __restrict__ out,
                                                           int H, int l, int L, int T_max, float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, r = blockIdx.z;
  const float* qb = q + (((long long)r * H + h) * l) * D;
  const float* kb = kc + (((long long)r * H + h) * T_max) * D;
  const float* vb = vc + (((long long)r * H + h) * T_max) * D;

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
      if (k0 + j < L) kv = ld4(kb + (long long)(k0 + j) * D + dq * 4);
      sm.Kt[dq * 4 + 0][j] = kv.x;
      sm.Kt[dq * 4 + 1][j] = kv.y;
      sm.Kt[dq * 4 + 2][j] = kv.z;
      sm.Kt[dq * 4 + 3][j] = kv.w;
    }
    for (int it = 0; it < 4; ++it) {
      int item = it * 256 + tid;
      int j = item >> 4, dq = item & 15;          // lanes walk the head dimension: coalesced, natural layout
      float4 vv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + j < L) vv = ld4(vb + (long long)(k0 + j) * D + dq * 4);
      *reinterpret_cast<float4*>(&sm.V[j][dq * 4]) = vv;
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
      float4 b = *reinterpret_cast<const float4*>(&sm.V[j][tx * 4]);
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
      st4(out + ((long long)r * l + t) * C + h * D + tx * 4, v);
    }
  }
}
}  // namespace

extern "C" int cvar_attn_kvcache(const float* q, const float* k_cache, const float* v_cache, float* out, int R, int H,
                                 int l, int L, int T_max, float scale, void* stream) {
  CVAR_REQUIRE(R > 0 && H > 0 && l > 0 && L >= l && L <= T_max, "cvar_attn_kvcache: bad shape l=%d L=%d T=%d", l, L,
               T_max);
  CVAR_REQUIRE(R <= 65535 && H <= 65535, "cvar_attn_kvcache: grid too large");
  cudaError_t e = cudaFuncSetAttribute(attn_kvcache_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(AttnSmem));
  CVAR_REQUIRE(e == cudaSuccess, "cvar_attn_kvcache: cannot raise shared memory: %s", cudaGetErrorString(e));
  dim3 grid(cdiv(l, BQ), H, R);
  attn_kvcache_kernel<<<grid, 256, sizeof(AttnSmem), (cudaStream_t)stream>>>(q, k_cache, v_cache, out, H, l, L, T_max,
                                                                            scale);
  CVAR_CHECK_LAUNCH("cvar_attn_kvcache");
  return 0;
}
