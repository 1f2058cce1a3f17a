// CFG mix + top-k + top-p + multinomial(1) for one scale (control_var.py:501-505, helpers.py:6-19).
// One CTA per (sample, token) row of V = 4096 logits.  The combined logits never go back to HBM: they live in
// shared memory between the CFG mix, the k-th-largest radix select, the ascending sort that top-p needs, and
// the final argmax(softmax(v) / q).
#include "common.cuh"

using namespace cvar;

namespace {
constexpr int V_FIXED = 4096;
constexpr int NT = 256;
constexpr int PER = V_FIXED / NT;   // 16

__device__ __forceinline__ unsigned orderable(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(unsigned u) {
  unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(b);
}

struct SampleSmem {
  unsigned long long keys[V_FIXED];   // (orderable value << 32) | index
  float vals[V_FIXED];
  unsigned hist[256];
  float redf[NT / 32];
  int redi[NT / 32];
  unsigned sel_prefix;
  unsigned sel_remaining;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) t += red[w];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) t = fmaxf(t, red[w]);
  return t;
}

struct MixCoef {
  float c[4];
};

__global__ void __launch_bounds__(NT) cfg_sample_kernel(const float* __restrict__ logits, const float* __restrict__ qn,
                                                        int64_t* __restrict__ idx_out, int B, int l, int groups,
                                                        MixCoef coef, int replicas, int top_k, int use_top_p, float thr,
                                                        const int64_t* __restrict__ forced_first,
                                                        const int64_t* __restrict__ forced_second, int forced_replicas,
                                                        float* __restrict__ masked_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SampleSmem& sm = *reinterpret_cast<SampleSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row = blockIdx.x;
  const int b = (int)(row / l), t = (int)(row % l);
  // teacher forcing (control_var.py:309-321): replicas whose token is given need no sample; when every replica of
  // this row is forced the whole distribution is skipped
  const int half = l >> 1;
  const int64_t* forced = t < half ? forced_first : forced_second;
  const long long forced_at = (long long)b * half + (t < half ? t : t - half);
  if (forced != nullptr && forced_replicas >= replicas && masked_out == nullptr) {
    if (tid < replicas) idx_out[(long long)tid * B * l + row] = forced[forced_at];
    return;
  }

  // sum_g coef[g] * logits[g*B + b]: every product rounded, summed left to right.  Two groups with (1 + t, -t) is
  // (1 + t) * logits[:B] - t * logits[B:] of control_var.py:502; four groups is the mix of control_var.py:295-298
  // (a - x == a + (-x) and (-c) * y == -(c * y) exactly in IEEE arithmetic).
  {
    const float* lg = logits + ((long long)b * l + t) * V_FIXED;
    const long long gstride = (long long)B * l * V_FIXED;
    // (unrolled over the four possible groups: the run-time loop cost ~45 instructions per element, 10 % of the kernel)
    const float* lg1 = lg + (groups > 1 ? gstride : 0);
    const float* lg2 = lg + (groups > 2 ? 2 * gstride : 0);
    const float* lg3 = lg + (groups > 3 ? 3 * gstride : 0);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int e = tid + i * NT;
      float v = __fmul_rn(coef.c[0], lg[e]);
      if (groups > 1) v = __fadd_rn(v, __fmul_rn(coef.c[1], lg1[e]));
      if (groups > 2) v = __fadd_rn(v, __fmul_rn(coef.c[2], lg2[e]));
      if (groups > 3) v = __fadd_rn(v, __fmul_rn(coef.c[3], lg3[e]));
      sm.vals[e] = v;
    }
  }
  __syncthreads();

  // ---- top-k: logits < (k-th largest) -> -inf (ties with the k-th kept)                    helpers.py:8-10
  if (top_k > 0 && top_k < V_FIXED) {
    if (tid == 0) {
      sm.sel_prefix = 0u;
      sm.sel_remaining = (unsigned)top_k;
    }
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      sm.hist[tid] = 0u;
      __syncthreads();
      const unsigned prefix = sm.sel_prefix;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        unsigned u = orderable(sm.vals[tid + i * NT]);
        bool match = (pass == 0) || ((u >> (shift + 8)) == (prefix >> (shift + 8)));
        if (match) atomicAdd(&sm.hist[(u >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (warp == 0) {
        // lane owns bins [255 - 8*lane - 7, 255 - 8*lane], scanned from the top bin down
        unsigned cnt[8], local = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          cnt[j] = sm.hist[255 - 8 * lane - j];
          local += cnt[j];
        }
        unsigned incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += n;
        }
        unsigned excl = incl - local;
        const unsigned rem = sm.sel_remaining;
        __syncwarp();     // every lane has read sel_remaining before the owning lane rewrites it below
        if (excl < rem && rem <= incl) {
          unsigned c = excl;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (c < rem && rem <= c + cnt[j]) {
              sm.sel_prefix = prefix | ((unsigned)(255 - 8 * lane - j) << shift);
              sm.sel_remaining = rem - c;
            }
            c += cnt[j];
          }
        }
      }
      __syncthreads();
    }
    const unsigned kth = sm.sel_prefix;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int e = tid + i * NT;
      if (orderable(sm.vals[e]) < kth) sm.vals[e] = -INFINITY;
    }
    __syncthreads();
  }

  float vmax;
  // ---- top-p: remove the ascending prefix whose softmax mass is <= 1 - top_p               helpers.py:11-15
  if (use_top_p) {
    // Entries already removed by top-k are -inf: in the ascending order they come first, contribute exp(-inf) = 0 to every
    // sum below and are written back as -inf whatever their order.  So instead of sorting all 4096 keys, the survivors
    // are compacted into the LAST slots and only the last kSortTail slots are sorted when they all fit there (top_k = 900
    // leaves ~900): the sorted tail - hence every running sum, taken per thread over the same 16 consecutive slots - is
    // bit-identical to that of the full sort, at a fifth of the compare-exchange work.
    constexpr int kSortTail = 1024;
    int nsurv_local = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) nsurv_local += (sm.vals[tid + i * NT] != -INFINITY) ? 1 : 0;
    // block-wide exclusive scan of the per-thread survivor counts (thread order, then element order i)
    int incl_c = nsurv_local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, incl_c, o);
      if (lane >= o) incl_c += n;
    }
    __syncthreads();
    if (lane == 31) sm.redi[warp] = incl_c;
    __syncthreads();
    int woff_c = 0, nsurv = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
      if (w < warp) woff_c += sm.redi[w];
      nsurv += sm.redi[w];
    }
    int pos_s = woff_c + incl_c - nsurv_local;            // rank of this thread's first survivor among all survivors
    int pos_d = (tid * PER) - pos_s;                      // rank of its first removed entry (elements before it: tid*PER)
    // NB: ranks are taken in (thread, i) order, not in index order - any order will do (see above)
    __syncthreads();                                      // sm.redi is reused by the final arg-max
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int e = tid + i * NT;
      const float v = sm.vals[e];
      const unsigned long long key = ((unsigned long long)orderable(v) << 32) | (unsigned)e;
      if (v != -INFINITY)
        sm.keys[V_FIXED - nsurv + pos_s++] = key;
      else
        sm.keys[pos_d++] = key;
    }
    __syncthreads();
    const bool tail_only = nsurv <= kSortTail;            // uniform
    const int base = tail_only ? V_FIXED - kSortTail : 0;
    const int nsort = tail_only ? kSortTail : V_FIXED;
    for (int k = 2; k <= nsort; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int p = tid; p < nsort / 2; p += NT) {
          int i = 2 * p - (p & (j - 1));
          int ixj = i + j;
          bool up = (i & k) == 0;
          unsigned long long a = sm.keys[base + i], c = sm.keys[base + ixj];
          if ((a > c) == up) {
            sm.keys[base + i] = c;
            sm.keys[base + ixj] = a;
          }
        }
        __syncthreads();
      }
    }
    vmax = from_orderable((unsigned)(sm.keys[V_FIXED - 1] >> 32));
    // softmax over the sorted values, then the running sum in ascending order
    float ev[PER];
    float loc = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      float v = from_orderable((unsigned)(sm.keys[tid * PER + i] >> 32));
      ev[i] = v == -INFINITY ? 0.f : expf(v - vmax);       // expf(-inf) is exactly 0: same value, no range reduction
      loc += ev[i];
    }
    const float total = block_sum(loc, sm.redf);
    float run = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      // 0 / total is exactly 0, but the IEEE division sends a zero numerator down its slow path (FCHK): three quarters of
      // the entries are removed ones, and that path was 19 % of the kernel's instructions (profiles/r02_sample.md)
      ev[i] = ev[i] == 0.f ? 0.f : ev[i] / total;
      run += ev[i];
      ev[i] = run;
    }
    // exclusive scan of the per-thread totals
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    __syncthreads();
    if (lane == 31) sm.redf[warp] = incl;
    __syncthreads();
    float woff = 0.f;
    for (int w = 0; w < warp; ++w) woff += sm.redf[w];
    const float off = woff + (incl - run);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int s = tid * PER + i;
      float cum = off + ev[i];
      if (cum <= thr && s != V_FIXED - 1) sm.vals[(unsigned)(sm.keys[s] & 0xffffffffu)] = -INFINITY;
    }
    __syncthreads();
  } else {
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < PER; ++i) m = fmaxf(m, sm.vals[tid + i * NT]);
    vmax = block_max(m, sm.redf);
  }

  // more_smooth (control_var.py:513-515): the caller wants the logits as sample_with_top_k_top_p_ leaves them IN PLACE
  // (helpers.py:10, 15: removed entries -inf) - they feed the Gumbel-softmax mixture of cvar_gumbel_embed
  if (masked_out != nullptr) {
#pragma unroll
    for (int i = 0; i < PER; ++i) masked_out[row * V_FIXED + tid + i * NT] = sm.vals[tid + i * NT];
  }

  // ---- multinomial(softmax(v), 1) = argmax(softmax(v) / q)                                 helpers.py:19
  float ev[PER];
  float loc = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const float v = sm.vals[tid + i * NT];
    ev[i] = v == -INFINITY ? 0.f : expf(v - vmax);         // removed entries: expf(-inf) is exactly 0
    loc += ev[i];
  }
  const float total = block_sum(loc, sm.redf);
  for (int rep = 0; rep < replicas; ++rep) {      // logits_BlV.repeat(replicas, 1, 1): one distribution, independent draws
    const long long orow = (long long)rep * B * l + row;
    if (forced != nullptr && rep < forced_replicas) {
      if (tid == 0) idx_out[orow] = forced[forced_at];
      continue;
    }
    const float* qr = qn + orow * V_FIXED;
    float best = -INFINITY;
    int besti = V_FIXED;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int e = tid + i * NT;
      // A removed entry has probability exactly 0, so (0 / total) / q is 0 for every q > 0: it can never beat the largest
      // surviving entry ((1 / total) / q > 0).  Its noise value is not loaded and its two divisions - which take the IEEE
      // slow path for a zero numerator - are not evaluated.
      float r = ev[i] == 0.f ? 0.f : (ev[i] / total) / qr[e];
      if (r > best) {        // ascending e: the first maximum wins, like torch.argmax
        best = r;
        besti = e;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) {
        best = ob;
        besti = oi;
      }
    }
    __syncthreads();
    if (lane == 0) {
      sm.redf[warp] = best;
      sm.redi[warp] = besti;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < NT / 32; ++w) {
        if (sm.redf[w] > best || (sm.redf[w] == best && sm.redi[w] < besti)) {
          best = sm.redf[w];
          besti = sm.redi[w];
        }
      }
      idx_out[orow] = (int64_t)(besti < V_FIXED ? besti : 0);
    }
  }
}
}  // namespace

static int launch_cfg_sample(const float* logits, const float* q_noise, int64_t* idx_out, int B, int l, int groups,
                             MixCoef coef, int replicas, int top_k, double top_p, const int64_t* forced_first,
                             const int64_t* forced_second, int forced_replicas, cudaStream_t stream, const char* name,
                             float* masked_out = nullptr) {
  // python evaluates (1 - top_p) in double, then the scalar meets an fp32 tensor as an fp32 value
  const float thr = (float)(1.0 - top_p);
  cudaError_t e = cudaFuncSetAttribute(cfg_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(SampleSmem));
  CVAR_REQUIRE(e == cudaSuccess, "%s: cannot raise shared memory: %s", name, cudaGetErrorString(e));
  cfg_sample_kernel<<<(unsigned)((long long)B * l), NT, sizeof(SampleSmem), stream>>>(
      logits, q_noise, idx_out, B, l, groups, coef, replicas, top_k, top_p > 0.0 ? 1 : 0, thr, forced_first,
      forced_second, forced_replicas, masked_out);
  CVAR_CHECK_LAUNCH(name);
  return 0;
}

extern "C" int cvar_cfg_sample(const float* logits, const float* q_noise, int64_t* idx_out, int B, int l, int V,
                               double t, int top_k, double top_p, void* stream) {
  CVAR_REQUIRE(V == V_FIXED, "cvar_cfg_sample: V must be %d (got %d)", V_FIXED, V);
  CVAR_REQUIRE(B > 0 && l > 0 && top_k >= 0 && top_k <= V, "cvar_cfg_sample: bad arguments");
  // python evaluates (1 + t) in double, then the scalar meets an fp32 tensor as an fp32 value
  MixCoef coef{{(float)(1.0 + t), -(float)t, 0.f, 0.f}};
  return launch_cfg_sample(logits, q_noise, idx_out, B, l, 2, coef, 1, top_k, top_p, nullptr, nullptr, 0,
                           (cudaStream_t)stream, "cvar_cfg_sample");
}

extern "C" int cvar_cfg_sample_multi(const float* logits, const float* q_noise, int64_t* idx_out, int B, int l, int V,
                                     int groups, const float* host_coef, int replicas, int top_k, double top_p,
                                     const int64_t* forced_first, const int64_t* forced_second, int forced_replicas,
                                     void* stream) {
  CVAR_REQUIRE(V == V_FIXED, "cvar_cfg_sample_multi: V must be %d (got %d)", V_FIXED, V);
  CVAR_REQUIRE(B > 0 && l > 0 && l % 2 == 0 && top_k >= 0 && top_k <= V, "cvar_cfg_sample_multi: bad arguments");
  CVAR_REQUIRE(groups >= 1 && groups <= 4 && host_coef != nullptr, "cvar_cfg_sample_multi: 1..4 logit groups");
  CVAR_REQUIRE(replicas >= 1 && replicas <= NT && forced_replicas >= 0 && forced_replicas <= replicas,
               "cvar_cfg_sample_multi: bad replicas / forced_replicas");
  MixCoef coef{{0.f, 0.f, 0.f, 0.f}};
  for (int g = 0; g < groups; ++g) coef.c[g] = host_coef[g];
  return launch_cfg_sample(logits, q_noise, idx_out, B, l, groups, coef, replicas, top_k, top_p, forced_first,
                           forced_second, forced_replicas, (cudaStream_t)stream, "cvar_cfg_sample_multi");
}

// cvar_cfg_sample_multi that also returns the mixed logits as sample_with_top_k_top_p_ leaves them (B*l, V; removed entries
// -inf): the input of the more_smooth path.
extern "C" int cvar_cfg_sample_masked(const float* logits, const float* q_noise, int64_t* idx_out, float* masked_out, int B,
                                      int l, int V, int groups, const float* host_coef, int replicas, int top_k, double top_p,
                                      const int64_t* forced_first, const int64_t* forced_second, int forced_replicas,
                                      void* stream) {
  CVAR_REQUIRE(V == V_FIXED, "cvar_cfg_sample_masked: V must be %d (got %d)", V_FIXED, V);
  CVAR_REQUIRE(masked_out != nullptr, "cvar_cfg_sample_masked: masked_out is null");
  CVAR_REQUIRE(B > 0 && l > 0 && l % 2 == 0 && top_k >= 0 && top_k <= V, "cvar_cfg_sample_masked: bad arguments");
  CVAR_REQUIRE(groups >= 1 && groups <= 4 && host_coef != nullptr, "cvar_cfg_sample_masked: 1..4 logit groups");
  CVAR_REQUIRE(replicas >= 1 && replicas <= NT && forced_replicas >= 0 && forced_replicas <= replicas,
               "cvar_cfg_sample_masked: bad replicas / forced_replicas");
  MixCoef coef{{0.f, 0.f, 0.f, 0.f}};
  for (int g = 0; g < groups; ++g) coef.c[g] = host_coef[g];
  return launch_cfg_sample(logits, q_noise, idx_out, B, l, groups, coef, replicas, top_k, top_p, forced_first,
                           forced_second, forced_replicas, (cudaStream_t)stream, "cvar_cfg_sample_masked", masked_out);
}

// ------------------------------------------------------------------------------------------------ more_smooth
// h = softmax((logits * mul + (-log e)) / tau) @ embedding      control_var.py:514-515, helpers.py:26-28 (hard = False)
// One CTA per output row; rows_in distinct logit rows are repeated (row % rows_in: logits_BlV.repeat(4, 1, 1) of
// conditional_infer_cfg, control_var.py:306), every output row has its own Exp(1) noise.  Rounding follows the reference's
// elementwise steps (mul, log, neg, add, div, then softmax = exp(x - max) / sum); the 4096-term products with the code
// vectors are summed in a different order than ATen's matmul (a float path: compared within tolerance).
namespace {
constexpr int kGumCv = 32;
__global__ void __launch_bounds__(NT) gumbel_embed_kernel(const float* __restrict__ masked, const float* __restrict__ e_noise,
                                                          const float* __restrict__ emb, float* __restrict__ h_out,
                                                          long long rows_in, float mul, float tau) {
  __shared__ float y_s[V_FIXED];
  __shared__ float red[NT / 32];
  __shared__ float part[NT / 32][kGumCv];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row = blockIdx.x;
  const float* lg = masked + (row % rows_in) * V_FIXED;
  const float* en = e_noise + row * V_FIXED;
  float v[PER];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int e = tid + i * NT;
    const float g = -logf(en[e]);
    v[i] = __fdiv_rn(__fadd_rn(__fmul_rn(lg[e], mul), g), tau);
    m = fmaxf(m, v[i]);
  }
  const float vmax = block_max(m, red);
  float loc = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    v[i] = v[i] == -INFINITY ? 0.f : expf(v[i] - vmax);
    loc += v[i];
  }
  const float total = block_sum(loc, red);
#pragma unroll
  for (int i = 0; i < PER; ++i) y_s[tid + i * NT] = v[i] == 0.f ? 0.f : v[i] / total;
  __syncthreads();
  // warp w owns entries [w * 512, (w + 1) * 512), lane = channel: one 128-byte code vector per step
  float acc = 0.f;
  const int e0 = warp * (V_FIXED / (NT / 32));
  for (int e = e0; e < e0 + V_FIXED / (NT / 32); ++e) {
    const float y = y_s[e];
    if (y != 0.f) acc = fmaf(y, emb[(long long)e * kGumCv + lane], acc);
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += part[w][lane];
    h_out[row * kGumCv + lane] = t;
  }
}
}  // namespace

extern "C" int cvar_gumbel_embed(const float* masked_logits, const float* e_noise, const float* embedding, float* h_out,
                                 long long rows_in, long long rows_out, int V, int Cvae, double mul, double tau,
                                 void* stream) {
  CVAR_REQUIRE(V == V_FIXED && Cvae == kGumCv, "cvar_gumbel_embed: V must be %d and Cvae %d", V_FIXED, kGumCv);
  CVAR_REQUIRE(masked_logits && e_noise && embedding && h_out, "cvar_gumbel_embed: null pointer");
  CVAR_REQUIRE(rows_in > 0 && rows_out > 0 && rows_out % rows_in == 0 && rows_out < (1LL << 31) && tau > 0.0,
               "cvar_gumbel_embed: bad arguments");
  // python forms (1 + ratio) and gum_t in double; they meet the fp32 tensors as fp32 values
  gumbel_embed_kernel<<<(unsigned)rows_out, NT, 0, (cudaStream_t)stream>>>(masked_logits, e_noise, embedding, h_out, rows_in,
                                                                        (float)mul, (float)tau);
  CVAR_CHECK_LAUNCH("cvar_gumbel_embed");
  return 0;
}
