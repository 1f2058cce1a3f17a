// Multi-scale VQ step of the sampler (control_var.py:512-560 + quant.py:243-270) as ONE kernel per scale:
// embedding gather -> bicubic upsample to hw x hw -> Phi (0.5*h + 0.5*conv3x3) -> f_hat += -> area pool to the
// next scale -> word_embed (32 -> C) + level/position embedding, written for both CFG halves.
// In the reference this is ~12 tiny ATen kernels per stream per scale (launch bound); here everything between the
// token indices and the next scale's transformer input stays in shared memory.
#include "common.cuh"

using namespace cvar;

namespace {
constexpr int CV = 32;          // Cvae
constexpr int MAXHW = 16;

struct VqSmemLayout {
  // offsets in floats
  int w, h, t, hu, U, total;
};
__host__ __device__ inline VqSmemLayout vq_layout(int hw) {
  VqSmemLayout L;
  int o = 0;
  L.w = o, o += CV * 9 * CV;                    // phi weights as [ci][tap][co]
  L.h = o, o += CV * hw * hw;                   // gathered codes [c][pn*pn]; later the pooled map [c][pn'*pn']
  L.t = o, o += CV * hw * hw;                   // bicubic intermediate [c][y][X]; later the updated f_hat [c][hw*hw]
  L.hu = o, o += CV * (hw + 2) * (hw + 2);      // upsampled map with a zero halo
  L.U = o, o += hw * hw;                        // U[X][x]
  L.total = o;
  return L;
}

__global__ void __launch_bounds__(256) vq_step_kernel(const int64_t* __restrict__ idx, const float* __restrict__ emb,
                                                      const float* __restrict__ U, const float* __restrict__ phi_w,
                                                      const float* __restrict__ phi_b, const float* __restrict__ word_w,
                                                      const float* __restrict__ word_b,
                                                      const float* __restrict__ lvl_pos_next, float* __restrict__ f_hat,
                                                      float* __restrict__ f_rest, float* __restrict__ x_next, int B,
                                                      int x_rep, int pn, int pn_next, int hw, int C) {
  extern __shared__ __align__(16) float smf[];
  const VqSmemLayout L = vq_layout(hw);
  float* w_s = smf + L.w;
  float* h_s = smf + L.h;
  float* t_s = smf + L.t;
  float* hu_s = smf + L.hu;
  float* U_s = smf + L.U;
  const int tid = threadIdx.x;
  const int b = blockIdx.x, s = blockIdx.y, S = gridDim.y;   // S streams stacked along H: (control, image) or one map
  const int npix = pn * pn, HW = hw * hw, hp = hw + 2;

  for (int i = tid; i < CV * 9 * CV; i += 256) {
    int co = i % CV, r = i / CV;
    int tap = r % 9, ci = r / 9;
    w_s[i] = phi_w[(co * CV + ci) * 9 + tap];
  }
  if (pn != hw)
    for (int i = tid; i < hw * pn; i += 256) U_s[i] = U[i];
  for (int i = tid; i < CV * hp * hp; i += 256) hu_s[i] = 0.f;
  const int64_t* ib = idx + (long long)b * (S * npix) + s * npix;
  for (int i = tid; i < CV * npix; i += 256) {
    int c = i % CV, p = i / CV;                  // consecutive lanes read one 128-byte codebook row
    h_s[c * npix + p] = emb[ib[p] * CV + c];
  }
  __syncthreads();

  if (pn != hw) {
    // bicubic, x direction then y direction (the order ATen's upsample_bicubic2d accumulates in)
    for (int i = tid; i < CV * pn * hw; i += 256) {
      int X = i % hw, r = i / hw;
      int y = r % pn, c = r / pn;
      const float* hr = h_s + c * npix + y * pn;
      const float* ur = U_s + X * pn;
      float a = 0.f;
      for (int x = 0; x < pn; ++x) a = fmaf(ur[x], hr[x], a);
      t_s[(c * pn + y) * hw + X] = a;
    }
    __syncthreads();
    for (int i = tid; i < CV * HW; i += 256) {
      int X = i % hw, r = i / hw;
      int Y = r % hw, c = r / hw;
      const float* ur = U_s + Y * pn;
      float a = 0.f;
      for (int y = 0; y < pn; ++y) a = fmaf(ur[y], t_s[(c * pn + y) * hw + X], a);
      hu_s[(c * hp + Y + 1) * hp + X + 1] = a;
    }
  } else {
    for (int i = tid; i < CV * HW; i += 256) {
      int X = i % hw, r = i / hw;
      int Y = r % hw, c = r / hw;
      hu_s[(c * hp + Y + 1) * hp + X + 1] = h_s[c * npix + Y * pn + X];
    }
  }
  __syncthreads();

  // Phi: 0.5*hu + 0.5*conv3x3(hu) ; f_hat += phi                                   quant.py:269-270, 255
  float* f_s = t_s;   // t_s is dead from here on
  for (int p = tid; p < HW; p += 256) {
    int Y = p / hw, X = p % hw;
    float acc[CV];
#pragma unroll
    for (int co = 0; co < CV; ++co) acc[co] = phi_b[co];
    for (int ci = 0; ci < CV; ++ci) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        float v = hu_s[(ci * hp + Y + tap / 3) * hp + X + tap % 3];
        const float4* wv = reinterpret_cast<const float4*>(w_s + (ci * 9 + tap) * CV);
#pragma unroll
        for (int c4 = 0; c4 < CV / 4; ++c4) {
          float4 w4 = wv[c4];
          acc[4 * c4 + 0] = fmaf(v, w4.x, acc[4 * c4 + 0]);
          acc[4 * c4 + 1] = fmaf(v, w4.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(v, w4.z, acc[4 * c4 + 2]);
          acc[4 * c4 + 3] = fmaf(v, w4.w, acc[4 * c4 + 3]);
        }
      }
    }
#pragma unroll
    for (int co = 0; co < CV; ++co) {
      float hv = hu_s[(co * hp + Y + 1) * hp + X + 1];
      float phi = __fadd_rn(__fmul_rn(hv, 0.5f), __fmul_rn(acc[co], 0.5f));
      long long g = (((long long)b * CV + co) * (S * hw) + s * hw + Y) * hw + X;
      float fn = __fadd_rn(f_hat[g], phi);
      f_hat[g] = fn;
      f_s[co * HW + p] = fn;
      if (f_rest != nullptr) f_rest[g] = __fsub_rn(f_rest[g], phi);        // f_rest.sub_(h)   quant.py:211
    }
  }
  if (pn_next <= 0) return;
  __syncthreads();

  // area (adaptive average) pooling to pn_next x pn_next                           quant.py:256
  const int nn = pn_next * pn_next;
  float* nxt_s = h_s;   // [c][nn]
  for (int i = tid; i < CV * nn; i += 256) {
    int j = i % nn, c = i / nn;
    int iy = j / pn_next, ix = j % pn_next;
    int y0 = (iy * hw) / pn_next, y1 = ((iy + 1) * hw + pn_next - 1) / pn_next;
    int x0 = (ix * hw) / pn_next, x1 = ((ix + 1) * hw + pn_next - 1) / pn_next;
    float sum = 0.f;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) sum += f_s[c * HW + y * hw + x];
    nxt_s[c * nn + j] = sum / (float)(y1 - y0) / (float)(x1 - x0);
  }
  __syncthreads();

  // word_embed + lvl_pos, written x_rep times (the two CFG halves of control_var.py:560; once at :345-347)
  for (int co = tid; co < C; co += 256) {
    float w[CV];
#pragma unroll
    for (int c4 = 0; c4 < CV / 4; ++c4) {
      float4 t4 = ld4(word_w + (long long)co * CV + c4 * 4);
      w[4 * c4 + 0] = t4.x, w[4 * c4 + 1] = t4.y, w[4 * c4 + 2] = t4.z, w[4 * c4 + 3] = t4.w;
    }
    const float bias = word_b[co];
    for (int j = 0; j < nn; ++j) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < CV; ++c) a = fmaf(w[c], nxt_s[c * nn + j], a);
      int tok = s * nn + j;
      float v = __fadd_rn(__fadd_rn(a, bias), lvl_pos_next[(long long)tok * C + co]);
      for (int rep = 0; rep < x_rep; ++rep)
        x_next[((long long)(rep * B + b) * (S * nn) + tok) * C + co] = v;
    }
  }
}

// F.interpolate(f_rest, size=(pn, pn), mode='area').permute(0, 2, 3, 1).reshape(-1, C)      quant.py:199
__global__ void area_pool_nc_kernel(const float* __restrict__ f, float* __restrict__ z, int hw, int pn, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % CV);
  long long r = i / CV;
  int j = (int)(r % (pn * pn));
  long long b = r / (pn * pn);
  int iy = j / pn, ix = j % pn;
  int y0 = (iy * hw) / pn, y1 = ((iy + 1) * hw + pn - 1) / pn;
  int x0 = (ix * hw) / pn, x1 = ((ix + 1) * hw + pn - 1) / pn;
  const float* fc = f + (b * CV + c) * hw * hw;
  float sum = 0.f;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) sum += fc[y * hw + x];
  z[i] = sum / (float)(y1 - y0) / (float)(x1 - x0);
}

// L2 nearest code (quant.py:203-206): one warp per latent vector, lanes stride the codebook.
__global__ void __launch_bounds__(256) vq_nearest_kernel(const float* __restrict__ z, const float* __restrict__ emb,
                                                         int64_t* __restrict__ idx_out, int N, int V) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long n = (long long)blockIdx.x * 8 + warp;
  if (n >= N) return;
  float zr[CV];
  float zz = 0.f;
#pragma unroll
  for (int c = 0; c < CV; ++c) {
    zr[c] = z[n * CV + c];
    zz = fmaf(zr[c], zr[c], zz);
  }
  float best = INFINITY;
  int besti = 0x7fffffff;
  for (int v = lane; v < V; v += 32) {
    const float4* er = reinterpret_cast<const float4*>(emb + (long long)v * CV);
    float dot = 0.f, ee = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < CV / 4; ++c4) {
      float4 e = er[c4];
      dot = fmaf(zr[4 * c4 + 0], e.x, dot), ee = fmaf(e.x, e.x, ee);
      dot = fmaf(zr[4 * c4 + 1], e.y, dot), ee = fmaf(e.y, e.y, ee);
      dot = fmaf(zr[4 * c4 + 2], e.z, dot), ee = fmaf(e.z, e.z, ee);
      dot = fmaf(zr[4 * c4 + 3], e.w, dot), ee = fmaf(e.w, e.w, ee);
    }
    float d = fmaf(-2.f, dot, __fadd_rn(zz, ee));
    if (d < best) {
      best = d;
      besti = v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob < best || (ob == best && oi < besti)) {
      best = ob;
      besti = oi;
    }
  }
  if (lane == 0) idx_out[n] = besti;
}
}  // namespace

extern "C" int cvar_vq_step_ex(const int64_t* idx, const float* embedding, const float* U, const float* phi_w,
                               const float* phi_b, const float* word_w, const float* word_b, const float* lvl_pos_next,
                               float* f_hat, float* f_rest, float* x_next, int B, int streams, int x_replicas, int pn,
                               int pn_next, int hw, int Cvae, int C, void* stream) {
  CVAR_REQUIRE(Cvae == CV, "cvar_vq_step: Cvae must be %d", CV);
  CVAR_REQUIRE(B > 0 && (streams == 1 || streams == 2) && x_replicas >= 1, "cvar_vq_step: bad B / streams / x_replicas");
  CVAR_REQUIRE(hw >= 1 && hw <= MAXHW && pn >= 1 && pn <= hw && pn_next <= hw, "cvar_vq_step: bad sizes pn=%d hw=%d", pn,
               hw);
  CVAR_REQUIRE(pn == hw || U != nullptr, "cvar_vq_step: interpolation matrix missing");
  CVAR_REQUIRE(pn_next <= 0 || (x_next && word_w && word_b && lvl_pos_next), "cvar_vq_step: next-scale buffers missing");
  size_t smem = (size_t)vq_layout(hw).total * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(vq_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  CVAR_REQUIRE(e == cudaSuccess, "cvar_vq_step: cannot raise shared memory: %s", cudaGetErrorString(e));
  vq_step_kernel<<<dim3(B, streams), 256, smem, (cudaStream_t)stream>>>(idx, embedding, U, phi_w, phi_b, word_w, word_b,
                                                                        lvl_pos_next, f_hat, f_rest, x_next, B,
                                                                        x_replicas, pn, pn_next, hw, C);
  CVAR_CHECK_LAUNCH("cvar_vq_step");
  return 0;
}

extern "C" int cvar_vq_step(const int64_t* idx, const float* embedding, const float* U, const float* phi_w,
                            const float* phi_b, const float* word_w, const float* word_b, const float* lvl_pos_next,
                            float* f_hat, float* x_next, int B, int pn, int pn_next, int hw, int Cvae, int C,
                            void* stream) {
  return cvar_vq_step_ex(idx, embedding, U, phi_w, phi_b, word_w, word_b, lvl_pos_next, f_hat, nullptr, x_next, B, 2, 2,
                         pn, pn_next, hw, Cvae, C, stream);
}

extern "C" int cvar_area_pool_nc(const float* f_nchw, float* z_NC, int B, int Cvae, int hw, int pn, void* stream) {
  CVAR_REQUIRE(Cvae == CV && B > 0 && pn >= 1 && pn <= hw, "cvar_area_pool_nc: bad shape");
  long long total = (long long)B * pn * pn * CV;
  area_pool_nc_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(f_nchw, z_NC, hw, pn, total);
  CVAR_CHECK_LAUNCH("cvar_area_pool_nc");
  return 0;
}

extern "C" int cvar_vq_nearest(const float* z_NC, const float* embedding, int64_t* idx_out, int N, int Cvae, int V,
                               void* stream) {
  CVAR_REQUIRE(Cvae == CV && N > 0 && V > 0, "cvar_vq_nearest: bad shape");
  vq_nearest_kernel<<<cdiv(N, 8), 256, 0, (cudaStream_t)stream>>>(z_NC, embedding, idx_out, N, V);
  CVAR_CHECK_LAUNCH("cvar_vq_nearest");
  return 0;
}
