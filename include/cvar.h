/*
 * cvar.h - C ABI of libcvar_sm100.so: the sm_100a kernels behind ControlVAR's next-scale sampling hot path.
 *
 * The reference (lxa9867/ControlVAR) has no plugin / FFI layer: its boundary is the Python method surface
 *   ControlVAR.autoregressive_infer_cfg      models/control_var.py:356-565
 *   VQVAE.fhat_to_img                        models/vqvae.py:88-89
 * and, for pixel-level control,
 *   ControlVAR.conditional_infer_cfg         models/control_var.py:223-354
 *   VQVAE.img_to_idxBl                       models/vqvae.py:73-75, models/quant.py:184-215
 * (SURVEY.md section 8b).  The host mirror in controlvar_b200/ keeps those signatures and calls the entry points
 * below through ctypes.  Each entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - all launches are asynchronous on the caller's stream (a cudaStream_t passed as void*);
 *   - nothing is allocated or freed inside the library; scratch memory is caller-owned;
 *   - return value 0 = ok, negative = error; cvar_last_error() returns a thread-local message;
 *   - tensors are dense fp32 row-major unless a stride is given; token / class indices are int64.
 */
#ifndef CVAR_H_
#define CVAR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVAR_ABI_VERSION 1
#if defined(__GNUC__)
#define CVAR_API __attribute__((visibility("default")))
#else
#define CVAR_API
#endif

CVAR_API int cvar_abi_version(void);
CVAR_API const char* cvar_last_error(void);
/* Number of kernel launches issued by this library in the calling process since load (bench.py: gpu_launches). */
CVAR_API long long cvar_launch_count(void);
/* Which GEMM engine serves cvar_gemm / cvar_qkv_project / cvar_conv2d when the shape and operands allow:
 *   0 = SIMT fp32 FFMA;  1 = tcgen05 3xTF32, one CTA per tile (fp32-class);  2 = reserved (bf16, not implemented);
 *   3 = as 1, and dense layers whose activation operand is supplied pre-split (A_lo != NULL) run on the 2-CTA
 *       (cta_group::2) all-TMA kernel.  Anything an engine does not take falls through to the next lower one.
 *   4 = as 3 for TF32 operands; tells the HOST to hand the dense layers FP16-pair operands (A16_* / W16_*, see
 *       cvar_split_f16), which run on the same 2-CTA kernel with kind::f16 MMAs at twice the TF32 rate.  The library
 *       itself picks the FP16 path whenever a call carries A16_hi, under any engine other than 0.
 * Default: 4 (environment variable CVAR_GEMM_ENGINE = 0 | 1 | 3 | 4 overrides it at load).  Returns the previous value. */
CVAR_API int cvar_set_gemm_engine(int engine);
CVAR_API int cvar_get_gemm_engine(void);
/* Epilogue of the 2-CTA tcgen05 kernels (dense layers, QKV, convolutions).  1 (default): every epilogue warp first pulls
 * its slice of the accumulators out of tensor memory into registers and releases tensor memory, so that its global-memory
 * work overlaps the next tile's MMAs; 0: tensor memory is held until the tile is stored (the round-1 behaviour, kept for
 * A/B timing).  Results are bit-identical.  Environment variable CVAR_EPI_OVERLAP = 0 | 1 sets the initial value.
 * Returns the previous value. */
CVAR_API int cvar_set_epilogue_overlap(int on);
/* Adds n to the launch counter: a host that replays a captured CUDA graph of this library's kernels accounts for the
 * replayed launches with it (the library counts only the launches it issues itself). */
CVAR_API int cvar_add_launch_count(long long n);
/* Fast mode - NOT a parity mode.  0 (default): three MMAs per product on FP16 pairs (fp32-class results, the mode every
 * parity claim refers to).  1: the 2-CTA GEMM / convolution kernels and the f16 attention kernel use the `hi` halves only,
 * one kind::f16 MMA per product with fp32 accumulation (SURVEY.md section 7.2 "fast mode"): half-precision operand
 * rounding (2^-11 relative), tokens diverge from the reference.  Reported separately by bench.py --fast.  Returns the
 * previous value.  Environment variable CVAR_FAST_MODE = 1 sets it at load. */
CVAR_API int cvar_set_fast_mode(int on);
CVAR_API int cvar_get_fast_mode(void);
/* K-block of the tcgen05 engine: 32 (128-byte swizzle, default) or 16 (64-byte swizzle, deeper pipeline). Returns the
 * previous value. */
CVAR_API int cvar_set_tc_kblock(int bk);
/* Diagnostics: device buffer of 4*64*2 int64 that CTA 0 of the tcgen05 engine fills with clock64() stamps of its
 * pipeline hand-overs (NULL switches the trace off). */
CVAR_API int cvar_debug_set_trace(long long* dev_buf);
/* Same for the tensor-core attention kernels: 3*32*8 int64, CTA (0,0,0), per KV tile (NULL switches it off). */
CVAR_API int cvar_debug_set_attn_trace(long long* dev_buf);

/* ---- prologue: control_var.py:381-383, 399-409 -------------------------------------------------------------
 * lvl_pos[t,:] = lvl_embed[lvl_1L[t],:] + pos_1LC[t,:]                                   (control_var.py:383) */
CVAR_API int cvar_lvl_pos(const float* lvl_embed, const int64_t* lvl_1L, const float* pos_1LC, float* lvl_pos,
                 int T, int C, void* stream);
/* rows r < B use (label[r], cond_type[r]); rows r >= B use (num_classes, 4) - the CFG "unconditional" half.
 * cond_BD[r,:]   = class_emb[lab,:]                                  (control_var.py:381)
 * silu_cond[r,:] = SiLU(cond_BD[r,:])                                (input of every ada_lin, basic_var.py:198)
 * x0[r,0,:] = cond_embed[ct,:] + pos_start[0,:] + lvl_pos[0,:]       (control_var.py:402-409)
 * x0[r,1,:] = class_emb[lab,:] + pos_start[1,:] + lvl_pos[1,:] */
CVAR_API int cvar_prologue(const float* class_emb, const float* cond_embed, const float* pos_start, const float* lvl_pos,
                  const int64_t* label_B, const int64_t* cond_type_B, int B, int C, int num_classes,
                  float* cond_BD, float* silu_cond, float* x0, void* stream);

/* The same for explicit per-row ids: row r uses (label_R[r], cond_type_R[r]).  conditional_infer_cfg runs four guidance
 * replicas [class+type | type only | none | none] (control_var.py:252-269); the host concatenates the ids. */
CVAR_API int cvar_prologue_rows(const float* class_emb, const float* cond_embed, const float* pos_start,
                       const float* lvl_pos, const int64_t* label_R, const int64_t* cond_type_R, int R, int C,
                       float* cond_BD, float* silu_cond, float* x0, void* stream);

/* ---- AdaLN-modulated LayerNorm: basic_var.py:208-209, control_var.py:699-701 ---------------------------------
 * y[m,:] = LayerNorm(x[m,:], eps, no affine) * (scale[r,:] + 1) + shift[r,:],  r = m / rows_per_sample.
 * scale/shift are slices of an ada_lin output, so they carry a row stride (6C or 2C floats). */
CVAR_API int cvar_ln_modulate(const float* x, const float* scale, const float* shift, long long mod_row_stride,
                     float* y, float* y_lo, void* y16_hi, void* y16_lo, int M, int C, int rows_per_sample, float eps,
                     void* stream);
/* y_lo (optional): when not NULL the result is written as a TF32 split, y = hi, y_lo = lo (hi + lo == value exactly) -
 * the operand format the 2-CTA GEMM fetches by TMA.
 * y16_hi / y16_lo (optional, IEEE half [M,C]): the result as an FP16 pair (see cvar_split_f16); y may then be NULL. */

/* ---- dense layers: F.linear call sites of basic_var.py:51,92,119 and control_var.py:221 -----------------------
 * out = epilogue(A[M,K] @ W[N,K]^T + bias[N]).  A, W, out row-major with leading dimensions lda/ldw/ldo.
 * batch > 1 runs independent problems separated by the given strides (decoder attention).
 * w_is_kn != 0 means W is stored [K,N] row-major (used for P @ V in vae_modules.py:89). */
enum {
  CVAR_EPI_BIAS = 0,            /* out = acc*alpha + bias                                                  */
  CVAR_EPI_BIAS_GELU = 1,       /* out = GELU_tanh(acc + bias)               (FFN.fc1 + act, basic_var.py:51) */
  CVAR_EPI_BIAS_GAMMA_RESID = 2,/* out[m,n] += gamma[m/rows_per_sample, n] * (acc + bias[n])  (basic_var.py:208-209) */
  CVAR_EPI_BIAS_RESID = 3       /* out[m,n] = resid[m,n] + (acc + bias[n])                    (vae_modules.py:60,92) */
};
typedef struct {
  const float* A; long long lda; long long strideA;
  /* optional: A is the TF32 'hi' part of the activation and A_lo the remainder (same shape / lda), as written by a
   * producer called with its *_lo output (cvar_ln_modulate, cvar_attn_kvcache, a cvar_gemm with out_lo) */
  const float* A_lo;
  const float* W; long long ldw; long long strideW; int w_is_kn;
  /* optional TF32 split of W made once by cvar_split_tf32 (same shape / ldw): the tcgen05 engine takes the problem only
   * when both are given; W itself always stays valid for the SIMT engine */
  const float* W_hi; const float* W_lo;
  const float* bias;                      /* [N] or NULL */
  float* out; long long ldo; long long strideO;
  int M, N, K, batch;
  int epilogue; float alpha;
  const float* gamma; long long gamma_row_stride; int rows_per_sample;   /* CVAR_EPI_BIAS_GAMMA_RESID */
  const float* resid; long long ldr; long long strideR;                  /* CVAR_EPI_BIAS_RESID */
  float* out_lo;   /* optional, CVAR_EPI_BIAS / CVAR_EPI_BIAS_GELU only: write the result split (out = hi, out_lo = lo) */
  /* FP16-pair operands (engine 4, "f16x3"): when A16_hi is set the product runs as three kind::f16 tcgen05 MMAs on the
   * pairs (A16_hi, A16_lo) [M, lda halves] and (W16_hi, W16_lo) [N, ldw halves] made by cvar_split_f16 or by a producer's
   * *16 outputs; A / A_lo / W / W_hi / W_lo are ignored.  Needs K % 64 == 0, batch == 1, W as [N,K]; any M, N.
   * out16_hi / out16_lo (optional, BIAS / BIAS_GELU): write the result as an FP16 pair [M, ldo halves]; out may be NULL. */
  const void* A16_hi; const void* A16_lo;
  const void* W16_hi; const void* W16_lo;
  void* out16_hi; void* out16_lo;
} cvar_gemm_args;
CVAR_API int cvar_gemm(const cvar_gemm_args* args, void* stream);

/* ---- QKV projection fused with the KV-cache append: basic_var.py:92-108 ------------------------------------
 * qkv = A[M,C] @ Wqkv[3C,C]^T + [q_bias, k_bias, v_bias];  M = R*l, row m = r*l + t.
 * q  -> q_out[r, h, t, :]                         (R, H, l, 64)
 * k  -> k_hi / k_lo[r, h, L_prev + t, :]          (R, H, T_max, 64)   TF32 split, k_hi + k_lo == k exactly
 * v  -> vt_hi / vt_lo[r, h, :, L_prev + t]        (R, H, 64, T_max)   same split, stored TRANSPOSED (keys contiguous)
 * This is the in-place replacement of the reference's torch.cat cache growth; the split / transposed layout is the
 * operand format of the tensor-core attention kernel (K tiles and V^T tiles are fetched by TMA as they are).
 * T_max must be a multiple of 4.  The caller zero-initialises the cache once (stale tail keys are masked, but must be
 * finite).  cos_attn != 0 (depth 30, basic_var.py:99-104): q = normalize(q) * exp(min(scale_mul[h], ln 100)),
 * k = normalize(k).  A16_* / W16_* (optional): FP16-pair operands as in cvar_gemm_args; the cache format is unchanged. */
CVAR_API int cvar_qkv_project(const float* A, const float* A_lo, const void* A16_hi, const void* A16_lo,
                     const float* Wqkv, const float* Wqkv_hi, const float* Wqkv_lo, const void* W16_hi, const void* W16_lo,
                     const float* q_bias, const float* k_bias, const float* v_bias,
                     float* q_out, float* k_hi, float* k_lo, float* vt_hi, float* vt_lo,
                     int R, int l, int L_prev, int T_max, int H, int cos_attn, const float* scale_mul_H,
                     void* stream);

/* ---- KV-cached attention: F.scaled_dot_product_attention at basic_var.py:117 --------------------------------
 * out[r, t, h*64:(h+1)*64] = softmax(q[r,h,t,:] . K[r,h,0:L,:]^T * scale) @ V[r,h,0:L,:]
 * q (R,H,l,64); cache arrays as written by cvar_qkv_project; out (R,l,H*64) - the layout proj consumes.  No mask: the
 * cache holds exactly the keys of scales <= current, which is the block-causal pattern of control_var.py:168.
 * engine: -1 = library default (tensor cores when l >= 64), 0 = SIMT fp32, 1 = tcgen05 3xTF32. */
CVAR_API int cvar_attn_kvcache(const float* q, const float* k_hi, const float* k_lo, const float* vt_hi, const float* vt_lo,
                      float* out, float* out_lo, void* out16_hi, void* out16_lo, int R, int H, int l, int L, int T_max,
                      float scale, int engine, void* stream);
/* out_lo (optional): write the result as a TF32 split (out = hi, out_lo = lo) for the 2-CTA proj GEMM.
 * out16_hi / out16_lo (optional, IEEE half (R,l,H*64)): the result as an FP16 pair; out may then be NULL. */

/* ---- the same two steps on FP16-pair operands end to end (engine 4) ---------------------------------------------
 * cvar_qkv_project16: A16 / W16 pairs in (as cvar_gemm_args), and q, K, V^T written as FP16 pairs with the index layout
 * of cvar_qkv_project: q16 (R,H,l,64), k16 (R,H,T_max,64), vt16 (R,H,64,T_max).  V^T is a standard pair (cvar_split_f16:
 * x ~= hi + lo * 2^-11).  q and K are "qk pairs": 16 x = hi + lo with hi = half_rn(16 x), lo = half_rn(16 x - hi) - the
 * residual is NOT scaled, so q.k = (hi hi + hi lo + lo hi) / 256 accumulates in a single tensor-core accumulator; the
 * factor 16 keeps the residual of an O(1) value a normal fp16 number (|x| < 4094).
 * Half the bytes of the TF32 split, and exactly the tiles the f16 attention kernel fetches by TMA.  T_max % 8 == 0; the
 * caller zero-initialises the cache once.  Needs a tensor-core engine (cvar_set_gemm_engine != 0). */
CVAR_API int cvar_qkv_project16(const void* A16_hi, const void* A16_lo, const void* W16_hi, const void* W16_lo,
                       const float* q_bias, const float* k_bias, const float* v_bias,
                       void* q16_hi, void* q16_lo, void* k16_hi, void* k16_lo, void* vt16_hi, void* vt16_lo,
                       int R, int l, int L_prev, int T_max, int H, int cos_attn, const float* scale_mul_H, void* stream);
/* cvar_attn_kvcache16: softmax(q K^T * scale) V on those pairs; out (R,l,H*64) fp32 and / or out16 pair (either may be
 * NULL, not both).  engine: -1 = default (tensor cores unless the library engine is 0), 0 = SIMT fp32 on the same operands,
 * 1 = tcgen05 kind::f16: three MMAs per product, CTA = 128 queries, two CTAs per SM (256 TMEM columns each: two S buffers
 * + O main / cross), Q / K / V^T tiles by TMA, P written back to TMEM over the S cells it came from. */
CVAR_API int cvar_attn_kvcache16(const void* q16_hi, const void* q16_lo, const void* k16_hi, const void* k16_lo,
                        const void* vt16_hi, const void* vt16_lo, float* out, void* out16_hi, void* out16_lo,
                        int R, int H, int l, int L, int T_max, float scale, int engine, void* stream);

/* cvar_attn_blockcausal16: the block-causal full-sequence pass of ControlVAR.forward (control_var.py:158-198, 622-636, the
 * released mask: a query of scale s attends to the keys of scales 0..s) in ONE launch on the tcgen05 kernel above.  The
 * mask is a step function of the query's scale, so no L x L bias tensor exists: the launch carries a table of query tiles
 * (first query, count, visible keys).  q16 (R, H, l_total, 64) and the caches hold the whole pyramid, as written by
 * cvar_qkv_project16 with L_prev = 0, l = l_total.  host_scale_lens: n_scales ints on the HOST, summing to l_total.
 * out / out16 as cvar_attn_kvcache16.  Not the 'indep' / 'separate_decoding' masks (no released configuration). */
CVAR_API int cvar_attn_blockcausal16(const void* q16_hi, const void* q16_lo, const void* k16_hi, const void* k16_lo,
                            const void* vt16_hi, const void* vt16_lo, float* out, void* out16_hi, void* out16_lo,
                            int R, int H, int l_total, int T_max, float scale, int n_scales, const int* host_scale_lens,
                            void* stream);

/* ---- CFG + top-k/top-p + multinomial(1): control_var.py:501-505, helpers.py:6-19 -----------------------------
 * logits (2B, l, V): rows [0,B) conditional, [B,2B) unconditional.  v = (1+t)*lc - t*lu; top-k keeps v >= k-th
 * largest (ties kept); top-p removes the ascending-sorted prefix whose softmax mass is <= 1-top_p (the largest is
 * always kept); idx = argmax(softmax(v) / q_noise) - the ATen multinomial(num_samples=1) rule.  q_noise (B*l, V) is
 * Exp(1) noise supplied by the caller's generator.  V must be 4096.  idx_out (B, l) int64.  t and top_p are doubles
 * because the reference forms (1+t) and (1-top_p) in Python double precision before they meet the fp32 tensors. */
CVAR_API int cvar_cfg_sample(const float* logits, const float* q_noise, int64_t* idx_out,
                    int B, int l, int V, double t, int top_k, double top_p, void* stream);

/* Generalisation used by conditional_infer_cfg (control_var.py:288-321): logits (groups*B, l, V);
 *   v = coef[0]*L_0 + coef[1]*L_1 + ... (each product rounded, summed left to right - the evaluation order of :295-298;
 *   a subtracted term is passed with a negated coefficient, which is bit-identical in IEEE arithmetic);
 * the masked distribution of row (b, t) is then sampled `replicas` times with independent noise rows
 * (logits_BlV.repeat(repeat_num, 1, 1), :306-307): q_noise (replicas*B*l, V), idx_out (replicas*B, l).
 * Teacher forcing (:309-321): for replicas g < forced_replicas, tokens t < l/2 are overwritten with
 * forced_first[b, t] and tokens t >= l/2 with forced_second[b, t - l/2] (either may be NULL: keep the sample).
 * coef (host pointer, `groups` floats, groups <= 4). */
CVAR_API int cvar_cfg_sample_multi(const float* logits, const float* q_noise, int64_t* idx_out, int B, int l, int V,
                          int groups, const float* host_coef, int replicas, int top_k, double top_p,
                          const int64_t* forced_first, const int64_t* forced_second, int forced_replicas, void* stream);

/* cvar_cfg_sample_multi that ALSO returns the mixed logits as sample_with_top_k_top_p_ leaves them in place
 * (helpers.py:10, 15: masked_out (B*l, V) fp32, removed entries -inf): the input of the more_smooth path
 * (control_var.py:513-515, 329-331).  Rows whose replicas are all forced are still evaluated. */
CVAR_API int cvar_cfg_sample_masked(const float* logits, const float* q_noise, int64_t* idx_out, float* masked_out, int B, int l,
                           int V, int groups, const float* host_coef, int replicas, int top_k, double top_p,
                           const int64_t* forced_first, const int64_t* forced_second, int forced_replicas, void* stream);

/* ---- more_smooth: gumbel_softmax_with_rng(logits.mul(1 + ratio), tau, hard=False) @ embedding ---------------------
 * control_var.py:514-515 (and :330-331), helpers.py:22-36.  masked_logits (rows_in, V) from cvar_cfg_sample_masked;
 * e_noise (rows_out, V): the Exp(1) tensor helpers.py:26 draws (gumbel = -log e); rows_out = k * rows_in: output row r
 * uses logits row r % rows_in (logits_BlV.repeat(4, 1, 1), control_var.py:306).  h_out (rows_out, Cvae) fp32:
 *   h = softmax((masked * mul + gumbel) / tau) @ embedding,   mul = 1 + ratio, tau = max(0.27 (1 - 0.95 ratio), 0.005). */
CVAR_API int cvar_gumbel_embed(const float* masked_logits, const float* e_noise, const float* embedding, float* h_out,
                      long long rows_in, long long rows_out, int V, int Cvae, double mul, double tau, void* stream);

/* ---- multi-scale VQ step: control_var.py:512-560 + quant.py:243-270 ------------------------------------------
 * For sample b and stream s in {0: control, 1: image}:
 *   h  = embedding[idx[b, s*pn*pn + i], :] as a (32, pn, pn) map                      (control_var.py:512-524)
 *   hu = bicubic(h -> hw x hw) = U h U^T   (identity at the last scale)               (quant.py:254)
 *   f_hat[b, :, s*hw:(s+1)*hw, :] += 0.5*hu + 0.5*(conv3x3(hu; phi_w, phi_b))         (quant.py:269-270, 255)
 *   nxt = area(f_hat_s -> pn_next x pn_next)  (adaptive average pooling bins)         (quant.py:256)
 *   x_next[b (and B+b), s*pn_next^2 + j, :] = word_embed(nxt[:, j]) + lvl_pos_next[s*pn_next^2 + j, :]   (control_var.py:555-560)
 * U (hw x pn) is the 1-D interpolation matrix of F.interpolate(mode='bicubic', align_corners=False); f_hat is (B, 32, 2*hw, hw) NCHW as in the reference.  pn_next == 0 marks the last scale (no x_next). */
CVAR_API int cvar_vq_step(const int64_t* idx, const float* embedding, const float* U, const float* phi_w, const float* phi_b,
                 const float* word_w, const float* word_b, const float* lvl_pos_next,
                 float* f_hat, float* x_next, int B, int pn, int pn_next, int hw, int Cvae, int C, void* stream);

/* General form.  streams: 2 = (control, image) halves stacked along H as above, 1 = a single (B, 32, hw, hw) map.
 * x_replicas: how many row groups of x_next receive the result (2 = the CFG halves of autoregressive_infer_cfg, which
 * repeats the map at control_var.py:560; 1 = conditional_infer_cfg, where every replica has its own f_hat).
 * f_rest (optional, same layout as f_hat): the residual of the encoder-side quantiser, f_rest -= phi (quant.py:211). */
CVAR_API int cvar_vq_step_ex(const int64_t* idx, const float* embedding, const float* U, const float* phi_w,
                    const float* phi_b, const float* word_w, const float* word_b, const float* lvl_pos_next,
                    float* f_hat, float* f_rest, float* x_next, int B, int streams, int x_replicas, int pn, int pn_next,
                    int hw, int Cvae, int C, void* stream);
/* z[(b*pn*pn + j), c] = area-pooled f[b, c, :, :] at bin j (F.interpolate(mode='area') = adaptive average pooling,
 * quant.py:199), written as the (N, 32) row matrix cvar_vq_nearest consumes; pn == hw is a plain NCHW -> NHWC copy. */
CVAR_API int cvar_area_pool_nc(const float* f_nchw, float* z_NC, int B, int Cvae, int hw, int pn, void* stream);

/* L2 nearest code: quant.py:203-206.  idx[n] = argmin_v (|z_n|^2 + |e_v|^2 - 2 z_n.e_v), first index on ties. */
CVAR_API int cvar_vq_nearest(const float* z_NC, const float* embedding, int64_t* idx_out, int N, int Cvae, int V, void* stream);

/* ---- VQVAE decoder: vae_modules.py:18-28,57-60,73-92,210-226; vqvae.py:88-89 ---------------------------------
 * Activations are NHWC inside the library. */
/* (B,C,H,W) -> (B,H,W,C) */
CVAR_API int cvar_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W,
                      long long in_batch_stride, void* stream);
/* The same with the channel dimension zero-padded to Cpad >= C (the 3-channel image entering Encoder.conv_in runs as a
 * 16-channel NHWC tensor; cvar_repack_conv_weight_pad pads the weight to match). */
CVAR_API int cvar_nchw_to_nhwc_pad(const float* in, float* out, int B, int C, int H, int W, int Cpad, void* stream);
/* GroupNorm statistics folded with the affine: a[n,c] = rstd[n,g]*gamma[c], b[n,c] = beta[c] - mean[n,g]*a[n,c],
 * so GroupNorm(x)[n,:,:,c] = x*a + b.  scratch: 2*B*groups*chunks doubles (chunks = cvar_gn_chunks(HW)). */
CVAR_API int cvar_gn_chunks(int HW);
CVAR_API int cvar_gn_stats(const float* x_nhwc, const float* gamma, const float* beta, float* a_out, float* b_out,
                  double* scratch, int B, int HW, int C, int groups, float eps, void* stream);
/* conv2d, stride 1, 'same' padding, ks in {1,3}; x (B,Hin,Win,Cin) NHWC; w repacked (Cout, ks*ks*Cin) tap-major
 * (cvar_repack_conv_weight); optional fused input transform in = [silu](x*a[n,c] + b[n,c]) (GroupNorm [+SiLU]);
 * upsample2x != 0 applies nearest x2 to the input first (Upsample2x, vae_modules.py:27-28), output is
 * (B, Hout, Wout, Cout) with Hout = Hin * (upsample2x ? 2 : 1); resid (same shape as out) is added when not NULL.
 * out_mode 0: NHWC fp32.  out_mode 1: final image - clamp(-1,1), (v+1)*0.5, written NCHW into
 * out[n, c, row_offset + y, x] of a (B, Cout, out_rows_total, Wout) tensor   (vqvae.py:89, control_var.py:563-565).
 * out_mode 2: as 1 without the (v+1)*0.5 step (plain VQVAE.fhat_to_img).
 * out_mode 3: NCHW planes as in 1, but neither clamp nor shift (quant_conv output for the quantiser, vqvae.py:74). */
typedef struct {
  const float* x; const float* w; const float* w_hi; const float* w_lo; const float* bias; float* out;
  const float* in_a; const float* in_b; int in_silu;
  const float* resid;
  int B, Hin, Win, Cin, Cout, ks, upsample2x;
  int out_mode, out_rows_total, row_offset;
  int engine;   /* -1 = library default (cvar_set_gemm_engine); 0 = force the SIMT fp32 engine for this call */
  /* FP16-pair operands (engine 4): when x16_hi is set the convolution runs on the 2-CTA tcgen05 kernel as an implicit
   * GEMM whose activation windows are fetched by 4-D TMA (zero fill = padding).  x16_* is the INPUT as NHWC half pairs
   * (B, Hin, Win, Cin), already normalised / activated / upsampled by its producer (cvar_affine_nc,
   * cvar_upsample2x_split_f16, cvar_split_f16): x, in_a, in_b, upsample2x must be NULL / 0.  w16_* is the pair of the
   * repacked weight.  Shapes: see cvar_conv2d_f16_supported. */
  const void* x16_hi; const void* x16_lo; const void* w16_hi; const void* w16_lo;
  /* Downsample2x of the encoder (vae_modules.py:31-37): F.pad(x, (0,1,0,1)) + 3x3 conv with stride 2 and no padding.
   * Output is (B, Hin/2, Win/2, Cout).  ks = 3, even Hin / Win, fp32 NHWC input; always on the SIMT fp32 engine. */
  int downsample2x;
  /* FP16-pair path only: 3 = run the three kernel rows of a 3x3 convolution as three accumulations whose partial sums are
   * added in fp32 (round to nearest) in the output - for the K = 9*640 layers, where one truncating tensor-core
   * accumulation of 360 steps costs too much accuracy.  0 / 1 = one accumulation. out_mode 0 only. */
  int ksplit;
  /* Image output (out_mode 1..3) of a STACKED batch: when > 0, image n = s * out_samples + b of the B images lands in
   * sample b of an (out_samples, Cout, out_rows_total, Wout) tensor at rows s * Hout + row_offset.  Decodes the control
   * and image halves of control_var.py:563-565 in one pass of 2 x out_samples images.  0: image n is sample n. */
  int out_samples;
  /* FP16-pair path, out_mode 0: GroupNorm statistics of the OUTPUT, produced by the epilogue (saves the read pass of
   * cvar_gn_stats over the activation).  gn_part: 2 * B * gn_groups * (Hout*Wout/32) doubles - per image, group and
   * 32-pixel slot the (sum, sum of squares) of the stored values; feed it to cvar_gn_finalize_parts.  Only for layers
   * cvar_conv2d_gn_fusable() accepts (an error otherwise: never silently skipped).  NULL: off. */
  double* gn_part; int gn_groups;
} cvar_conv_args;
CVAR_API int cvar_conv2d(const cvar_conv_args* args, void* stream);
/* 1 when the FP16-pair kernel takes this layer: ks in {1,3}, Cin % 32 == 0, Cout a multiple of one of
 * {256,160,128,224,192,96,64,32}, and W | 128 or 128 | W with whole 128-pixel tiles inside an image. */
CVAR_API int cvar_conv2d_f16_supported(int H, int W, int Cin, int Cout, int ks);
/* 1 when the FP16-pair kernel can emit GroupNorm partials for this layer (cvar_conv_args.gn_part). */
CVAR_API int cvar_conv2d_gn_fusable(int H, int W, int Cin, int Cout, int ks, int groups);
/* a[n,c], b[n,c] as cvar_gn_stats, from the partials a convolution wrote into gn_part (B images of HW pixels, C channels). */
CVAR_API int cvar_gn_finalize_parts(const double* gn_part, const float* gamma, const float* beta, float* a_out, float* b_out,
                           int B, int HW, int C, int groups, float eps, void* stream);
/* nearest x2 upsampling (Upsample2x, vae_modules.py:27-28) fused with the FP16-pair split:
 * x (B,H,W,C) fp32 NHWC -> hi / lo (B,2H,2W,C) halves.  C % 4 == 0. */
CVAR_API int cvar_upsample2x_split_f16(const float* x_nhwc, void* hi, void* lo, int B, int H, int W, int C, void* stream);
/* hi = w with the 13 low mantissa bits cleared (what a TF32 tensor-core operand keeps), lo = w - hi (exact): the
 * error-compensated 3xTF32 operands of the tcgen05 engine.  n must be a multiple of 4. */
CVAR_API int cvar_split_tf32(const float* w, float* w_hi, float* w_lo, long long n, void* stream);
/* FP16 pair of an fp32 array: hi = half_rn(x), lo = half_rn((x - hi) * 2^11), so x ~= hi + lo * 2^-11 with
 * |error| <= 2^-24 |x| (the residual is scaled to stay in fp16's normal range; the engine folds 2^-11 back in its
 * epilogue).  Three kind::f16 MMAs on such pairs (hi*hi, hi*lo, lo*hi) give fp32-class products at twice the TF32
 * tensor-core rate.  |x| must be < 65504 (saturates otherwise).  n must be a multiple of 4. */
CVAR_API int cvar_split_f16(const float* x, void* hi, void* lo, long long n, void* stream);
/* (Cout,Cin,ks,ks) -> (Cout, ks*ks*Cin), k index = (ky*ks+kx)*Cin + ci */
CVAR_API int cvar_repack_conv_weight(const float* w_oihw, float* w_out, int Cout, int Cin, int ks, void* stream);
/* (Cout,Cin,ks,ks) -> (Cout, ks*ks*Cin_pad) with zero weights for the padding channels */
CVAR_API int cvar_repack_conv_weight_pad(const float* w_oihw, float* w_out, int Cout, int Cin, int ks, int Cin_pad,
                                void* stream);
/* y = [silu](x*a[n,c] + b[n,c]): GroupNorm (+ SiLU) applied once (vae_modules.py:58-59, AttnBlock.norm).
 * y16_hi / y16_lo (optional, halves, same shape): the result as an FP16 pair; y may then be NULL. */
CVAR_API int cvar_affine_nc(const float* x_nhwc, const float* a, const float* b, float* y, void* y16_hi, void* y16_lo,
                   int B, int HW, int C, int silu, void* stream);
/* in-place row softmax of a (rows, cols) matrix (AttnBlock, vae_modules.py:84) */
CVAR_API int cvar_softmax_rows(float* x, int rows, int cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CVAR_H_ */
