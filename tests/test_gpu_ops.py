"""GPU: every C-ABI entry point against the CPU oracle (fp32 ATen) on seeded inputs.

Tolerances (floating point path): fp32 kernels differ from the oracle only by summation
order, so activations must agree to ~1e-5 relative; token / index outputs must be IDENTICAL wherever the
oracle's own top-2 margin is above the numerical resolution (1e-5), and the tests count how many rows are below it.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from controlvar_b200 import ops, weights as W
from controlvar_b200.config import PathConfig
from controlvar_b200.control_var import bicubic_matrix
from oracle import controlvar_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def g(t):
    return t.to(DEV).contiguous()


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


@pytest.fixture(params=[0, 1], ids=["simt", "tc3xtf32"])
def engine(request):
    old = ops.set_gemm_engine(request.param)
    yield request.param
    ops.set_gemm_engine(old)


# ------------------------------------------------------------------------------------------------ basics
@pytest.mark.parametrize("M,C,l", [(6, 256, 3), (260, 768, 130), (1024, 1536, 512), (40, 1920, 8)])
def test_ln_modulate(M, C, l):
    torch.manual_seed(0)
    R = M // l + (1 if M % l else 0)
    x = torch.randn(M, C) * 3 + 0.5
    ada = torch.randn(R, 6 * C) * 0.3
    scale, shift = ada[:, 2 * C:3 * C], ada[:, 4 * C:5 * C]
    rows = torch.arange(M) // l
    ref = O.ln_modulate(x, scale[rows], shift[rows])
    ada_g = g(ada)
    out = torch.empty(M, C, device=DEV)
    ops.ln_modulate(g(x), ada_g[:, 2 * C:3 * C], ada_g[:, 4 * C:5 * C], 6 * C, out, M, C, l, 1e-6)
    assert (out.cpu() - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("C,l", [(768, 77), (1024, 50), (1536, 338), (1920, 200), (2048, 128)])
def test_ln_modulate_streaming_kernel_is_bit_identical(C, l):
    """Large-M FP16-pair calls take the persistent bulk-copy kernel (csrc/ln_stream.cu); the fp32-output call of the same
    rows takes the warp-per-row kernel.  Same arithmetic: the pair must be exactly the split of the fp32 result, for a row count
    that is not a multiple of anything (ragged last round of every warp's ring), and the fp32 result must match the oracle."""
    torch.manual_seed(5)
    M = 2 * 148 * 16 + 1237
    R = M // l + 1
    x = g(torch.randn(M, C) * 3 + 0.5)
    ada = g(torch.randn(R, 6 * C) * 0.3)
    scale, shift = ada[:, 2 * C:3 * C], ada[:, 4 * C:5 * C]
    y32 = torch.empty(M, C, device=DEV)
    ops.ln_modulate(x, scale, shift, 6 * C, y32, M, C, l, 1e-6)
    pair = ops.F16Pair.empty((M, C), DEV)
    pair.hi.fill_(7.0), pair.lo.fill_(7.0)
    ops.ln_modulate(x, scale, shift, 6 * C, None, M, C, l, 1e-6, out16=pair)
    want = ops.F16Pair.from_tensor(y32)
    assert torch.equal(pair.hi, want.hi) and torch.equal(pair.lo, want.lo)
    rows = torch.arange(M) // l
    ref = O.ln_modulate(x.cpu(), scale.cpu()[rows], shift.cpu()[rows])
    assert (y32.cpu() - ref).abs().max().item() < 2e-5


def test_prologue_and_lvl_pos():
    cfg = PathConfig(depth=2, patch_nums=(1, 2, 3))
    sd = W.synthetic_var_state_dict(cfg, 0)
    C, B = cfg.C, 3
    lp = ops.lvl_pos(g(sd["lvl_embed.weight"]), g(sd["lvl_1L"]), g(sd["pos_1LC"]), torch.empty(cfg.L, C, device=DEV))
    lp_ref = F.embedding(sd["lvl_1L"], sd["lvl_embed.weight"]) + sd["pos_1LC"]
    assert torch.equal(lp.cpu(), lp_ref[0])
    label, ct = torch.tensor([5, 999, 1000]), torch.tensor([0, 3, 2])
    cond, silu, x0 = (torch.empty(2 * B, C, device=DEV), torch.empty(2 * B, C, device=DEV),
                      torch.empty(2 * B, 2, C, device=DEV))
    ops.prologue(g(sd["class_emb.weight"]), g(sd["cond_embed.weight"]), g(sd["pos_start"]), lp, g(label), g(ct), 1000,
                 cond, silu, x0)
    lab2 = torch.cat((label, torch.full_like(label, 1000)))
    ct2 = torch.cat((ct, torch.full_like(ct, 4)))
    sos = F.embedding(lab2, sd["class_emb.weight"])
    ref = torch.cat([F.embedding(ct2, sd["cond_embed.weight"]).unsqueeze(1), sos.unsqueeze(1)], 1) \
        + sd["pos_start"].expand(2 * B, 2, -1) + lp_ref[:, :2]
    assert torch.equal(cond.cpu(), sos)
    assert torch.equal(x0.cpu(), ref)
    assert (silu.cpu() - F.silu(sos)).abs().max().item() < 1e-6


# -------------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(4, 64, 32), (100, 192, 64), (300, 256, 128), (257, 4096, 256), (1000, 160, 1440),
                                   (2048, 768, 3072)])
def test_gemm_bias_and_gelu(engine, M, N, K):
    torch.manual_seed(1)
    A, Wt, b = torch.randn(M, K), torch.randn(N, K) / math.sqrt(K), torch.randn(N)
    ref = (A.double() @ Wt.double().T + b.double())
    out = torch.empty(M, N, device=DEV)
    ops.gemm(g(A), ops.SplitWeight(g(Wt)), g(b), out, M, N, K)
    assert rel_err(out.cpu().double(), ref) < 1e-5
    ops.gemm(g(A), ops.SplitWeight(g(Wt)), g(b), out, M, N, K, epilogue=ops.EPI_BIAS_GELU)
    assert rel_err(out.cpu().double(), F.gelu(ref, approximate="tanh")) < 1e-5


def test_gemm_gamma_residual(engine):
    torch.manual_seed(2)
    R, l, C, K = 3, 50, 256, 512
    M = R * l
    A, Wt, b = torch.randn(M, K), torch.randn(C, K) / math.sqrt(K), torch.randn(C)
    x0 = torch.randn(M, C)
    ada = torch.randn(R, 6 * C)
    gamma = ada[:, C:2 * C]
    ref = x0.double() + (A.double() @ Wt.double().T + b.double()) * gamma.double().repeat_interleave(l, 0)
    x = g(x0)
    ada_g = g(ada)
    ops.gemm(g(A), ops.SplitWeight(g(Wt)), g(b), x, M, C, K, epilogue=ops.EPI_BIAS_GAMMA_RESID, gamma=ada_g[:, C:2 * C],
             gamma_row_stride=6 * C, rows_per_sample=l)
    assert rel_err(x.cpu().double(), ref) < 1e-5


def test_gemm_batched_strided_and_kn(engine):
    torch.manual_seed(3)
    Bn, HW, Cn = 3, 256, 640
    qkv = torch.randn(Bn, HW, 3 * Cn) * 0.2
    q, k, v = qkv[..., :Cn], qkv[..., Cn:2 * Cn], qkv[..., 2 * Cn:]
    alpha = Cn ** -0.5
    S_ref = torch.bmm(q.double(), k.double().transpose(1, 2)) * alpha
    qg = g(qkv)
    S = torch.empty(Bn, HW, HW, device=DEV)
    ops.gemm(qg, qg[:, :, Cn:], None, S, HW, HW, Cn, lda=3 * Cn, ldw=3 * Cn, ldo=HW, alpha=alpha, batch=Bn,
             strideA=HW * 3 * Cn, strideW=HW * 3 * Cn, strideO=HW * HW)
    assert rel_err(S.cpu().double(), S_ref) < 1e-5
    ops.softmax_rows(S, Bn * HW, HW)
    P_ref = S_ref.float().softmax(-1)
    assert (S.cpu() - P_ref).abs().max().item() < 1e-6
    h = torch.empty(Bn, HW, Cn, device=DEV)
    ops.gemm(S, qg[:, :, 2 * Cn:], None, h, HW, Cn, HW, lda=HW, ldw=3 * Cn, ldo=Cn, w_is_kn=True, batch=Bn,
             strideA=HW * HW, strideW=HW * 3 * Cn, strideO=HW * Cn)
    h_ref = torch.bmm(S.cpu().double(), v.double())
    assert rel_err(h.cpu().double(), h_ref) < 1e-5
    resid = torch.randn(Bn * HW, Cn)
    Wp, bp = torch.randn(Cn, Cn) / math.sqrt(Cn), torch.randn(Cn)
    out = torch.empty(Bn * HW, Cn, device=DEV)
    ops.gemm(h.view(-1, Cn), ops.SplitWeight(g(Wp)), g(bp), out, Bn * HW, Cn, Cn, epilogue=ops.EPI_BIAS_RESID, resid=g(resid))
    ref = resid.double() + h.cpu().double().view(-1, Cn) @ Wp.double().T + bp.double()
    assert rel_err(out.cpu().double(), ref) < 1e-5


# --------------------------------------------------------------------------------------- QKV + attention
@pytest.mark.parametrize("cos_attn", [False, True])
def test_qkv_project_and_kvcache_attention(engine, cos_attn):
    """Three consecutive scales through cvar_qkv_project + cvar_attn_kvcache vs SelfAttention.forward with
    torch.cat cache growth (basic_var.py:89-119), incl. ragged l / L that are not multiples of the 64-key tile."""
    torch.manual_seed(4)
    R, H = 3, 4
    C = H * 64
    T = 2 + 50 + 130
    sd = {"q_bias": torch.randn(C) * 0.1, "zero_k_bias": torch.zeros(C), "v_bias": torch.randn(C) * 0.1,
          "mat_qkv.weight": torch.randn(3 * C, C) / math.sqrt(C), "proj.weight": torch.eye(C), "proj.bias": torch.zeros(C),
          "scale_mul_1H11": torch.tensor([1.0, 1.386, 3.0, 5.0]).view(1, H, 1, 1)}
    scale = 1.0 if cos_attn else 0.25 / math.sqrt(64)
    cache = {}
    kv = ops.KVCache(R, H, T, DEV)
    sdg = {k: g(v) for k, v in sd.items()}
    sm = sdg["scale_mul_1H11"].reshape(-1).contiguous()
    wq = ops.SplitWeight(sdg["mat_qkv.weight"])
    L = 0
    for l in (2, 50, 130):
        x = torch.randn(R, l, C)
        ref = O.self_attention(x, sd, "", H, cache, cos_attn, scale)      # proj is the identity here
        q = torch.empty(R, H, l, 64, device=DEV)
        ops.qkv_project(g(x), wq, sdg["q_bias"], sdg["zero_k_bias"], sdg["v_bias"], q, kv, R, l, L, H, cos_attn,
                        sm if cos_attn else None)
        L += l
        assert (kv.keys(L).cpu() - cache["k"]).abs().max().item() < 2e-5
        assert (kv.values(L).cpu() - cache["v"]).abs().max().item() < 2e-5
        # Tolerance scales with the logit range.  Attention output error is ~ max|S| x (relative error of q, k): measured on
        # B200 (tools/diag_cos_attn.py, profiles/r01_cos_attn_error.md) the all-SIMT path gives 2e-7 / 2e-6 / 1e-5 for
        # heads with max|S| = 1.4 / 10 / 51, i.e. fp32 resolution of the logit itself.  Cosine attention multiplies q by
        # up to 100 (basic_var.py:100), so a flat absolute bound is wrong for it.  Both attention engines are within
        # 1.1x of each other on identical inputs; the tcgen05 GEMM's q is ~4.5x less accurate than the SIMT GEMM's.
        s_max = (q.double().cpu() @ kv.keys(L).double().cpu().transpose(-1, -2)).abs().max().item() * scale
        tol = max(3e-5, 1.5e-6 * s_max)
        for eng in ((0, 1) if l >= 50 else (0,)):
            out = torch.empty(R, l, C, device=DEV)
            ops.attn_kvcache(q, kv, out, R, H, l, L, scale, engine=eng)
            err = (out.cpu() - ref).abs().max().item()
            assert err < tol, f"l={l} L={L} engine={eng}: {err:.3e} (max|S| {s_max:.1f}, tol {tol:.1e})"


# ---------------------------------------------------------------------------------------------- sampling
@pytest.mark.parametrize("top_k,top_p,t", [(900, 0.96, 1.5), (0, 0.0, 0.0), (50, 0.0, 3.0), (0, 0.5, 0.75),
                                           (4096, 0.96, 1.0), (1, 0.0, 1.0)])
def test_cfg_sample_matches_reference_rule(top_k, top_p, t):
    torch.manual_seed(5)
    B, l, V = 3, 37, 4096
    logits = torch.randn(2 * B, l, V) * 1.5
    gen = torch.Generator().manual_seed(11)
    q = torch.empty(B * l, V).exponential_(1, generator=gen)
    mixed = O.cfg_combine(logits, B, t)
    masked = O.mask_top_k_top_p_(mixed.clone(), top_k, top_p)
    ref = O.multinomial1_with_noise(masked.softmax(-1).view(-1, V), q).view(B, l)
    margin = O.sampling_margin(masked, q).view(B, l)
    idx = torch.empty(B, l, dtype=torch.int64, device=DEV)
    ops.cfg_sample(g(logits), g(q), idx, B, l, V, t, top_k, top_p)
    clear = margin > 1e-5
    assert torch.equal(idx.cpu()[clear], ref[clear])
    assert (~clear).sum().item() <= 1

    # adversarial noise: tokens the reference masks out get q -> 0+, so any token that wrongly survives the
    # top-k / top-p filter would win the argmax
    removed = torch.isinf(masked).view(-1, V)
    q_adv = torch.where(removed, torch.full_like(q, 1e-30), q)
    ref_adv = O.multinomial1_with_noise(masked.softmax(-1).view(-1, V), q_adv).view(B, l)
    ops.cfg_sample(g(logits), g(q_adv), idx, B, l, V, t, top_k, top_p)
    n_bad = (idx.cpu() != ref_adv).sum().item()
    assert n_bad <= (1 if top_p > 0 else 0), f"{n_bad} rows picked a token the reference filtered out"


def test_cfg_sample_masked_and_gumbel_embed():
    """more_smooth (control_var.py:511-515, helpers.py:22-36): cvar_cfg_sample_masked returns the logits as
    sample_with_top_k_top_p_ leaves them in place and the same tokens as cvar_cfg_sample; cvar_gumbel_embed turns them
    into the Gumbel-softmax mixture of code vectors (4 replicas re-using the logit rows, as conditional_infer_cfg does)."""
    torch.manual_seed(21)
    B, l, V, Cv, t, top_k, top_p = 2, 18, 4096, 32, 1.2, 900, 0.96
    logits = torch.randn(2 * B, l, V) * 1.5
    gen = torch.Generator().manual_seed(3)
    q = torch.empty(B * l, V).exponential_(1, generator=gen)
    masked_ref = O.mask_top_k_top_p_(O.cfg_combine(logits, B, t).clone(), top_k, top_p)
    idx0 = torch.empty(B, l, dtype=torch.int64, device=DEV)
    ops.cfg_sample(g(logits), g(q), idx0, B, l, V, t, top_k, top_p)
    idx1 = torch.empty(B, l, dtype=torch.int64, device=DEV)
    masked = torch.empty(B * l, V, device=DEV)
    f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))
    ops.cfg_sample_masked(g(logits), g(q), idx1, masked, B, l, V, (f32(1 + t), -f32(t)), 1, top_k, top_p)
    assert torch.equal(idx0, idx1)
    mr = masked_ref.view(B * l, V)
    same_mask = torch.isinf(masked.cpu()) == torch.isinf(mr)
    assert (~same_mask).sum().item() <= 2                       # a top-p boundary entry may fall either way at fp32 resolution
    keep = ~torch.isinf(mr) & same_mask
    assert torch.equal(masked.cpu()[keep], mr[keep])            # the guidance mix is elementwise: bit-exact

    emb = torch.randn(V, Cv)
    rep, ratio = 4, 5 / 9
    e = torch.empty(rep * B * l, V).exponential_(1, generator=gen)
    ref = O.gumbel_soft_embedding(mr.view(B, l, V).repeat(rep, 1, 1), ratio, e, emb).view(rep * B * l, Cv)
    h = torch.empty(rep * B * l, Cv, device=DEV)
    ops.gumbel_embed(g(mr), g(e), g(emb), h, B * l, rep * B * l, V, Cv, 1 + ratio, max(0.27 * (1 - ratio * 0.95), 0.005))
    assert (h.cpu() - ref).abs().max().item() < 2e-5 * ref.abs().max().item()


# ----------------------------------------------------------------------------------------------- VQ step
def test_vq_step_all_scales():
    cfg = PathConfig(depth=2)
    pn_list = cfg.patch_nums
    vsd = W.synthetic_vae_state_dict(cfg, 0, with_encoder=False)
    sd = W.synthetic_var_state_dict(cfg, 0)
    C, B, hw, Cv = cfg.C, 2, pn_list[-1], 32
    torch.manual_seed(6)
    lvl_pos = torch.randn(cfg.L, C)
    f_hat_ref = torch.zeros(B, Cv, 2 * hw, hw)
    f_hat = torch.zeros(B, Cv, 2 * hw, hw, device=DEV)
    emb_g, lvl_g = g(vsd["quantize.embedding.weight"]), g(lvl_pos)
    ww, wb = g(sd["word_embed.weight"]), g(sd["word_embed.bias"])
    cur = 0
    for si, pn in enumerate(pn_list):
        cur += 2 * pn * pn
        idx = torch.randint(0, 4096, (B, 2 * pn * pn))
        h = F.embedding(idx, vsd["quantize.embedding.weight"]).transpose(1, 2)
        h1, h2 = h[:, :, :pn * pn].reshape(B, Cv, pn, pn), h[:, :, pn * pn:].reshape(B, Cv, pn, pn)
        f1, n1 = O.get_next_autoregressive_input(si, pn_list, f_hat_ref[:, :, :hw], h1, vsd)
        f2, n2 = O.get_next_autoregressive_input(si, pn_list, f_hat_ref[:, :, hw:], h2, vsd)
        f_hat_ref = torch.cat((f1, f2), 2)
        k = cfg.phi_index(si)
        pw, pb = g(vsd[f"quantize.quant_resi.qresi_ls.{k}.weight"]), g(vsd[f"quantize.quant_resi.qresi_ls.{k}.bias"])
        last = si == len(pn_list) - 1
        pn_next = 0 if last else pn_list[si + 1]
        U = None if pn == hw else g(bicubic_matrix(pn, hw))
        x_next = None if last else torch.empty(2 * B, 2 * pn_next * pn_next, C, device=DEV)
        ops.vq_step(g(idx), emb_g, U, pw, pb, ww, wb, None if last else lvl_g[cur:], f_hat, x_next, B, pn, pn_next, hw,
                    Cv, C)
        assert (f_hat.cpu() - f_hat_ref).abs().max().item() < 2e-5, f"f_hat at scale {si}"
        if not last:
            ntm = torch.cat((n1, n2), 2).reshape(B, Cv, -1).transpose(1, 2)
            ref = F.linear(ntm, sd["word_embed.weight"], sd["word_embed.bias"]) + lvl_pos[cur:cur + 2 * pn_next ** 2]
            assert (x_next[:B].cpu() - ref).abs().max().item() < 2e-5, f"next map at scale {si}"
            assert torch.equal(x_next[:B], x_next[B:])


def test_vq_nearest():
    cfg = PathConfig()
    emb = W.synthetic_vae_state_dict(cfg, 0, with_encoder=False)["quantize.embedding.weight"]
    torch.manual_seed(7)
    z = torch.randn(777, 32) * 1.2
    ref = O.vq_nearest(z, emb)
    idx = ops.vq_nearest(g(z), g(emb), torch.empty(777, dtype=torch.int64, device=DEV)).cpu()
    d = torch.cdist(z.double(), emb.double()) ** 2
    top2 = d.topk(2, dim=1, largest=False)[0]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-4
    assert torch.equal(idx[clear], ref[clear]) and clear.float().mean().item() > 0.99


# ----------------------------------------------------------------------------------------------- decoder
@pytest.mark.parametrize("cin,cout,ks,up,H", [(32, 32, 3, False, 16), (160, 160, 3, False, 24), (320, 160, 1, False, 16),
                                              (160, 160, 3, True, 12), (640, 320, 3, False, 16), (32, 640, 3, False, 16)])
def test_conv2d_with_fused_groupnorm_silu(engine, cin, cout, ks, up, H):
    torch.manual_seed(8)
    B, Wd = 2, H + 3
    x = torch.randn(B, cin, H, Wd) * 2 + 0.3
    w = torch.randn(cout, cin, ks, ks) / math.sqrt(cin * ks * ks)
    b = torch.randn(cout)
    gam, bet = torch.rand(cin) + 0.5, torch.randn(cin) * 0.1
    xin = F.silu(F.group_norm(x, 32, gam, bet, 1e-6))
    if up:
        xin = F.interpolate(xin, scale_factor=2, mode="nearest")
    ref = F.conv2d(xin.double(), w.double(), b.double(), padding=ks // 2)
    Ho, Wo = ref.shape[2], ref.shape[3]
    resid = torch.randn(B, cout, Ho, Wo)
    ref = ref + resid.double()
    x_nhwc = g(x.permute(0, 2, 3, 1))
    wp = ops.SplitWeight(ops.repack_conv_weight(g(w), torch.empty(cout, ks * ks * cin, device=DEV)))
    a = torch.empty(B, cin, device=DEV)
    bb = torch.empty(B, cin, device=DEV)
    scratch = torch.empty(2 * B * 32 * ops.gn_chunks(H * Wd), dtype=torch.float64, device=DEV)
    ops.gn_stats(x_nhwc, g(gam), g(bet), a, bb, scratch, B, H * Wd, cin)
    out = torch.empty(B, Ho, Wo, cout, device=DEV)
    ops.conv2d(x_nhwc, wp, g(b), out, B, H, Wd, cin, cout, ks, in_a=a, in_b=bb, in_silu=True,
               resid=g(resid.permute(0, 2, 3, 1)), upsample2x=up)
    got = out.cpu().permute(0, 3, 1, 2).double()
    assert rel_err(got, ref) < 2e-5


def test_conv_out_image_mode():
    torch.manual_seed(9)
    B, cin, H = 2, 160, 20
    x = torch.randn(B, cin, H, H)
    w = torch.randn(3, cin, 3, 3) / math.sqrt(cin * 9) * 3
    b = torch.randn(3) * 0.1
    ref = F.conv2d(x, w, b, padding=1).clamp_(-1, 1).add_(1).mul_(0.5)
    img = torch.zeros(B, 3, 2 * H, H, device=DEV)
    wp = ops.repack_conv_weight(g(w), torch.empty(3, 9 * cin, device=DEV))
    ops.conv2d(g(x.permute(0, 2, 3, 1)), wp, g(b), img, B, H, H, cin, 3, 3, out_mode=1, out_rows_total=2 * H,
               row_offset=H)
    assert (img[:, :, H:].cpu() - ref).abs().max().item() < 1e-5
    assert img[:, :, :H].abs().max().item() == 0


@pytest.mark.parametrize("H,Wd,cin", [(8, 128, 32), (12, 256, 160), (4, 128, 16)])
def test_conv_out_image_mode_wide(H, Wd, cin):
    """Wide images (W a multiple of 128, H of 4) take the 4-rows-per-thread shared-memory kernel: zero padding on all four sides,
    several row tiles, several 16-channel chunks, stacked halves (out_samples) - against fp64."""
    torch.manual_seed(13)
    B = 4
    x = torch.randn(B, cin, H, Wd)
    w = torch.randn(3, cin, 3, 3) / math.sqrt(cin * 9) * 3
    b = torch.randn(3) * 0.1
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1).clamp_(-1, 1).add_(1).mul_(0.5)
    img = torch.zeros(B // 2, 3, 2 * H, Wd, device=DEV)
    wp = ops.repack_conv_weight(g(w), torch.empty(3, 9 * cin, device=DEV))
    # maps 0..1 are the upper halves of images 0..1, maps 2..3 the lower halves (the sampler's stacked decode)
    ops.conv2d(g(x.permute(0, 2, 3, 1)), wp, g(b), img, B, H, Wd, cin, 3, 3, out_mode=1, out_rows_total=2 * H, row_offset=0,
               out_samples=B // 2)
    got = torch.cat((img[:, :, :H], img[:, :, H:]), 0).cpu().double()
    assert (got - ref).abs().max().item() < 1e-5


def test_fhat_to_img_matches_oracle_decoder():
    """VQVAE.fhat_to_img end to end (vqvae.py:88-89): pixels within 1e-4 abs of the fp32 oracle."""
    from controlvar_b200 import VQVAE
    cfg = PathConfig()
    vsd = W.synthetic_vae_state_dict(cfg, 0)
    vae = VQVAE(ch=160).to(DEV)
    vae.load_state_dict(vsd)
    torch.manual_seed(10)
    f_hat = torch.randn(1, 32, 32, 16) * 1.5
    ref = O.fhat_to_img(f_hat[:, :, 16:, :], vsd)
    got = vae.fhat_to_img(g(f_hat)[:, :, 16:, :])
    assert got.shape == ref.shape == (1, 3, 256, 256)
    assert (got.cpu() - ref).abs().max().item() < 1e-4


def test_gemm_ragged_scalar_path():
    """K / ld not multiples of 4 (decoder attention of truncated pyramids: 3x3 -> HW = 9, 5x5 -> HW = 25)."""
    torch.manual_seed(12)
    Bn, HW, Cn = 2, 25, 64
    qkv = torch.randn(Bn, HW, 3 * Cn)
    qg = g(qkv)
    S = torch.empty(Bn, HW, HW, device=DEV)
    ops.gemm(qg, qg[:, :, Cn:], None, S, HW, HW, Cn, lda=3 * Cn, ldw=3 * Cn, ldo=HW, alpha=0.125, batch=Bn,
             strideA=HW * 3 * Cn, strideW=HW * 3 * Cn, strideO=HW * HW)
    S_ref = torch.bmm(qkv[..., :Cn].double(), qkv[..., Cn:2 * Cn].double().transpose(1, 2)) * 0.125
    assert rel_err(S.cpu().double(), S_ref) < 1e-5
    h = torch.empty(Bn, HW, Cn, device=DEV)
    ops.gemm(S, qg[:, :, 2 * Cn:], None, h, HW, Cn, HW, lda=HW, ldw=3 * Cn, ldo=Cn, w_is_kn=True, batch=Bn,
             strideA=HW * HW, strideW=HW * 3 * Cn, strideO=HW * Cn)
    assert rel_err(h.cpu().double(), torch.bmm(S.cpu().double(), qkv[..., 2 * Cn:].double())) < 1e-5
