"""GPU: the FP16-pair attention path of engine 4 - cvar_qkv_project16 (q / K / V^T written as pairs) and
cvar_attn_kvcache16 (tcgen05 kind::f16 kernel, two CTAs per SM, and the SIMT kernel on the same operands) - against fp64
SDPA and against SelfAttention.forward with torch.cat cache growth (basic_var.py:89-119)."""
import math

import pytest
import torch
import torch.nn.functional as F

from controlvar_b200 import ops
from oracle import controlvar_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def g(t):
    return t.to(DEV).contiguous()


def pair(x):
    return ops.F16Pair.from_tensor(g(x))


def pair_qk(x):
    """q / K operand format of the f16 attention path: 16 x = hi + lo (include/cvar.h)."""
    return ops.F16Pair.from_tensor_qk(g(x))


def fill_cache(kv, k, v, L):
    """Write K (R,H,L,64) / V into a KVCache16 through the library's own split."""
    R, H = kv.R, kv.H
    kp = pair_qk(k)
    vp = pair(v.transpose(2, 3))                     # (R,H,64,L)
    kv.k_hi.view(R, H, kv.T, 64)[:, :, :L] = kp.hi
    kv.k_lo.view(R, H, kv.T, 64)[:, :, :L] = kp.lo
    kv.vt_hi.view(R, H, 64, kv.T)[:, :, :, :L] = vp.hi
    kv.vt_lo.view(R, H, 64, kv.T)[:, :, :, :L] = vp.lo


@pytest.mark.parametrize("R,H,l,L", [(2, 2, 64, 64), (2, 3, 128, 310), (1, 2, 200, 510), (2, 2, 338, 848),
                                     (1, 4, 512, 1360), (3, 1, 72, 182), (1, 1, 130, 131), (2, 5, 8, 10), (2, 3, 32, 60),
                                     (3, 2, 50, 110)])
def test_attn16_matches_sdpa(R, H, l, L):
    torch.manual_seed(l + L)
    T = L + 5
    q = torch.randn(R, H, l, 64) * 2
    k = torch.randn(R, H, L, 64)
    v = torch.randn(R, H, L, 64)
    scale = 1 / 32
    ref = F.scaled_dot_product_attention(q.double(), k.double(), v.double(), scale=scale).transpose(1, 2).reshape(R, l, H * 64)
    kv = ops.KVCache16(R, H, T, DEV)
    fill_cache(kv, k, v, L)
    # poison the stale tail [L, T): it must be masked, not read into the result
    kv.k_hi.view(R, H, kv.T, 64)[:, :, L:] = 1e4
    kv.vt_hi.view(R, H, 64, kv.T)[:, :, :, L:] = -1e4
    q16 = pair_qk(q)
    res = {}
    for eng in (1, 0):
        out = torch.full((R, l, H * 64), float("nan"), device=DEV)
        o16 = ops.F16Pair.empty((R, l, H * 64), DEV)
        ops.attn_kvcache16(q16, kv, out, R, H, l, L, scale, engine=eng, out16=o16)
        torch.cuda.synchronize()
        res[eng] = (out.cpu().double() - ref).abs().max().item()
        assert (o16.float() - out).abs().max().item() < 1e-6       # the pair output is the split of the fp32 one
    print(f"\n[attn16-accuracy] R={R} H={H} l={l} L={L}: " + "  ".join(f"engine {e} err {x:.3e}" for e, x in res.items()))
    assert all(x < 2e-5 for x in res.values()), res
    # fp32 class: the pair format carries 22 mantissa bits, the reference fp32 SDPA has ~1e-6 here
    assert all(x < 4e-6 for x in res.values()), res


@pytest.mark.parametrize("L", [576, 600, 1360])
def test_attn16_reference_rebase_paths(L):
    """The tensor-core kernel keeps a lazily updated softmax reference and leaves O accumulating in TMEM for 4 tiles.
    Keys whose logits grow tile after tile force the reference to move at every position of the 4-tile drain cycle
    (rescale of registers AND of the row's TMEM accumulators); some rows never move it (queries pointing the other way)."""
    torch.manual_seed(L)
    R, H, l = 2, 2, 128
    scale = 1.0
    q = torch.randn(R, H, l, 64)
    k = torch.randn(R, H, L, 64) * 0.3
    # a direction every query shares (with either sign) and that keys follow more strongly the later they come
    u = torch.nn.functional.normalize(torch.randn(64), dim=0)
    sign = torch.where(torch.arange(l) % 3 == 0, -1.0, 1.0).view(1, 1, l, 1)
    q = q * 0.5 + 6.0 * sign * u
    ramp = (torch.arange(L) // 64).float().view(1, 1, L, 1)            # + ~7 nats per tile for the '+' rows
    k = k + 1.2 * ramp * u
    v = torch.randn(R, H, L, 64)
    ref = F.scaled_dot_product_attention(q.double(), k.double(), v.double(), scale=scale).transpose(1, 2).reshape(R, l, H * 64)
    kv = ops.KVCache16(R, H, L, DEV)
    fill_cache(kv, k, v, L)
    q16 = pair_qk(q)
    # reference with the operands the kernel really sees (the pair format is exact to 2^-24, the logits reach ~150)
    qd, kd, vd = q16.float_qk().double().cpu(), kv.keys(L).double().cpu(), kv.values(L).double().cpu()
    ref_pair = F.scaled_dot_product_attention(qd, kd, vd, scale=scale).transpose(1, 2).reshape(R, l, H * 64)
    errs = {}
    for eng in (1, 0):
        out = torch.empty(R, l, H * 64, device=DEV)
        ops.attn_kvcache16(q16, kv, out, R, H, l, L, scale, engine=eng)
        errs[eng] = ((out.cpu().double() - ref_pair).abs().max().item(), (out.cpu().double() - ref).abs().max().item())
    s_rng = (qd @ kd.transpose(-1, -2)).abs().max().item()
    print(f"\n[attn16-rebase] L={L} max|S| {s_rng:.0f}: tcgen05 err {errs[1][0]:.2e} (vs fp64 of fp32 inputs {errs[1][1]:.2e})"
          f"   SIMT err {errs[0][0]:.2e}")
    # the logits themselves are fp32 numbers of size max|S| (ulp 1.5e-5 at 190): the error of exp() scales with them.
    # Measured on B200: tcgen05 0.9e-7..1.0e-7 x max|S|, SIMT 2.2e-7..3.2e-7 x max|S| (fp32 FMA chains of 64 terms)
    assert errs[1][0] < 2.5e-7 * s_rng and errs[0][0] < 6e-7 * s_rng
    assert errs[1][1] < max(3e-5, 1.5e-6 * s_rng)


def test_attn16_pair_only_output_and_default_engine():
    """out = None (what the sampler passes) and engine = -1 (tensor cores for l >= 64, SIMT below)."""
    torch.manual_seed(0)
    R, H, T = 2, 2, 200
    kv = ops.KVCache16(R, H, T, DEV)
    k, v = torch.randn(R, H, T, 64), torch.randn(R, H, T, 64)
    fill_cache(kv, k, v, T)
    for l, L in ((18, 28), (128, 200)):
        q = torch.randn(R, H, l, 64)
        ref = F.scaled_dot_product_attention(q.double(), k[:, :, :L].double(), v[:, :, :L].double(), scale=0.125) \
            .transpose(1, 2).reshape(R, l, H * 64)
        o16 = ops.F16Pair.empty((R, l, H * 64), DEV)
        ops.attn_kvcache16(pair_qk(q), kv, None, R, H, l, L, 0.125, out16=o16)
        assert (o16.float().cpu().double() - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("cos_attn", [False, True])
def test_qkv_project16_and_attention_three_scales(cos_attn):
    """Three consecutive scales through cvar_qkv_project16 + cvar_attn_kvcache16 vs SelfAttention.forward, incl. ragged
    l / L that are not multiples of the 64-key tile and the cosine-attention normalisation of depth 30."""
    torch.manual_seed(4)
    R, H = 3, 4
    C = H * 64
    T = 2 + 50 + 130
    sd = {"q_bias": torch.randn(C) * 0.1, "zero_k_bias": torch.zeros(C), "v_bias": torch.randn(C) * 0.1,
          "mat_qkv.weight": torch.randn(3 * C, C) / math.sqrt(C), "proj.weight": torch.eye(C), "proj.bias": torch.zeros(C),
          "scale_mul_1H11": torch.tensor([1.0, 1.386, 3.0, 5.0]).view(1, H, 1, 1)}
    scale = 1.0 if cos_attn else 0.25 / math.sqrt(64)
    cache = {}
    kv = ops.KVCache16(R, H, T, DEV)
    sdg = {k: g(v) for k, v in sd.items()}
    sm = sdg["scale_mul_1H11"].reshape(-1).contiguous()
    wq = ops.SplitWeight(sdg["mat_qkv.weight"], f16=True)
    L = 0
    for l in (2, 50, 130):
        x = torch.randn(R, l, C)
        ref = O.self_attention(x, sd, "", H, cache, cos_attn, scale)      # proj is the identity here
        q16 = ops.F16Pair.empty((R, H, l, 64), DEV)
        ops.qkv_project16(pair(x.reshape(R * l, C)), wq, sdg["q_bias"], sdg["zero_k_bias"], sdg["v_bias"], q16, kv, R, l, L,
                          H, cos_attn, sm if cos_attn else None)
        L += l
        assert (kv.keys(L).cpu() - cache["k"]).abs().max().item() < 2e-5
        assert (kv.values(L).cpu() - cache["v"]).abs().max().item() < 2e-5
        # tolerance scales with the logit range, as in test_gpu_ops.test_qkv_project_and_kvcache_attention
        qf = q16.float_qk()
        s_max = (qf.double().cpu() @ kv.keys(L).double().cpu().transpose(-1, -2)).abs().max().item() * scale
        tol = max(3e-5, 1.5e-6 * s_max)
        for eng in ((0, 1) if l >= 50 else (0,)):
            out = torch.empty(R, l, C, device=DEV)
            ops.attn_kvcache16(q16, kv, out, R, H, l, L, scale, engine=eng)
            err = (out.cpu() - ref).abs().max().item()
            assert err < tol, f"l={l} L={L} engine={eng}: {err:.3e} (max|S| {s_max:.1f}, tol {tol:.1e})"


@pytest.mark.parametrize("scale,qmul", [(1 / 32, 1.0), (1.0, 3.0)])
def test_attn16_last_scale_shape_two_ctas_per_sm(scale, qmul):
    """The bench's last scale on a FULL machine: 3072 CTAs, every SM holding two, the MMA / TMA threads lagging behind
    the softmax warps (the regime in which a one-bit mbarrier parity can alias: two versions of this kernel passed every
    small-grid test and failed here).  Run twice: bit-identical; and equal to the SIMT kernel on the same operands.
    The second parameter set has logits of +-100: the softmax reference moves, with rescales of the TMEM accumulators."""
    torch.manual_seed(9)
    R, H, l, L = 32, 24, 512, 1360
    kv = ops.KVCache16(R, H, L, DEV)
    kv.k_hi.normal_().mul_(16), kv.vt_hi.normal_()
    kv.k_lo.normal_().mul_(2.0 ** -8), kv.vt_lo.normal_()       # qk pairs: residual of a 16 x ~ N(0, 16) value is ~2^-8
    q16 = ops.F16Pair.empty((R, H, l, 64), DEV)
    q16.hi.normal_().mul_(16 * qmul), q16.lo.normal_().mul_(2.0 ** -8)
    outs = []
    for _ in range(2):
        o = torch.empty(R, l, H * 64, device=DEV)
        ops.attn_kvcache16(q16, kv, o, R, H, l, L, scale, engine=1)
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    o0 = torch.empty(R, l, H * 64, device=DEV)
    ops.attn_kvcache16(q16, kv, o0, R, H, l, L, scale, engine=0)
    err = (outs[0] - o0).abs().max().item()
    s_max = 8.0 * qmul * 8.0 * scale * 4          # |q||k| ~ 8 * 8 per unit qmul, a few sigma
    print(f"\n[attn16-load] scale {scale:.3f} qmul {qmul}: max |tcgen05 - SIMT| = {err:.2e}")
    assert err < max(2e-6, 4e-7 * s_max)


@pytest.mark.parametrize("pns,R,H", [((1, 2, 3, 4, 5, 6, 8, 10, 13, 16), 2, 3), ((1, 2, 3), 3, 2), ((1, 2, 3, 4, 5, 6, 8, 10), 1, 1)])
def test_attn16_block_causal_one_launch(pns, R, H):
    """cvar_attn_blockcausal16: the whole pyramid in one launch, a query of scale s seeing the keys of scales <= s
    (control_var.py:168) - against fp64 SDPA with the explicit boolean mask."""
    lens = [2 * p * p for p in pns]
    L = sum(lens)
    torch.manual_seed(L)
    q, k, v = torch.randn(R, H, L, 64) * 2, torch.randn(R, H, L, 64), torch.randn(R, H, L, 64)
    lvl = torch.cat([torch.full((n,), i) for i, n in enumerate(lens)])
    mask = lvl.view(L, 1) >= lvl.view(1, L)
    scale = 1 / 32
    ref = F.scaled_dot_product_attention(q.double(), k.double(), v.double(), attn_mask=mask, scale=scale)
    ref = ref.transpose(1, 2).reshape(R, L, H * 64)
    kv = ops.KVCache16(R, H, L, DEV)
    fill_cache(kv, k, v, L)
    out = torch.full((R, L, H * 64), float("nan"), device=DEV)
    o16 = ops.F16Pair.empty((R, L, H * 64), DEV)
    ops.attn_blockcausal16(pair_qk(q), kv, out, R, H, lens, scale, out16=o16)
    e = ((out.cpu().double() - ref).abs().max() / ref.abs().max()).item()
    assert e < 2e-6, e
    assert torch.equal(o16.float().cpu(), ops.F16Pair.from_tensor(out).float().cpu())
    with pytest.raises(Exception):
        ops.attn_blockcausal16(pair_qk(q), kv, out, R, H, lens + [8], scale)         # more tokens than the cache holds
