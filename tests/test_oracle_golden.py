"""CPU: the oracle restatement reproduces the reference's own outputs (goldens made by oracle/make_golden.py)."""
import pytest
import torch

from controlvar_b200 import weights as W
from oracle import controlvar_oracle as O
from golden_util import golden_names, load_golden

FAST = [n for n in golden_names() if "pn10" not in n]


def test_goldens_exist():
    assert len(golden_names()) >= 4


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    m, cfg = g["meta"], g["cfg"]
    sd = W.synthetic_var_state_dict(cfg, m["weight_seed"])
    vsd = W.synthetic_vae_state_dict(cfg, m["weight_seed"], with_encoder=False)
    out = O.autoregressive_infer_cfg(sd, vsd, cfg.patch_nums, cfg.depth, m["B"], torch.tensor(m["labels"]),
                                     torch.tensor(m["cond"]), m["cfg"], m["top_k"], m["top_p"],
                                     O.cpu_generator_noise(m["seed"]), embed_dim=cfg.embed_dim, num_heads=cfg.heads,
                                     more_smooth=bool(m.get("more_smooth", False)))
    for si, (a, b) in enumerate(zip(g["idx"], out["idx"])):
        assert torch.equal(a, b), f"tokens differ at scale {si}"
    sub = m["img_sub"]
    assert list(out["img"].shape) == m["img_shape"]
    # same ATen ops in the same order: bit-exact on the host that made the golden, <=1e-6 across CPU ISAs
    assert (out["img"][:, :, ::sub, ::sub] - g["img_sub"]).abs().max().item() <= 1e-6
    assert (out["f_hat"] - g["f_hat"]).abs().max().item() <= 1e-6


@pytest.mark.parametrize("name", golden_names("enc"))
def test_oracle_img_to_idxBl_matches_reference_golden(name):
    """VQVAE.img_to_idxBl (vqvae.py:73-75, quant.py:184-215): encoder output and the tokens of every scale."""
    g = load_golden(name)
    m, cfg = g["meta"], g["cfg"]
    vsd = W.synthetic_vae_state_dict(cfg, m["weight_seed"])
    img = W.synthetic_image(m["B"], cfg.img_hw, m["img_seed"])
    f = O.img_to_f(img, vsd)
    assert (f - g["f"]).abs().max().item() <= 1e-6
    # tokens from the golden's own f: independent of sub-ulp differences of the encoder across CPU ISAs
    for si, (a, b) in enumerate(zip(g["idx"], O.f_to_idxBl(g["f"], cfg.patch_nums, vsd))):
        assert torch.equal(a, b), f"tokens differ at scale {si}"
    # idxBl_to_img / idxBl_to_h / img_to_recon (vqvae.py:77-104) from the golden's tokens
    sub = m["img_sub"]
    assert (O.idxBl_to_img_last(g["idx"], cfg.patch_nums, vsd)[:, :, ::sub, ::sub] - g["img_from_tokens_sub"]).abs().max().item() <= 1e-6
    for a, b in zip(g["h"], O.idxBl_to_var_input(g["idx"], cfg.patch_nums, vsd)):
        assert (a - b).abs().max().item() <= 1e-6
    recon = O.decoder_forward(O._conv(O.idxBl_to_fhat_list(g["idx"], cfg.patch_nums, vsd)[-1], vsd, "post_quant_conv", 1), vsd)
    assert (recon[:, :, ::sub, ::sub] - g["recon_sub"]).abs().max().item() <= 5e-5      # same f_hat reached by another op order


@pytest.mark.parametrize("name", [n for n in golden_names("cond") if "pn10" not in n])
def test_oracle_conditional_infer_matches_reference_golden(name):
    """ControlVAR.conditional_infer_cfg (control_var.py:223-354) with the golden's teacher-forced tokens."""
    g = load_golden(name)
    m, cfg = g["meta"], g["cfg"]
    sd = W.synthetic_var_state_dict(cfg, m["weight_seed"])
    vsd = W.synthetic_vae_state_dict(cfg, m["weight_seed"], with_encoder=False)
    out = O.conditional_infer_cfg(sd, vsd, cfg.patch_nums, cfg.depth, m["B"], torch.tensor(m["labels"]),
                                  torch.tensor(m["cond"]), m["cfg"], m["top_k"], m["top_p"],
                                  O.cpu_generator_noise(m["seed"]), c_mask=g["forced"] if m["c_mask"] else None,
                                  c_img=g["forced"] if m["c_img"] else None, more_smooth=bool(m.get("more_smooth", False)))
    for si, (a, b) in enumerate(zip(g["idx"], out["idx"])):
        assert torch.equal(a, b), f"tokens differ at scale {si}"
    sub = m["img_sub"]
    assert list(out["img"].shape) == m["img_shape"]
    assert (out["img"][:, :, ::sub, ::sub] - g["img_sub"]).abs().max().item() <= 1e-6
    assert (out["f_hat"][:m["B"]] - g["f_hat"]).abs().max().item() <= 1e-6


@pytest.mark.parametrize("name", golden_names("fwd"))
def test_oracle_forward_matches_reference_golden(name):
    """ControlVAR.forward (control_var.py:566-651): teacher-forced logits under the block-causal mask."""
    g = load_golden(name)
    m, cfg = g["meta"], g["cfg"]
    sd = W.synthetic_var_state_dict(cfg, m["weight_seed"])
    x = W.synthetic_teacher_input(cfg, m["B"], m["x_seed"])
    out = O.forward_teacher_forced(sd, cfg.patch_nums, cfg.depth, torch.tensor(m["labels"]), x, torch.tensor(m["cond"]),
                                   embed_dim=cfg.embed_dim, num_heads=cfg.heads, mask_first=bool(m.get("mask_first", True)))
    assert list(out.shape) == m["logits_shape"]
    assert (out[:, :, ::m["logits_sub"]] - g["logits_sub"]).abs().max().item() <= 2e-6


def test_multinomial_identity():
    """torch.multinomial(p, 1, replacement=True, generator=g) == argmax(p / Exp(1) drawn from g) (helpers.py:19)."""
    g1, g2 = torch.Generator(), torch.Generator()
    g1.manual_seed(5), g2.manual_seed(5)
    p = torch.rand(64, 4096).softmax(-1)
    a = torch.multinomial(p, 1, replacement=True, generator=g1)[:, 0]
    q = torch.empty(64, 4096).exponential_(1, generator=g2)
    assert torch.equal(a, O.multinomial1_with_noise(p, q))


def test_kv_cache_equals_teacher_forced_forward():
    """SURVEY.md section 4 identity: incremental KV-cached pass == one block-causal full-sequence pass."""
    from controlvar_b200.config import PathConfig
    cfg = PathConfig(depth=2, patch_nums=(1, 2, 3))
    sd = W.synthetic_var_state_dict(cfg, 0)
    C, H = cfg.C, cfg.num_heads
    torch.manual_seed(0)
    x = torch.randn(2, cfg.L, C)
    cond = torch.randn(2, C)
    full = x
    for bi in range(cfg.depth):
        full = O.adaln_block(full, cond, sd, f"blocks.{bi}.", H, None, False, cfg.attn_scale,
                             sd["attn_bias_for_masking"])
    caches = [dict() for _ in range(cfg.depth)]
    outs, cur = [], 0
    for n in cfg.scale_lens:
        xi = x[:, cur:cur + n]
        for bi in range(cfg.depth):
            xi = O.adaln_block(xi, cond, sd, f"blocks.{bi}.", H, caches[bi], False, cfg.attn_scale)
        outs.append(xi)
        cur += n
    assert (torch.cat(outs, 1) - full).abs().max().item() < 2e-5
