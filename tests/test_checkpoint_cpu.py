"""CPU: checkpoint compatibility - the reference's wrapper / prefix handling and the
VAR -> ControlVAR surgery of load_var_weight (train_control_var_hpu.py:472-534), checked against a line-by-line
restatement of that function and, where /root/reference is mounted, against the reference's own state_dict key set."""
import os
import sys
from collections import OrderedDict

import pytest
import torch

from controlvar_b200 import VQVAE, build_control_var, checkpoint as CK, weights as W
from controlvar_b200.config import PathConfig

PN = (1, 2, 3, 4)


def _var(cfg):
    vae = VQVAE(ch=160, v_patch_nums=cfg.patch_nums)
    return vae, build_control_var(vae, depth=cfg.depth, patch_nums=cfg.patch_nums, mask_type="interleave_append",
                                  multi_cond=True)


def _plain_var_state_dict(cfg, seed=3):
    """A VAR-style checkpoint: ControlVAR's keys with a half-length positional table and first_l = 1 (VAR has one
    token per position: models/var.py in the reference; only the keys the surgery touches matter)."""
    sd = OrderedDict(W.synthetic_var_state_dict(cfg, seed))
    L1 = sum(pn * pn for pn in cfg.patch_nums)
    g = torch.Generator().manual_seed(seed)
    sd["pos_1LC"] = torch.randn(1, L1, cfg.C, generator=g)
    sd["pos_start"] = torch.randn(1, 1, cfg.C, generator=g)
    sd["lvl_1L"] = torch.zeros(1, L1, dtype=torch.long)
    sd["attn_bias_for_masking"] = torch.zeros(1, 1, L1, L1)
    del sd["cond_embed.weight"]                       # VAR has no condition-type embedding
    return sd


def test_unwrap_handles_wrapper_and_ddp_prefix(tmp_path):
    cfg = PathConfig(depth=2, patch_nums=PN)
    sd = W.synthetic_var_state_dict(cfg, 0)
    wrapped = {"model_state_dict": OrderedDict(("module." + k, v) for k, v in sd.items()), "step": 7, "epoch": 1}
    path = tmp_path / "checkpoint_step_latest.pth"
    torch.save(wrapped, path)
    out = CK.unwrap_state_dict(str(path))
    assert list(out.keys()) == list(sd.keys())
    vae, var = _var(cfg)
    res = CK.load_checkpoint(var, str(path), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(var.get_parameter("head.weight"), sd["head.weight"])
    # released VQVAE weights are a bare state_dict
    vsd = W.synthetic_vae_state_dict(cfg, 0)
    res = CK.load_checkpoint(vae, vsd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


@pytest.mark.parametrize("interpos", [False, True])
def test_load_var_weight_surgery(interpos):
    cfg = PathConfig(depth=2, patch_nums=PN)
    vae, var = _var(cfg)
    fresh_pos_start = var.pos_start.clone()
    ck = _plain_var_state_dict(cfg)
    res = CK.load_var_weight(var, {"model_state_dict": OrderedDict(("module." + k, v) for k, v in ck.items())},
                             interpos=interpos)
    # ControlVAR-only tensors stay at their constructor values and are reported missing (strict=False)
    assert set(res.missing_keys) == {"pos_start", "lvl_1L", "attn_bias_for_masking", "cond_embed.weight"}
    assert not res.unexpected_keys
    assert torch.equal(var.pos_start, fresh_pos_start)
    pos, got = ck["pos_1LC"], var.pos_1LC
    assert got.shape == (1, cfg.L, cfg.C)
    if not interpos:
        assert torch.equal(got, torch.cat([pos, pos], dim=1))                      # train_control_var_hpu.py:519
    else:
        L1 = L2 = 0
        for pn in PN:                                                             # train_control_var_hpu.py:494-503
            n = pn * pn
            assert torch.equal(got[:, L2:L2 + n], pos[:, L1:L1 + n])
            assert torch.equal(got[:, L2 + n:L2 + 2 * n], pos[:, L1:L1 + n])
            L1, L2 = L1 + n, L2 + 2 * n
    assert torch.equal(var.get_parameter("blocks.1.ffn.fc2.weight"), ck["blocks.1.ffn.fc2.weight"])
    assert torch.equal(var.lvl_1L, W.lvl_1L(cfg))                                  # rebuilt by the constructor


def test_load_var_weight_rejects_unimplemented_and_malformed():
    cfg = PathConfig(depth=2, patch_nums=PN)
    _, var = _var(cfg)
    with pytest.raises(NotImplementedError):
        CK.load_var_weight(var, _plain_var_state_dict(cfg), separator=True)
    bad = _plain_var_state_dict(cfg)
    del bad["lvl_1L"]
    with pytest.raises(KeyError):                      # the reference's `del var_state_dict[key]` raises the same
        CK.load_var_weight(var, bad)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference not mounted (GPU box)")
def test_key_set_equals_the_reference_modules():
    """Our modules expose exactly the reference's state_dict keys and shapes, so its checkpoints load strict=True."""
    import contextlib
    import io
    sys.path.insert(0, "/root/reference")
    try:
        from models import VQVAE as RefVQVAE, build_control_var as ref_build
    finally:
        sys.path.remove("/root/reference")
    cfg = PathConfig(depth=2, patch_nums=PN)
    with contextlib.redirect_stdout(io.StringIO()):
        rvae = RefVQVAE(vocab_size=4096, z_channels=32, ch=160, test_mode=True, share_quant_resi=4, v_patch_nums=PN)
        rvar = ref_build(rvae, depth=2, patch_nums=PN, mask_type="interleave_append", multi_cond=True)
    vae, var = _var(cfg)
    for ours, ref in ((var, rvar), (vae, rvae)):
        a = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
        b = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        assert a == b
    # a checkpoint written by the reference's save_checkpoint layout loads into ours
    res = CK.load_checkpoint(var, {"model_state_dict": rvar.state_dict()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    res = CK.load_checkpoint(vae, rvae.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
