"""GPU: CUDA-graph capture / replay of the sampler (ControlVAR._sample).

The first call with a key runs eagerly, the second captures, later calls replay.  A replay must be BIT-identical to the
eager launch sequence for any seed / labels / condition types / forced tokens (they are inputs of the graph, not constants),
must keep counting its kernel launches, and must be dropped when a workspace it points into is reallocated."""
import pytest
import torch

from controlvar_b200 import VQVAE, build_control_var, ops, weights as W
from controlvar_b200.config import PathConfig

pytestmark = pytest.mark.gpu
DEV = "cuda"


def build(cfg):
    vae = VQVAE(ch=160, v_patch_nums=cfg.patch_nums)
    var = build_control_var(vae, depth=cfg.depth, patch_nums=cfg.patch_nums, mask_type="interleave_append", multi_cond=True)
    var.load_state_dict(W.synthetic_var_state_dict(cfg, 0))
    vae.load_state_dict(W.synthetic_vae_state_dict(cfg, 0))
    vae.to(DEV)
    var.to(DEV)
    return vae, var


KW = dict(cfg=1.5, top_k=900, top_p=0.96)


@pytest.mark.parametrize("rng_device", ["cuda", "cpu"])
def test_replay_equals_eager(rng_device):
    cfg = PathConfig(depth=2, patch_nums=(1, 2, 3, 4, 5))
    vae, var = build(cfg)
    var.rng_device = rng_device
    lab1, ct1 = torch.tensor([1, 2, 3]), torch.tensor([0, 1, 2])
    lab2, ct2 = torch.tensor([7, 500, 999]), torch.tensor([3, 3, 1])

    def eager(lab, ct, seed):
        var.use_graphs = False
        img = var.autoregressive_infer_cfg(3, lab, g_seed=seed, cond_type=ct, **KW)
        idx = [t.clone() for t in var.last_idx]
        var.use_graphs = True
        return img, idx

    ref_a, idx_a = eager(lab1, ct1, 5)
    ref_b, idx_b = eager(lab2, ct2, 6)
    assert not torch.equal(ref_a, ref_b)
    a1 = var.autoregressive_infer_cfg(3, lab1, g_seed=5, cond_type=ct1, **KW)      # eager (first call of the key)
    n0 = ops.launch_count()
    a2 = var.autoregressive_infer_cfg(3, lab1, g_seed=5, cond_type=ct1, **KW)      # capture + replay
    assert len(var._graphs) == 1 and next(iter(var._graphs.values()))["graph"] is not None
    n1 = ops.launch_count()
    a3 = var.autoregressive_infer_cfg(3, lab1, g_seed=5, cond_type=ct1, **KW)      # replay
    n2 = ops.launch_count()
    assert torch.equal(a1, ref_a) and torch.equal(a2, ref_a) and torch.equal(a3, ref_a)
    assert all(torch.equal(x, y) for x, y in zip(var.last_idx, idx_a))
    assert n2 - n1 > 100 and n1 - n0 >= n2 - n1          # a replay accounts for the kernels it launches
    # other inputs through the SAME graph
    b = var.autoregressive_infer_cfg(3, lab2, g_seed=6, cond_type=ct2, **KW)
    assert len(var._graphs) == 1
    assert torch.equal(b, ref_b) and all(torch.equal(x, y) for x, y in zip(var.last_idx, idx_b))
    # the returned image is the caller's: a later call must not overwrite it
    keep = b.clone()
    var.autoregressive_infer_cfg(3, lab1, g_seed=5, cond_type=ct1, **KW)
    assert torch.equal(b, keep)
    # a different guidance scale / top-k is a different key (they are baked into the captured launches)
    c1 = var.autoregressive_infer_cfg(3, lab1, g_seed=5, cond_type=ct1, cfg=3.0, top_k=50, top_p=0.0)
    c2 = var.autoregressive_infer_cfg(3, lab1, g_seed=5, cond_type=ct1, cfg=3.0, top_k=50, top_p=0.0)
    c3 = var.autoregressive_infer_cfg(3, lab1, g_seed=5, cond_type=ct1, cfg=3.0, top_k=50, top_p=0.0)
    assert len(var._graphs) == 2 and torch.equal(c1, c2) and torch.equal(c2, c3) and not torch.equal(c1, ref_a)


def test_graph_dropped_when_a_workspace_is_reallocated():
    cfg = PathConfig(depth=2, patch_nums=(1, 2, 3, 4))
    vae, var = build(cfg)
    lab, ct = torch.tensor([4, 5]), torch.tensor([1, 2])
    first = [var.autoregressive_infer_cfg(2, lab, g_seed=3, cond_type=ct, **KW) for _ in range(3)]
    assert torch.equal(first[0], first[2])
    big = torch.arange(6)
    var.autoregressive_infer_cfg(6, big, g_seed=3, cond_type=big % 4, **KW)        # grows every workspace
    again = [var.autoregressive_infer_cfg(2, lab, g_seed=3, cond_type=ct, **KW) for _ in range(3)]
    assert all(torch.equal(first[0], x) for x in again)
    var.release_workspace()
    assert not var._graphs
    assert torch.equal(first[0], var.autoregressive_infer_cfg(2, lab, g_seed=3, cond_type=ct, **KW))


def test_conditional_replay_equals_eager():
    cfg = PathConfig(depth=2, patch_nums=(1, 2, 3, 4))
    vae, var = build(cfg)
    B = 2
    g = torch.Generator().manual_seed(0)
    cm1 = [torch.randint(0, 4096, (B, pn * pn), generator=g).to(DEV) for pn in cfg.patch_nums]
    cm2 = [torch.randint(0, 4096, (B, pn * pn), generator=g).to(DEV) for pn in cfg.patch_nums]
    lab, ct = torch.tensor([10, 20]), torch.tensor([1, 3])
    kw = dict(cfg=(2.0, 1.5, 1.0), top_k=900, top_p=0.96, cond_type=ct)
    var.use_graphs = False
    r1 = var.conditional_infer_cfg(B, lab, g_seed=1, c_mask=cm1, **kw)
    r2 = var.conditional_infer_cfg(B, lab, g_seed=2, c_mask=cm2, **kw)
    var.use_graphs = True
    outs = [var.conditional_infer_cfg(B, lab, g_seed=1, c_mask=cm1, **kw) for _ in range(3)]
    assert all(torch.equal(o, r1) for o in outs)
    assert torch.equal(var.conditional_infer_cfg(B, lab, g_seed=2, c_mask=cm2, **kw), r2)
    assert not torch.equal(r1, r2)
