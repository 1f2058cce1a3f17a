"""CPU, world_size 2 over gloo: the multi-GPU host logic (batch slicing, one weight-arena broadcast, per-rank seeds)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_slices_cover_the_batch():
    from controlvar_b200.shard import shard_slice
    for total in (1, 7, 8, 64, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_slice(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_parameters_keeps_state_dict_and_values():
    from controlvar_b200 import VQVAE, build_control_var, weights as W
    from controlvar_b200.config import PathConfig
    from controlvar_b200.shard import pack_parameters
    pn = (1, 2, 3)
    cfg = PathConfig(depth=2, patch_nums=pn)
    vae = VQVAE(ch=160, v_patch_nums=pn)
    var = build_control_var(vae, depth=2, patch_nums=pn, mask_type="interleave_append", multi_cond=True)
    sd, vsd = W.synthetic_var_state_dict(cfg, 0), W.synthetic_vae_state_dict(cfg, 0)
    var.load_state_dict(sd), vae.load_state_dict(vsd)
    arena = pack_parameters([var, vae])
    assert arena.dim() == 1 and arena.numel() >= sum(v.numel() for v in sd.values() if v.dtype == torch.float32)
    for k, v in var.state_dict().items():
        assert torch.equal(v, sd[k]), k
        if v.dtype == torch.float32:
            assert v.untyped_storage().data_ptr() == arena.untyped_storage().data_ptr(), f"{k} is not a view of the arena"
            assert v.data_ptr() % 256 == arena.data_ptr() % 256, f"{k} lost its 256-byte alignment"
    assert list(vae.state_dict().keys()) == list(vsd.keys())


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from controlvar_b200 import VQVAE, build_control_var, weights as W
    from controlvar_b200.config import PathConfig
    from controlvar_b200 import shard
    pn = (1, 2)
    cfg = PathConfig(depth=1, patch_nums=pn)
    vae = VQVAE(ch=160, v_patch_nums=pn)
    var = build_control_var(vae, depth=1, patch_nums=pn, mask_type="interleave_append", multi_cond=True)
    if rank == 0:     # only rank 0 holds real weights before the broadcast
        var.load_state_dict(W.synthetic_var_state_dict(cfg, 0))
        vae.load_state_dict(W.synthetic_vae_state_dict(cfg, 0))
    arena = shard.pack_parameters([var, vae])
    # probe a real weight: the arena also carries attn_bias_for_masking (0 / -inf, control_var.py:168), identical on every
    # rank by construction, so a sum over the whole arena is inf everywhere and says nothing
    before = var.get_parameter("head.weight").abs().sum().item()
    shard.broadcast_weights(arena, src=0)
    ref = W.synthetic_var_state_dict(cfg, 0)
    ok = all(torch.equal(v, ref[k]) for k, v in var.state_dict().items())

    # sharded_infer's slicing / seeding, with the sampler replaced by a recorder (the kernels need a GPU)
    calls = []

    class Recorder:
        def autoregressive_infer_cfg(self, B, label_B, g_seed=None, cond_type=None, **kw):
            calls.append((B, label_B.tolist(), g_seed, cond_type.tolist()))
            return torch.full((B, 3, 4, 2), float(rank))

    labels, conds = torch.arange(5), torch.arange(5) % 4
    out = shard.sharded_infer(Recorder(), 5, labels, conds, g_seed=100, gather=True)
    q.put((rank, before, ok, calls, out[:, 0, 0, 0].tolist()))
    dist.destroy_process_group()


def test_two_ranks_broadcast_and_shard():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, before0, ok0, calls0, out0), (r1, before1, ok1, calls1, out1) = res
    assert before0 > 0 and before1 == 0            # rank 1 really started empty
    assert ok0 and ok1                             # ... and holds rank 0's weights after ONE broadcast
    assert calls0 == [(3, [0, 1, 2], 100, [0, 1, 2])]      # contiguous slices, seed = g_seed + rank
    assert calls1 == [(2, [3, 4], 101, [3, 0])]
    assert out0 == out1 == [0.0, 0.0, 0.0, 1.0, 1.0]       # uneven gather reassembles the batch in order
