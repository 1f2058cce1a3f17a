"""Multi-GPU plumbing for the sampling path: samples never interact, so the batch is cut into contiguous slices, one
per rank, exactly like the reference's rank-sharded ``validate`` (train_control_var_hpu.py:366-378).  The only
collective is ONE NCCL broadcast of the packed weight arena at start-up; nothing is exchanged
inside the sampling loop."""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn


def pack_parameters(modules: Iterable[nn.Module]) -> torch.Tensor:
    """Move every floating-point parameter / buffer of ``modules`` into one contiguous fp32 arena (views keep the
    state_dict layout).  Returns the arena so it can be broadcast with a single collective."""
    modules = list(modules)          # iterated twice below: a generator would skip the cache invalidation
    tensors: List[torch.Tensor] = []
    seen = set()
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            if t.dtype == torch.float32 and id(t) not in seen:
                seen.add(id(t))
                tensors.append(t)
    if not tensors:
        raise ValueError("nothing to pack")
    dev = tensors[0].device
    # 256-byte aligned slots so every tensor keeps the alignment the 128-bit loads and TMA descriptors need
    offs, total = [], 0
    for t in tensors:
        offs.append(total)
        total += (t.numel() + 63) // 64 * 64
    arena = torch.zeros(total, dtype=torch.float32, device=dev)
    for t, o in zip(tensors, offs):
        view = arena[o:o + t.numel()].view(t.shape)
        view.copy_(t.data)
        t.data = view
    invalidate_caches(modules)
    return arena


def invalidate_caches(modules: Iterable[nn.Module]) -> None:
    """Drop everything DERIVED from the parameters (operand-form weights: TF32 / FP16-pair splits, repacked conv weights,
    the lvl_pos table, captured CUDA graphs): after the parameters were re-pointed (pack_parameters) or overwritten in
    place (broadcast_weights) they would otherwise keep the old values."""
    for m in modules:
        for name in ("_consts", "_packed", "_packed16", "_graphs"):
            c = getattr(m, name, None)
            if c is not None:
                c.clear()
        if hasattr(m, "_ws_gen"):
            m._ws_gen += 1


def broadcast_weights(arena: torch.Tensor, src: int = 0, modules: Iterable[nn.Module] = ()) -> None:
    """The path's only collective: one ncclBroadcast of the weight arena (a no-op for a single process).  modules: the
    modules whose parameters live in the arena - their derived caches are invalidated (a warm-up before the broadcast
    would otherwise leave non-source ranks decoding with pre-broadcast conv weights)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(arena, src=src)
    invalidate_caches(list(modules))


def shard_slice(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [begin, end) of ``total`` samples owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


@torch.no_grad()
def sharded_infer(var, B_total: int, label_B: torch.Tensor, cond_type: torch.Tensor, g_seed: int, gather: bool = False,
                  **kw):
    """Each rank samples its slice with seed g_seed + rank; optional all_gather of the images."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    b0, b1 = shard_slice(B_total, rank, world)
    img = var.autoregressive_infer_cfg(b1 - b0, label_B[b0:b1], g_seed=g_seed + rank, cond_type=cond_type[b0:b1], **kw)
    if not gather or world == 1:
        return img
    sizes = [shard_slice(B_total, r, world) for r in range(world)]
    outs = [torch.empty((e - b,) + tuple(img.shape[1:]), device=img.device, dtype=img.dtype) for b, e in sizes]
    dist.all_gather(outs, img) if len({e - b for b, e in sizes}) == 1 else _gather_uneven(outs, img, sizes, rank)
    return torch.cat(outs, 0)


def _gather_uneven(outs, img, sizes, rank):
    for r, (b, e) in enumerate(sizes):
        buf = img if r == rank else outs[r]
        dist.broadcast(buf, src=r)
        if r == rank:
            outs[r].copy_(img)
