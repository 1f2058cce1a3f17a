"""GPU: tensor-core KV-cached attention (tcgen05, 3xTF32) against fp64 SDPA at the shapes of the real pyramid."""
import pytest
import torch
import torch.nn.functional as F

from controlvar_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("R,H,l,L", [(2, 2, 64, 64), (2, 3, 128, 310), (1, 2, 200, 510), (2, 2, 338, 848),
                                     (1, 4, 512, 1360), (3, 1, 72, 182), (1, 1, 130, 131)])
def test_attn_tc_matches_sdpa(R, H, l, L):
    torch.manual_seed(l + L)
    T = L + 5
    q = torch.randn(R, H, l, 64) * 2
    k = torch.randn(R, H, L, 64)
    v = torch.randn(R, H, L, 64)
    scale = 1 / 32
    ref = F.scaled_dot_product_attention(q.double(), k.double(), v.double(), scale=scale).transpose(1, 2).reshape(R, l, H * 64)
    kv = ops.KVCache(R, H, T, DEV)
    kg, vg = k.to(DEV), v.to(DEV)
    khi = kg.view(torch.int32).bitwise_and(-8192).view(torch.float32)
    vhi = vg.view(torch.int32).bitwise_and(-8192).view(torch.float32)
    kv.k_hi.view(R, H, kv.T, 64)[:, :, :L] = khi
    kv.k_lo.view(R, H, kv.T, 64)[:, :, :L] = kg - khi
    kv.vt_hi.view(R, H, 64, kv.T)[:, :, :, :L] = vhi.transpose(2, 3)
    kv.vt_lo.view(R, H, 64, kv.T)[:, :, :, :L] = (vg - vhi).transpose(2, 3)
    # poison the stale tail [L, T): it must be masked, not read into the result
    kv.k_hi.view(R, H, kv.T, 64)[:, :, L:] = 1e4
    kv.vt_hi.view(R, H, 64, kv.T)[:, :, :, L:] = -1e4
    res = {}
    for eng in (1, 0):
        out = torch.empty(R, l, H * 64, device=DEV)
        ops.attn_kvcache(q.to(DEV), kv, out, R, H, l, L, scale, engine=eng)
        torch.cuda.synchronize()
        res[eng] = (out.cpu().double() - ref).abs().max().item()
    print(f"\n[attn-accuracy] R={R} H={H} l={l} L={L}: tcgen05 err {res[1]:.3e}   SIMT err {res[0]:.3e}")
    assert res[1] < 2e-5 and res[0] < 2e-5
