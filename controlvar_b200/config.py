"""Static description of a ControlVAR + VQVAE pair (shapes only, no tensors).

Mirrors the construction contract of the reference factories:
  build_control_var  -> /root/reference/models/__init__.py:21-45  (embed_dim = 64*depth, num_heads = depth)
  ControlVAR.__init__ -> /root/reference/models/control_var.py:24-67 (L, first_l, begin_ends, cos_attn forced at depth 30)
  VQVAE.__init__      -> /root/reference/models/vqvae.py:17-48 (ch_mult (1,1,2,2,4), 2 res blocks, vocab 4096, Cvae 32)
"""
from dataclasses import dataclass, field
from typing import Tuple

DEFAULT_PATCH_NUMS = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)


@dataclass(frozen=True)
class PathConfig:
    depth: int = 16
    patch_nums: Tuple[int, ...] = DEFAULT_PATCH_NUMS
    num_classes: int = 1000
    vocab_size: int = 4096
    Cvae: int = 32
    vae_ch: int = 160
    vae_ch_mult: Tuple[int, ...] = (1, 1, 2, 2, 4)
    vae_num_res_blocks: int = 2
    share_quant_resi: int = 4
    quant_resi: float = 0.5
    mlp_ratio: float = 4.0
    norm_eps: float = 1e-6
    tau: float = 4.0
    mask_factor: int = 2          # mask_type='interleave_append'
    multi_cond: bool = True
    head_dim: int = 64
    # build_control_var fixes embed_dim = 64*depth, num_heads = depth; the ControlVAR constructor itself
    # (control_var.py:24-32) takes them separately, which the parity goldens use for a narrow depth-30
    # (cosine-attention) model.  0 means "derive from depth".
    embed_dim: int = 0
    heads: int = 0

    @property
    def C(self) -> int:
        return self.embed_dim if self.embed_dim else self.head_dim * self.depth

    @property
    def num_heads(self) -> int:
        return self.heads if self.heads else self.depth

    @property
    def cos_attn(self) -> bool:
        # control_var.py:35 rewrites cos_attn to True iff depth == 30
        return self.depth == 30

    @property
    def scale_lens(self) -> Tuple[int, ...]:
        return tuple(self.mask_factor * pn * pn for pn in self.patch_nums)

    @property
    def L(self) -> int:
        return sum(self.scale_lens)

    @property
    def first_l(self) -> int:
        return self.scale_lens[0]

    @property
    def hidden(self) -> int:
        return round(self.C * self.mlp_ratio)

    @property
    def attn_scale(self) -> float:
        # basic_var.py:66-71
        return 1.0 if self.cos_attn else 1.0 / (self.head_dim ** 0.5) / self.tau

    @property
    def img_hw(self) -> int:
        return self.patch_nums[-1] * (2 ** (len(self.vae_ch_mult) - 1))

    def phi_index(self, si: int) -> int:
        """PhiPartiallyShared.__getitem__ (quant.py:282-293): argmin |ticks - si/(SN-1)|."""
        import numpy as np
        K = self.share_quant_resi
        SN = len(self.patch_nums)
        ticks = np.linspace(1 / 3 / K, 1 - 1 / 3 / K, K) if K == 4 else np.linspace(1 / 2 / K, 1 - 1 / 2 / K, K)
        return int(np.argmin(np.abs(ticks - si / (SN - 1))).item())
