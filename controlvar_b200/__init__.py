"""controlvar_b200 - B200-native (sm_100a) implementation of ControlVAR's next-scale sampling hot path.

Drop-in surface (same names / arguments as /root/reference/models/__init__.py:21-45):
    VQVAE, ControlVAR, build_control_var
The compute lives in libcvar_sm100.so (C ABI: include/cvar.h); importing the ops without the built library raises.
"""
from .config import PathConfig, DEFAULT_PATCH_NUMS
from .vqvae import VQVAE
from .control_var import ControlVAR


def build_control_var(
    vae: VQVAE, depth: int,
    patch_nums=DEFAULT_PATCH_NUMS,
    aln=1, aln_gamma_init=1e-3, shared_aln=False, layer_scale=-1,
    tau=4, cos_attn=False,
    flash_if_available=True, fused_if_available=True,
    mask_type='replace', cond_drop_rate=0.1, bidirectional=False, separate_decoding=False, separator=False,
    type_pos=False, indep=False, multi_cond=False,
):
    """Factory with the reference's signature and defaults (models/__init__.py:21-45)."""
    if mask_type == 'replace':
        mask_factor = 1
    elif mask_type == 'interleave_append':
        mask_factor = 2
    else:
        raise NotImplementedError
    return ControlVAR(
        vae_local=vae, patch_nums=patch_nums,
        depth=depth, embed_dim=depth * 64, num_heads=depth, drop_path_rate=0.1 * depth / 24,
        aln=aln, aln_gamma_init=aln_gamma_init, shared_aln=shared_aln, layer_scale=layer_scale,
        tau=tau, cos_attn=cos_attn, cond_drop_rate=cond_drop_rate,
        flash_if_available=flash_if_available, fused_if_available=fused_if_available, mask_factor=mask_factor,
        bidirectional=bidirectional, separate_decoding=separate_decoding, separator=separator, type_pos=type_pos,
        indep=indep, multi_cond=multi_cond,
    )


__all__ = ["PathConfig", "DEFAULT_PATCH_NUMS", "VQVAE", "ControlVAR", "build_control_var"]
