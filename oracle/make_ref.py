"""Recipe: stage the UNMODIFIED reference under oracle/_ref/ so that it can travel to the GPU box.

The reference (lxa9867/ControlVAR) is a pure-Python package without a build step, so "building" it for the checker is a
byte-for-byte copy of the files its sampling path imports - models/*.py and dist.py - from where they lie under
/root/reference into oracle/_ref/ (git-ignored: reference sources never enter the history; NOT gpurun-ignored: the
directory ships with the snapshot like the built .so).  Nothing is edited; oracle/_ref/MANIFEST.json records the sha256 of
every file so that a run can state exactly what it timed.

Users: bench.py --impl reference / the cpu_baseline leg (the literal reference on the host cores, `kind: "reference"`),
tools/library_bar.py (the same modules on the B200: cuBLAS / cuDNN / SDPA - the library bar), tests that pin the oracle.
Product code never imports it (tests/test_abi.py checks).

    python oracle/make_ref.py        # in the build container; a no-op message when /root/reference is absent
"""
import hashlib
import json
import os
import shutil
import sys

SRC = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def make_ref(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "models")):
        if verbose:
            print(f"oracle/make_ref.py: {SRC} is not mounted here; keeping whatever oracle/_ref holds")
        return os.path.isdir(os.path.join(DST, "models"))
    os.makedirs(os.path.join(DST, "models"), exist_ok=True)
    manifest = {}
    files = [("dist.py", "dist.py")] + [(os.path.join("models", f), os.path.join("models", f))
                                        for f in sorted(os.listdir(os.path.join(SRC, "models"))) if f.endswith(".py")]
    for rel_src, rel_dst in files:
        s, d = os.path.join(SRC, rel_src), os.path.join(DST, rel_dst)
        shutil.copyfile(s, d)
        manifest[rel_dst] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files staged (unmodified)")
    return True


def ref_available() -> bool:
    return os.path.isfile(os.path.join(DST, "models", "control_var.py")) and os.path.isfile(os.path.join(DST, "dist.py"))


def import_reference():
    """-> (models module of the reference, its dist module).  Call with CUDA hidden for the CPU arm (dist.py binds the
    device string at import: 'cuda' whenever a GPU is visible, dist.py:11)."""
    if not ref_available():
        raise RuntimeError("oracle/_ref is empty: run `python oracle/make_ref.py` where /root/reference is mounted")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    sys.dont_write_bytecode = True
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        import dist as ref_dist          # noqa: F401  (the reference's dist.py, not torch.distributed)
        import models as ref_models
    return ref_models, ref_dist


def build_reference(depth: int, device: str, sd, vsd, patch_nums=(1, 2, 3, 4, 5, 6, 8, 10, 13, 16)):
    """The reference's VQVAE + ControlVAR (released configuration: interleave_append, multi_cond), our portable synthetic
    weights loaded with strict=True.  sd / vsd: state dicts (any device)."""
    import contextlib
    import io
    ref_models, _ = import_reference()
    with contextlib.redirect_stdout(io.StringIO()):      # the constructors print banners
        vae = ref_models.VQVAE(vocab_size=4096, z_channels=32, ch=160, test_mode=True, share_quant_resi=4,
                               v_patch_nums=patch_nums)
        var = ref_models.build_control_var(vae, depth=depth, patch_nums=patch_nums, mask_type="interleave_append",
                                           multi_cond=True)
    var.load_state_dict(sd, strict=True)
    missing = vae.load_state_dict(vsd, strict=False)      # the decode-only weight set has no encoder keys
    assert not [k for k in missing.missing_keys if not k.startswith(("encoder.", "quant_conv."))], missing.missing_keys
    vae.eval().to(device)
    var.eval().to(device)
    return vae, var


if __name__ == "__main__":
    ok = make_ref()
    sys.exit(0 if ok else 1)
