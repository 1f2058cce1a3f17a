"""CPU: the staged reference (oracle/_ref, oracle/make_ref.py) and the reference arm of bench.py.

oracle/_ref is test / bench infrastructure: unmodified copies of the reference's models/*.py and dist.py that travel to the
GPU box, where /root/reference does not exist.  Checked here: the copies are byte-identical to the mounted reference (when it
is mounted), the staged package reproduces the oracle bit for bit on a small model, and `bench.py --impl reference` prints
the contract's JSON line with kind "reference"."""
import hashlib
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_ref as R  # noqa: E402

needs_ref = pytest.mark.skipif(not R.ref_available(), reason="oracle/_ref not staged (run python oracle/make_ref.py)")


@needs_ref
def test_staged_files_are_unmodified_copies():
    man = json.load(open(os.path.join(R.DST, "MANIFEST.json")))
    assert "models/control_var.py" in man["files"] and "dist.py" in man["files"]
    for rel, sha in man["files"].items():
        assert hashlib.sha256(open(os.path.join(R.DST, rel), "rb").read()).hexdigest() == sha, rel
        src = os.path.join(R.SRC, rel)
        if os.path.exists(src):          # the build container: byte-for-byte the reference
            assert open(src, "rb").read() == open(os.path.join(R.DST, rel), "rb").read(), rel


@needs_ref
def test_staged_reference_equals_the_oracle_bit_for_bit():
    code = r"""
import sys, torch
sys.path.insert(0, %r)
from oracle import make_ref as R, controlvar_oracle as O
from controlvar_b200.config import PathConfig
from controlvar_b200 import weights as W
cfg = PathConfig(depth=2, patch_nums=(1, 2, 3, 4))
sd, vsd = W.synthetic_var_state_dict(cfg, 0), W.synthetic_vae_state_dict(cfg, 0, with_encoder=False)
vae, var = R.build_reference(2, 'cpu', sd, vsd, cfg.patch_nums)
lab, ct = torch.tensor([3, 77]), torch.tensor([1, 2])
with torch.no_grad():
    img = var.autoregressive_infer_cfg(2, lab, g_seed=0, cfg=1.5, top_k=900, top_p=0.96, cond_type=ct)
ref = O.autoregressive_infer_cfg(sd, vsd, cfg.patch_nums, 2, 2, lab, ct, 1.5, 900, 0.96, O.cpu_generator_noise(0))
assert torch.equal(img, ref['img']), (img - ref['img']).abs().max()
print('EQUAL')
""" % ROOT
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")      # dist.py binds the reference to 'cuda' when a GPU is visible
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "EQUAL" in out.stdout, out.stderr[-2000:]


@needs_ref
def test_bench_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "d12_b16",
                          "--steps", "1", "--warmup", "0", "--cpu-sample-batch", "1"], capture_output=True, text=True,
                         timeout=900, env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE")})
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == "d12_b16" and line["higher_is_better"] is True
