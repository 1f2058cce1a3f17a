"""CPU: host logic of the evaluation driver (train_control_var_hpu.py:338-408) with stub models."""
import numpy as np
import pytest
import torch

from controlvar_b200 import validate as V


def test_class_slice_partitions_like_the_reference():
    for gpus in (1, 3, 8):
        got = [c for r in range(gpus) for c in V.class_slice(r, gpus)]
        assert got == list(range(1000))
    assert V.class_slice(7, 8) == list(range(875, 1000)) and len(V.class_slice(2, 3)) == 334     # last rank takes the rest


def test_batch_plan_and_uint8_rule():
    assert V.batch_plan(50, 16) == [16, 16, 16, 2]
    assert V.batch_plan(50, 25) == [25, 25, 0]                      # the empty third batch is skipped by the caller
    with pytest.raises(AssertionError):
        V.batch_plan(50, 50)                                        # reference: assert 50 > args.batch_size
    x = torch.tensor([0.0, 0.999 / 255, 1.0 / 255, 0.5, 254.999 / 255, 1.0]).view(1, 1, 1, 6).repeat(1, 3, 1, 1)
    ref = x.clone().permute(0, 2, 3, 1).mul_(255).cpu().numpy().astype(np.uint8)
    assert np.array_equal(V.to_uint8_hwc(x), ref) and ref[0, 0, :, 0].tolist() == [0, 0, 1, 127, 254, 255]


class _StubVAR:
    """Records the calls validate makes; returns an image whose value encodes (cls, seed)."""
    device = torch.device("cpu")
    patch_nums = (1, 2)

    def __init__(self):
        self.calls = []

    def autoregressive_infer_cfg(self, B, label_B, cond_type, cfg, top_k, top_p, g_seed):
        self.calls.append(("ar", B, int(label_B[0]), int(cond_type[0]), cfg, top_k, top_p, g_seed))
        return torch.full((B, 3, 64, 32), (g_seed % 200) / 255.0)

    def conditional_infer_cfg(self, B, label_B, cfg, top_k, top_p, g_seed, c_mask, c_img, cond_type):
        self.calls.append(("cond", B, int(label_B[0]), c_mask is not None, c_img is not None, tuple(cfg), g_seed))
        return torch.full((B, 3, 64, 32), 0.25)


class _StubVAE:
    def img_to_idxBl(self, img, v_patch_nums):
        assert float(img.min()) >= -1.0 and float(img.max()) <= 1.0      # (x - 0.5) / 0.5 of an image in [0, 1]
        return [torch.zeros(img.shape[0], pn * pn, dtype=torch.long) for pn in v_patch_nums]


def test_validate_classes_sequence_seeds_and_files(tmp_path):
    var, vae = _StubVAR(), _StubVAE()
    out = V.validate_classes(var, vae, str(tmp_path), rank=0, gpus=1, batch_size=2, guidance_scale=(4.0, 4.0, 4.0),
                             seed=42, per_class=5, classes=[3, 9], cond_type="depth")
    ar = [c for c in var.calls if c[0] == "ar"]
    assert [c[1] for c in ar] == [2, 2, 1, 2, 2, 1]                  # 5 // 2 full batches + the remainder, per class
    assert all(c[3] == 2 and c[4] == 4.0 for c in ar)                # 'depth' -> 2, cfg = guidance_scale[0]
    # seed = seed + i * (cls + 1), cumulative across batches AND classes, exactly as the reference writes it (:377)
    seeds, s = [], 42
    for cls in (3, 9):
        for i in range(3):
            s = s + i * (cls + 1)
            seeds.append(s)
    assert [c[7] for c in ar] == seeds
    from PIL import Image
    files = sorted((tmp_path / "cfg_4.0" / "3").iterdir(), key=lambda p: int(p.stem))
    assert [p.name for p in files] == ["0.png", "1.png", "2.png", "3.png", "4.png"]
    img = np.asarray(Image.open(files[0]))
    assert img.shape == (32, 32, 3) and np.array_equal(img, out[3][0][0])       # the image half, as returned


def test_gibbs_refinement_alternates_mask_and_image_forcing(tmp_path):
    var, vae = _StubVAR(), _StubVAE()
    V.validate_classes(var, vae, str(tmp_path), batch_size=4, per_class=5, classes=[1], gibbs=2, save_val=False)
    kinds = [(c[0], c[3], c[4]) if c[0] == "cond" else (c[0],) for c in var.calls]
    one_batch = [("ar",)] + [("cond", True, False), ("cond", False, True)] * 2
    assert kinds == one_batch + one_batch                                        # two batches (4 + 1)


def test_pixel_conditioned_branch(tmp_path):
    var, vae = _StubVAR(), _StubVAE()
    loader = [dict(image=torch.zeros(2, 3, 32, 32), mask=torch.zeros(2, 3, 32, 32), cls=torch.tensor([5, 6]),
                   type=torch.tensor([1, 1]))] * 2
    out = V.validate_pixel_conditioned(var, vae, loader, str(tmp_path), val_cond="canny", c_mask=True, guidance_scale=(6, 6, 6))
    assert len(out) == 2 and out[0].shape == (2, 32, 32, 3)
    assert all(c[0] == "cond" and c[3] and not c[4] and c[5] == (6, 6, 6) for c in var.calls)
    assert sorted(p.name for p in (tmp_path / "cfg_6_6_6_canny" / "0").iterdir()) == ["0.png", "1.png", "2.png", "3.png"]
