"""The library bar: the UNMODIFIED reference (oracle/_ref) run ON the B200
through PyTorch's library kernels - the bar each hand-written sm_100a kernel has to beat.

  arm fp32      as the reference ships: fp32 weights / activations; cuBLAS fp32 GEMMs (torch default: TF32 matmul OFF),
                cuDNN convolutions with TF32 ON (torch default), F.scaled_dot_product_attention fp32 (basic_var.py:117)
  arm bf16      the same call under torch.autocast(bfloat16): linear layers / convs in bf16, attention through
                flash_attn_func (basic_var.py:95,113) when flash_attn imports
Per arm: images/s of ControlVAR.autoregressive_infer_cfg (CUDA events, warm-up 2) and a torch.profiler breakdown of one call
into kernel classes (GEMM / attention / conv / other) by kernel name.  Writes one JSON object to stdout (and --out).

    python tools/library_bar.py [--depth 24] [--batch 64] [--steps 3] [--out profiles/r02_library_bar.json]
"""
import argparse
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CLASSES = [
    ("attention", re.compile(r"fmha|flash|attention|sdpa|softmax", re.I)),
    ("conv", re.compile(r"conv|cudnn|implicit|wgrad|dgrad|im2col|nhwc|nchw|xmma", re.I)),
    ("gemm", re.compile(r"gemm|cublas|cutlass|sgemm|s1688|s16816|gemv|splitk|nvjet|matmul", re.I)),
]


def classify(name: str) -> str:
    for cls, rx in CLASSES:
        if rx.search(name):
            return cls
    return "other"


def run_arm(var, B, label, ct, steps, autocast):
    import contextlib
    ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if autocast else contextlib.nullcontext

    def call(seed):
        with torch.no_grad(), ctx():
            return var.autoregressive_infer_cfg(B, label, g_seed=seed, cfg=1.5, top_k=900, top_p=0.96, cond_type=ct)

    for w in range(2):
        call(w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        img = call(100 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res = {"ms_per_call": ms, "images_per_s": B / (ms / 1e3), "img_mean": float(img.float().mean()),
           "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30}
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            call(7)
            torch.cuda.synchronize()
        per, top = {}, {}
        for ev in prof.key_averages():
            t = getattr(ev, "device_time_total", None)
            if t is None:
                t = getattr(ev, "cuda_time_total", 0.0)
            if t <= 0:
                continue
            c = classify(ev.key)
            per[c] = per.get(c, 0.0) + t / 1e3
            top[ev.key] = top.get(ev.key, 0.0) + t / 1e3
        res["class_ms"] = {k: round(v, 2) for k, v in sorted(per.items(), key=lambda kv: -kv[1])}
        res["top_kernels_ms"] = {k[:110]: round(v, 2) for k, v in sorted(top.items(), key=lambda kv: -kv[1])[:14]}
    except Exception as e:   # noqa: BLE001
        res["profiler_error"] = repr(e)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=24)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cond", type=int, default=1)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    assert torch.cuda.is_available()
    from controlvar_b200 import weights as W
    from controlvar_b200.config import PathConfig
    from oracle import make_ref as R
    cfgp = PathConfig(depth=args.depth)
    sd = W.synthetic_var_state_dict(cfgp, 0, device="cuda")
    vsd = W.synthetic_vae_state_dict(cfgp, 0, with_encoder=False, device="cuda")
    _, var = R.build_reference(args.depth, "cuda", sd, vsd)
    del sd, vsd
    B = args.batch
    label = (torch.arange(B) % 1000).cuda()
    ct = torch.full((B,), args.cond, dtype=torch.long).cuda()
    out = {"what": "unmodified reference (oracle/_ref) on the B200 through PyTorch library kernels", "depth": args.depth,
           "batch": B, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
           "allow_tf32_matmul": torch.backends.cuda.matmul.allow_tf32, "allow_tf32_cudnn": torch.backends.cudnn.allow_tf32}
    try:
        import flash_attn  # noqa: F401
        out["flash_attn"] = flash_attn.__version__
    except Exception as e:   # noqa: BLE001
        out["flash_attn"] = "unavailable: " + repr(e)
    for arm, autocast in (("fp32", False), ("bf16_autocast", True)):
        torch.cuda.reset_peak_memory_stats()
        try:
            out[arm] = run_arm(var, B, label, ct, args.steps, autocast)
        except Exception as e:   # noqa: BLE001
            out[arm] = {"error": repr(e)[:500]}
            torch.cuda.empty_cache()
        print(arm, json.dumps(out[arm])[:600], file=sys.stderr, flush=True)
    s = json.dumps(out, indent=1)
    print(s)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        open(args.out, "w").write(s)


if __name__ == "__main__":
    main()
