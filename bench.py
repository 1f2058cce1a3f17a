#!/usr/bin/env python
"""bench.py - images/sec of ControlVAR next-scale sampling (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload d24_b64|d12_b16|...]

A "step" = one ControlVAR.autoregressive_infer_cfg call over one synthetic batch: prologue, 10-scale KV-cached
transformer loop, CFG + top-k/top-p sampling, VQ steps and BOTH decoder passes (control map + image).
  value  - whole-job images/s with labels / condition types already resident in HBM, result left in HBM;
  e2e    - same call through the public API with HOST buffers: pinned-host -> device copy of labels / condition types and
           device -> host copy of the (B,3,512,256) fp32 images inside the timed region;
  roofline - dominant kernel class of the step (CUDA events around every launch of that class, same timed region);
  cpu_baseline - the UNMODIFIED reference package (staged under oracle/_ref by oracle/make_ref.py; kind "reference") run on
           this box's host cores with CUDA hidden, on a bounded sample of the same workload (a child process);
           the oracle port (kind "port") only when oracle/_ref is absent.
--impl reference times that CPU arm alone on all host threads (8 images per step, up to 2 warm-up steps).
Multi-GPU: one process per GPU (torchrun), batch sharded with B per GPU fixed (weak scaling), ONE NCCL broadcast of
the weight arena before timing, no collective inside the sampling loop; time = max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the metric is quoted on (d24, 256x256, CFG 1.5, canny), fits one GPU
    "d24_b64": dict(depth=24, B=64, cond=1, cfg=1.5, desc="d24 ControlVAR 256x256 10-scale, batch 64/GPU, CFG=1.5, canny condition"),
    # BASELINE.json configs[1]
    "d12_b16": dict(depth=12, B=16, cond=0, cfg=1.5, desc="d12 ControlVAR 256x256 10-scale, batch 16/GPU, CFG=1.5, mask condition"),
    # BASELINE.json configs[4] per-GPU shard (cosine attention)
    "d30_b32": dict(depth=30, B=32, cond=2, cfg=1.5, desc="d30 ControlVAR 256x256 10-scale, batch 32/GPU, CFG=1.5, depth condition"),
    "d24_b8": dict(depth=24, B=8, cond=1, cfg=1.5, desc="d24 ControlVAR 256x256 10-scale, batch 8/GPU (small-batch probe)"),
    # pixel-conditioned sampling (conditional_infer_cfg: 4 guidance replicas = 64 transformer rows,
    # control tokens teacher-forced from synthetic token maps, guidance (1.5, 1.5, 1.5))
    "d24_cond_b16": dict(depth=24, B=16, cond=1, cfg=1.5, conditional=True,
                         desc="d24 ControlVAR.conditional_infer_cfg 256x256, batch 16/GPU (64 replica rows), c_mask forced, canny"),
}
TOP_K, TOP_P = 900, 0.96          # reference validate() defaults, train_control_var_hpu.py:338
CPU_BATCH_NOTE = ("round 1 measured the oracle port on a 16-core B200 host at 0.336 / 0.342 / 0.347 / 0.391 img/s for 1 / 2 / 4 / 8 images per call - still rising with batch, so this bounded sample is a LOWER bound on large-batch CPU throughput")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, f"/tmp/cvar_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), smax.append(float(f[2])), power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def work_per_image(depth):
    """Algorithmic FLOP per generated sample, for the config block."""
    C, T = 64 * depth, 1360
    lens = [2 * p * p for p in (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)]
    sum_lL, L = 0, 0
    for l in lens:
        L += l
        sum_lL += l * L
    linear = 2 * depth * 24 * C * C * T
    attn = 2 * depth * 4 * sum_lL * C
    head = 2 * T * 2 * C * 4096
    return dict(linear_tflop=linear / 1e12, attn_tflop=attn / 1e12, head_tflop=head / 1e12, decoder_tflop=0.786)


# ------------------------------------------------------------------------------------------------ reference arm
def _ref_staged():
    return os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "models", "control_var.py"))


def run_cpu_reference(depth, B_sample, cond, cfg_scale, steps, warmup, threads=None):
    """Times the reference's own CPU implementation of the path on B_sample images per step, all host threads.
    kind "reference": the UNMODIFIED reference package staged under oracle/_ref by oracle/make_ref.py
    (ControlVAR.autoregressive_infer_cfg, models/control_var.py:356-565, fp32, CUDA hidden);
    kind "port": the oracle restatement (only when oracle/_ref is absent).  -> (times, threads, kind)"""
    # dist.py:11 binds the reference's device string to 'cuda' whenever a GPU is visible (its generator and helper tensors
    # follow it): the CPU arm must not see the GPU.  Effective because torch has not been imported in this process yet.
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import torch
    assert not torch.cuda.is_available(), "the CPU reference arm must run in a process that has not initialised CUDA"
    from controlvar_b200.config import PathConfig
    from controlvar_b200 import weights as W
    # all host cores, explicitly: torchrun exports OMP_NUM_THREADS=1, which would silently make the reference arm
    # single-threaded (and ~16x slower) in every N > 1 launch                                   BASELINE.md section 5
    torch.set_num_threads(threads or os.cpu_count())
    cfgp = PathConfig(depth=depth)
    sd = W.synthetic_var_state_dict(cfgp, 0)
    vsd = W.synthetic_vae_state_dict(cfgp, 0, with_encoder=False)
    label = torch.arange(B_sample) % 1000
    ct = torch.full((B_sample,), cond)
    if _ref_staged():
        from oracle import make_ref as R
        kind = "reference"
        _, var = R.build_reference(depth, "cpu", sd, vsd)
        del sd, vsd

        def call(it, B=B_sample):
            with torch.no_grad():
                return var.autoregressive_infer_cfg(B, label[:B], g_seed=it, cfg=cfg_scale, top_k=TOP_K, top_p=TOP_P,
                                                    cond_type=ct[:B])
    else:
        from oracle import controlvar_oracle as O
        kind = "port"

        def call(it, B=B_sample):
            return O.autoregressive_infer_cfg(sd, vsd, cfgp.patch_nums, depth, B, label[:B], ct[:B], cfg_scale, TOP_K, TOP_P,
                                              O.cpu_generator_noise(it), decode=True)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        call(it)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, torch.get_num_threads(), kind


def main_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B_s = args.cpu_sample_batch
    warm = min(args.warmup, 2)
    times, cores, kind = run_cpu_reference(wl["depth"], B_s, wl["cond"], wl["cfg"], args.steps, warm)
    ms = 1e3 * sum(times) / len(times)
    v = B_s / (ms / 1e3)
    what = ("the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py): ControlVAR.autoregressive_infer_cfg, fp32, "
            "CUDA hidden" if kind == "reference" else "the oracle port (oracle/_ref not staged)")
    sample = f"{B_s} image(s) per step of the same workload ({wl['desc']}), {warm} warm-up step(s); {what}; " + CPU_BATCH_NOTE
    line = {"impl": "reference", "metric": f"images/sec (256x256, d{wl['depth']}, CFG=1.5)",
            "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": args.workload, "desc": wl["desc"], "top_k": TOP_K, "top_p": TOP_P,
                                            "images_per_step": B_s},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(args):
    """The cpu_baseline leg of our arm: the reference arm above in a CHILD process (this one has CUDA initialised; the
    reference must not see the GPU), one step on the bounded sample.  -> the child's cpu_baseline object."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--workload", args.workload, "--cpu-sample-batch", str(args.cpu_sample_batch)]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS"):
        env.pop(k, None)
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"error": "no JSON line from the reference arm", "stderr": out.stderr[-400:]}
    except Exception as e:   # noqa: BLE001
        return {"error": repr(e)}


def library_bar():
    """The unmodified reference on the B200 through PyTorch's library kernels (tools/library_bar.py), from the committed
    capture - a constant of the round, not measured in this run."""
    p = os.path.join(ROOT, "profiles", "r02_library_bar.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    out = {"source": "profiles/r02_library_bar.json (tools/library_bar.py on a B200 of this pool; NOT measured in this run)",
           "workload": f"d{d['depth']} B={d['batch']}"}
    for arm in ("fp32", "bf16_autocast"):
        if arm in d and "ms_per_call" in d[arm]:
            out[arm] = {"images_per_s": d[arm]["images_per_s"], "ms_per_call": d[arm]["ms_per_call"],
                        "class_ms": d[arm].get("class_ms")}
    return out


def fast_vs_parity(var, ops, torch, Bc, label_d, ct_d, kw):
    """Fast mode against the parity mode of this library on the same inputs and noise (the parity mode is what the oracle
    tests pin): the fast run is teacher-forced onto the parity run's tokens, so every scale sees the same trajectory.
    -> token flip rate per scale, max |pixel| difference of the decoded images."""
    lab, ct = label_d[:Bc].contiguous(), ct_d[:Bc].contiguous()
    ops.set_fast_mode(False)
    img_p = var.autoregressive_infer_cfg(Bc, lab, g_seed=77, cond_type=ct, **kw)
    idx_p = [t.clone() for t in var.last_idx]
    ops.set_fast_mode(True)
    var.debug_forced_idx = idx_p
    img_f = var.autoregressive_infer_cfg(Bc, lab, g_seed=77, cond_type=ct, **kw)
    var.debug_forced_idx = None
    flips = [float((a != b).float().mean()) for a, b in zip(idx_p, var.last_idx)]
    return {"against": "the parity mode of this library, same inputs and Exp(1) noise, teacher-forced", "images": Bc,
            "token_flip_rate_per_scale": [round(f, 5) for f in flips],
            "token_flip_rate": sum(float((a != b).sum()) for a, b in zip(idx_p, var.last_idx)) / sum(a.numel() for a in idx_p),
            "max_abs_pixel_diff_same_tokens": float((img_p - img_f).abs().max())}


# ----------------------------------------------------------------------------------------------------- our arm
def main_ours(args, wl):
    import torch
    import torch.distributed as dist
    from controlvar_b200 import VQVAE, build_control_var, ops, weights as W
    from controlvar_b200.config import PathConfig
    from controlvar_b200 import shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly ONE JSON line (rank 0).  NCCL prints its version banner (and, under NCCL_DEBUG, its log) to
    # stdout when the first communicator comes up: while the process group and the weight broadcast are set up, file
    # descriptor 1 points at stderr.
    saved_stdout = None
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    ops.set_gemm_engine(args.engine)

    depth, B = wl["depth"], (args.batch or wl["B"])
    cfgp = PathConfig(depth=depth)
    vae = VQVAE(ch=160).to(dev)
    var = build_control_var(vae, depth=depth, mask_type="interleave_append", multi_cond=True).to(dev)
    if rank == 0:     # rank 0 materialises the synthetic weights, everyone else receives them over NVLink
        var.load_state_dict(W.synthetic_var_state_dict(cfgp, 0, device=dev))
        vae.load_state_dict(W.synthetic_vae_state_dict(cfgp, 0, device=dev))
    arena = shard.pack_parameters([var, vae])
    t_b0 = time.perf_counter()
    shard.broadcast_weights(arena, src=0, modules=[var, vae])
    torch.cuda.synchronize()
    bcast_ms = (time.perf_counter() - t_b0) * 1e3
    if saved_stdout is not None:
        dist.barrier()                      # every communicator the bench uses exists now
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)

    label_h = (torch.arange(B) + rank * B) % 1000
    ct_h = torch.full((B,), wl["cond"], dtype=torch.long)
    label_d, ct_d = label_h.to(dev), ct_h.to(dev)
    label_pin, ct_pin = label_h.pin_memory(), ct_h.pin_memory()
    img_host = torch.empty(B, 3, 512, 256, dtype=torch.float32).pin_memory()
    kw = dict(cfg=wl["cfg"], top_k=TOP_K, top_p=TOP_P)
    conditional = bool(wl.get("conditional"))
    c_mask_h = c_mask_d = c_mask_pin = None
    if conditional:       # synthetic control-token maps (what VQVAE.img_to_idxBl would return for a condition image)
        gtok = torch.Generator().manual_seed(1234 + rank)
        c_mask_h = [torch.randint(0, 4096, (B, pn * pn), generator=gtok) for pn in cfgp.patch_nums]
        c_mask_d = [t.to(dev) for t in c_mask_h]
        c_mask_pin = [t.pin_memory() for t in c_mask_h]
        kw["cfg"] = (wl["cfg"],) * 3
    if args.fast:
        ops.set_fast_mode(True)
    if args.no_graphs:
        var.use_graphs = False

    def call(lab, ct, it, cm):
        if conditional:
            return var.conditional_infer_cfg(B, lab, g_seed=it * world + rank, cond_type=ct, c_mask=cm, **kw)
        return var.autoregressive_infer_cfg(B, lab, g_seed=it * world + rank, cond_type=ct, **kw)

    def step_device(it):
        return call(label_d, ct_d, it, c_mask_d)

    def step_e2e(it):
        lab = label_pin.to(dev, non_blocking=True)
        ct = ct_pin.to(dev, non_blocking=True)
        cm = [t.to(dev, non_blocking=True) for t in c_mask_pin] if conditional else None
        img = call(lab, ct, it, cm)
        img_host.copy_(img, non_blocking=True)
        return img

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, first_it, profile=False):
        barrier()
        if profile:
            ops.profile_begin()
        n0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.cuda.nvtx.range_push("bench_step")     # lets `ncu --nvtx --nvtx-include "bench_step/"` scope a launch list
        for i in range(steps):
            fn(first_it + i)
        torch.cuda.nvtx.range_pop()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ops.launch_count() - n0
        prof = ops.profile_end() if profile else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, launches, prof

    for w in range(args.warmup):          # call 1 eager, call 2 captures the CUDA graph, calls 3+ replay it
        step_device(w)
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # pass 1 (roofline): eager, CUDA events around every launch of our kernels (the profiler switches the graph off)
    prof_steps = min(args.steps, 3)
    ms_prof, _, prof = timed(step_device, prof_steps, 500, profile=True)
    # pass 2 (value): inputs resident in HBM, the public call (CUDA-graph replay unless --no-graphs)
    ms_total, launches, _ = (ms_prof * args.steps / prof_steps, 0, None) if args.profile_only else timed(step_device, args.steps, 1000)
    # pass 3 (e2e): pinned host inputs copied in, images copied out, inside the timed region
    # --profile-only (diagnostic, for ncu launch lists): pass 1 only; the line is then not a bench line
    ms_e2e, _, _ = (ms_total, 0, None) if args.profile_only else timed(step_e2e, args.steps, 2000)
    clk = clocks.stop() if rank == 0 else None
    mem_gb = torch.cuda.max_memory_allocated() / 2**30
    fast_check = None
    if args.fast and rank == 0 and not conditional:
        fast_check = fast_vs_parity(var, ops, torch, min(B, 8), label_d, ct_d, kw)

    if rank == 0:
        pk = peaks()
        ms_step = ms_total / args.steps
        value = world * B / (ms_step / 1e3)
        e2e_v = world * B / (ms_e2e / args.steps / 1e3)
        # dominant kernel class by device time inside the timed region
        classes = {k: v for k, v in prof.items() if k in ("gemm", "conv", "attn")}
        dom = max(classes, key=lambda k: classes[k]["ms"])
        d = classes[dom]
        ach_tf = d["work"] / (d["ms"] / 1e3) / 1e12
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if not os.path.exists(tp):
            tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):        # ncu --set full figure of ONE representative launch (the class mixes many shapes)
            tj = json.load(open(tp))
            traffic = {"traffic": tj["traffic"], "note": f"NOT measured in this run: constant from {os.path.basename(tp)} = dram "
                                                         f"read+write of one launch ({tj['launch']}), algorithmic "
                                                         f"{tj['algorithmic_bytes']} B; {tj['source']}"}
        roof = {"kernel": {"gemm": "dense-layer GEMM (cvar_gemm / cvar_qkv_project)", "conv": "decoder implicit-GEMM conv (cvar_conv2d)",
                           "attn": "KV-cached attention (cvar_attn_kvcache)"}[dom],
                "bound": "tensor", "achieved": ach_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach_tf / pk["tf_sust"],
                "traffic": traffic.get("traffic") if dom == "gemm" else None, "traffic_note": traffic.get("note"),
                "peak_source": pk["src"] + ", sustained bf16 (kernel timed inside a long step)",
                "launches": d["launches"], "avg_launch_ms": d["ms"] / d["launches"],
                "share_of_step": d["ms"] / ms_prof,
                "timed_in": f"pass 1 of this run: {prof_steps} eager step(s) with CUDA events around every launch of this "
                            f"library ({ms_prof / prof_steps:.1f} ms/step; the value / e2e passes replay a CUDA graph)"}
        kernels = {}
        for k, v in prof.items():
            e = {"launches": v["launches"], "ms_per_step": v["ms"] / prof_steps, "share_of_step": v["ms"] / ms_prof}
            if v["work"] > 0:
                e["tflops"] = v["work"] / (v["ms"] / 1e3) / 1e12
                e["frac_of_bf16_sustained"] = e["tflops"] / pk["tf_sust"]
            if v["bytes"] > 0:
                e["algorithmic_gbs"] = v["bytes"] / (v["ms"] / 1e3) / 1e9
                e["frac_of_hbm"] = e["algorithmic_gbs"] / pk["hbm"]
            kernels[k] = e
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_baseline_subprocess(args)
        engine_name = {0: "simt-fp32", 1: "tcgen05-3xtf32 (1 CTA per tile)",
                       3: "tcgen05-3xtf32 (2-CTA all-TMA dense layers, 1-CTA convs)",
                       4: "tcgen05 f16x3 (FP16 pairs): 2-CTA all-TMA dense layers and decoder convs (K-split for the long "
                          "accumulations), kind::f16 attention on an FP16-pair KV cache"}.get(ops.get_gemm_engine(), "?")
        line = {"metric": f"images/sec (256x256, d{depth}, CFG=1.5)", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 operands, one MMA per product, f32 accumulate (FAST MODE - not the parity mode)" if args.fast else "f32",
                "data": "synthetic",
                "config": {"workload": args.workload, "desc": wl["desc"], "batch_per_gpu": B, "global_batch": world * B,
                           "top_k": TOP_K, "top_p": TOP_P, "gemm_engine": engine_name,
                           "cuda_graph": bool(var.use_graphs), "mode": "fast (cvar_set_fast_mode: hi halves only)" if args.fast else "parity (f16x3, fp32-class)",
                           "l2": "working set >> L2: weights %.1f GB + KV arena %.1f GB streamed every step" % (
                               arena.numel() * 4 / 1e9, 2 * depth * 2 * B * 1360 * 64 * depth * 4 / 1e9),
                           "parallelism": f"dp{world} (batch sharded, one NCCL weight broadcast: {bcast_ms:.1f} ms)",
                           "algorithmic_tflop_per_image": work_per_image(depth), "peak_mem_gib": round(mem_gb, 1)},
                "e2e": {"value": e2e_v, "unit": "images/s", "h2d_bytes_per_step": int(16 * B * world),
                        "d2h_bytes_per_step": int(B * 3 * 512 * 256 * 4 * world)},
                "gpu_launches": launches, "roofline": roof, "kernels": kernels, "cpu_baseline": cpu, "clocks": clk,
                "library_bar": library_bar()}
        if args.fast:
            line["fast_mode_check"] = fast_check
            line["note"] = ("FAST MODE line: a separate, clearly labelled measurement (SURVEY.md 7.2); tokens diverge from the "
                            "reference by construction - no parity claim attaches to it")
        if args.profile_only:
            line["e2e"] = None
            line["note"] = "--profile-only run: diagnostic, not a bench line"
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="d24_b64", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch of the workload")
    ap.add_argument("--engine", type=int, default=int(os.environ.get("CVAR_GEMM_ENGINE", "4")),
                    help="0 SIMT fp32, 1 tcgen05 3xTF32 (1 CTA per tile), 3 = 1 plus the 2-CTA all-TMA kernel for dense layers, "
                         "4 = 3 with FP16-pair (f16x3) dense layers")
    ap.add_argument("--cpu-sample-batch", type=int, default=8,
                    help="images per CPU-reference step (bounded sample: ~20 s of CPU work at d24 on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="diagnostic: device pass only (for ncu launch lists)")
    ap.add_argument("--fast", action="store_true",
                    help="FAST MODE (not parity): single-MMA fp16 operands; prints a separately labelled line")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel from the host instead of replaying a CUDA graph")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        main_reference(args, wl)
    else:
        main_ours(args, wl)
