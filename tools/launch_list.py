"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of ONE bench step by kernel class.
usage: python tools/launch_list.py gpurun_out/launches.csv  ->  markdown table on stdout
The rows between the 1st and the 2nd prologue_kernel are one autoregressive_infer_cfg call; when the capture holds a
single call (bench.py --steps 1 --warmup 0 --profile-only) everything after the 1st prologue_kernel is used."""
import csv
import re
import sys

CLASSES = [
    (r"tc_gemm2_kernel<.*DenseRow", "dense layers, tcgen05 f16x3 2-CTA, row epilogue (tc_gemm2_kernel<DenseRow<mode>, F16>)"),
    (r"tc_gemm2_kernel<.*QkvRow", "QKV projection + KV append, tcgen05 f16x3 2-CTA, row epilogue (tc_gemm2_kernel<QkvRow, F16>)"),
    (r"tc_conv2_kernel<.*ConvRow", "decoder conv, tcgen05 f16x3 2-CTA implicit GEMM by 4-D TMA, row epilogue (tc_conv2_kernel<ConvRow>)"),
    (r"tc_gemm2_kernel<.*DenseEpilogue, 1>", "dense layers, tcgen05 f16x3 2-CTA (tc_gemm2_kernel<DenseEpilogue, F16>)"),
    (r"tc_gemm2_kernel<.*QkvEpilogue, 1>", "QKV projection + KV append, tcgen05 f16x3 2-CTA (tc_gemm2_kernel<QkvEpilogue, F16>)"),
    (r"tc_gemm2_kernel<.*DenseEpilogue", "dense layers, tcgen05 3xTF32 2-CTA"),
    (r"tc_gemm2_kernel<.*QkvEpilogue", "QKV projection + KV append, tcgen05 3xTF32 2-CTA"),
    (r"tc_conv2_kernel", "decoder conv, tcgen05 f16x3 2-CTA implicit GEMM by 4-D TMA (tc_conv2_kernel)"),
    (r"tc_gemm_kernel<.*ConvALoader", "decoder conv, tcgen05 3xTF32 1-CTA (tc_gemm_kernel<ConvALoader>)"),
    (r"tc_gemm_kernel<", "dense layers, tcgen05 3xTF32 1-CTA (tc_gemm_kernel)"),
    (r"attn16_tc_kernel", "attention, tcgen05 kind::f16 on FP16 pairs (attn16_tc_kernel, l >= 64)"),
    (r"attn16_simt_kernel", "attention, SIMT on FP16 pairs (attn16_simt_kernel, l < 64)"),
    (r"attn_tc_kernel", "attention, tcgen05 (attn_tc_kernel, l >= 64)"),
    (r"attn_kvcache_kernel", "attention, SIMT (attn_kvcache_kernel, l < 64)"),
    (r"sgemm_kernel", "SIMT fp32 GEMM / conv (small, ragged or accuracy-pinned layers)"),
    (r"ln_modulate_stream_kernel", "ln_modulate_stream_kernel (large scales: persistent, rows by bulk copy, modulation vectors in shared memory)"),
    (r"ln_modulate_kernel", "ln_modulate_kernel (small scales: one warp per row)"),
    (r"conv3x3_small_cout", "Decoder.conv_out, 3 output channels (conv3x3_small_cout_*_kernel, SIMT fp32)"),
    (r"affine_nc_kernel", "affine_nc_kernel (GroupNorm + SiLU, writes the FP16 pair)"),
    (r"upsample2x_split_kernel", "upsample2x_split_kernel"),
    (r"split_f16_kernel", "split_f16_kernel"),
    (r"split_tf32_kernel|tc_split", "TF32 split (one-time weight packing)"),
    (r"gn_partial_kernel", "gn_partial_kernel"),
    (r"gn_finalize_kernel", "gn_finalize_kernel"),
    (r"cfg_sample_kernel", "cfg_sample_kernel"),
    (r"vq_step", "vq_step_kernel"),
    (r"cos_attn_normalize", "cos_attn_normalize_kernel"),
    (r"softmax_rows_kernel", "softmax_rows_kernel"),
    (r"nchw_to_nhwc_kernel", "nchw_to_nhwc_kernel"),
    (r"prologue_kernel", "prologue_kernel"),
    (r"lvl_pos_kernel", "lvl_pos_kernel"),
    (r"repack_conv_weight", "repack_conv_weight_kernel"),
    (r"exponential", "[torch] exponential_ (Exp(1) noise for multinomial)"),
]


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    header = next(rd)
    ki, vi, ui = header.index("Kernel Name"), header.index("Metric Value"), header.index("Metric Unit")
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        rows.append((r[ki], ms))
    pro = [i for i, (k, _) in enumerate(rows) if "prologue_kernel" in k]
    if not pro:
        sys.exit("no prologue_kernel in the capture")
    lo, hi = pro[0], (pro[1] if len(pro) > 1 else len(rows))
    call, before = rows[lo:hi], rows[:lo]
    agg = {}
    for k, ms in call:
        name = next((label for pat, label in CLASSES if re.search(pat, k)), "[other] " + k[:60])
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ms
        a[1] += 1
    total = sum(a[0] for a in agg.values())
    print(f"Kernel time of the call: {total:.1f} ms over {sum(a[1] for a in agg.values())} launches "
          f"({len(before)} launches before the first prologue_kernel - one-time weight packing - excluded).\n")
    print("| share | ms | launches | kernel class |\n|---|---|---|---|")
    for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"| {100 * ms / total:5.1f} % | {ms:8.2f} | {n:4d} | {name} |")
    n_s = sum(a[1] for k, a in agg.items() if k.startswith("cfg_sample"))
    n_v = sum(a[1] for k, a in agg.items() if k.startswith("vq_step"))
    n_ln = sum(a[1] for k, a in agg.items() if k.startswith("ln_modulate"))
    print(f"\ncounts: cfg_sample {n_s}, vq_step {n_v}, ln_modulate {n_ln} (one d24 call: 10, 10, 490)")


if __name__ == "__main__":
    main(sys.argv[1])
