mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_silu.log 2>&1; tail -3 gpurun_out/r02_gputests_silu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_silu.json 2> gpurun_out/r02_bench_silu.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_silu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in d['kernels'].items(): print(k, round(v['ms_per_step'],1), v.get('tflops'))
PY
