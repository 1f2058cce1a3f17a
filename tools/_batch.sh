mkdir -p gpurun_out
for w in d12_b16 d30_b32 d24_b8 d24_cond_b16; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r02_bench_${w}_final.json 2> gpurun_out/r02_bench_${w}_final.err
  python - $w <<'PY'
import json,sys
w=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02_bench_{w}_final.json').read().strip().splitlines()[-1])
print(w, round(d['value'],2), 'img/s', round(d['ms_per_step'],1), 'ms  e2e', round(d['e2e']['value'],2), 'clk', d['clocks']['sm_mhz'], {k: round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
PY
done
