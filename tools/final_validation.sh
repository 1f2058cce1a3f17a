# One box, final build: full GPU suite, default bench (both arms), smoke, launch list of one call, fast-mode line.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_final.log 2>&1; tail -3 gpurun_out/r02_gputests_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; tail -2 gpurun_out/r02_smoke_final.log
timeout 900 ncu --nvtx --nvtx-include "bench_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 1 --warmup 0 --profile-only --no-cpu-baseline --no-graphs > gpurun_out/r02_launches_final.out 2>&1; tail -c 200 gpurun_out/r02_launches_final.out
timeout 600 python bench.py --fast --no-cpu-baseline > gpurun_out/r02_bench_fast_final.json 2> gpurun_out/r02_bench_fast_final.err; tail -c 300 gpurun_out/r02_bench_fast_final.json
