set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_final.log 2>&1; tail -3 gpurun_out/r02_gputests_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 600 gpurun_out/r02_bench_final.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; tail -2 gpurun_out/r02_smoke_final.log
bash tools/traffic_sweep.sh 1 > gpurun_out/r02_traffic_final.log 2>&1; cat gpurun_out/r02_traffic_final.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 1 --warmup 0 --profile-only --no-graphs > gpurun_out/r02_launches_final.out 2>&1; tail -c 300 gpurun_out/r02_launches_final.out
