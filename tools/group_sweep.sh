mkdir -p gpurun_out
for g in 1 2 4 8; do
  CVAR_GROUP_M=$g timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:tc_gemm2_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_fc1_group$g.csv python tools/prof_kernels.py gemm > /dev/null 2>&1
  echo "group $g"; tail -4 gpurun_out/r02_fc1_group$g.csv | awk -F'","' '{print "  " $(NF-2), $NF}'
  CVAR_GROUP_M=$g timeout 200 python tools/gemm_ab.py 2>&1 | grep "overlap=1" | cut -c1-100
done
