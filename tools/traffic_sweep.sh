# DRAM traffic of the five dense-layer shapes for several rasterisation groups (CVAR_GROUP_M > 0: row groups, < 0: column groups)
mkdir -p gpurun_out
for g in "$@"; do
  CVAR_GROUP_M=$g timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tc_gemm2_kernel --csv --log-file gpurun_out/r02_traffic_g$g.csv python tools/traffic_shapes.py > gpurun_out/r02_traffic_g$g.txt 2>&1
  echo "group $g"; python - "$g" <<'PY'
import csv, sys
g = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(f"gpurun_out/r02_traffic_g{g}.csv") if not l.startswith("=="))]
h = rows[0]; ki, mi, vi = h.index("ID"), h.index("Metric Name"), h.index("Metric Value")
d = {}
for r in rows[1:]:
    if len(r) > vi: d.setdefault(int(r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
algs = [ln for ln in open(f"gpurun_out/r02_traffic_g{g}.txt") if "algorithmic" in ln]
for (i, m), ln in zip(sorted(d.items()), algs):
    alg = int(ln.split("algorithmic")[1].split()[0])
    rd, wr, t = m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0), m.get("gpu__time_duration.sum", 0)
    print(f"  {ln.split(':')[0]:34s} read {rd/1e9:6.3f} GB  write {wr/1e9:6.3f} GB  = {(rd+wr)/alg:5.2f} x algorithmic  ({t/1e6:.3f} ms under ncu)")
PY
done
