# Which operand do the excess DRAM reads of the dense GEMM belong to?  CVAR_DEBUG_TRAFFIC=1: all tiles read the A rows of
# row tile 0 (A traffic ~ 0); =2: all tiles read the weight rows of column tile 0 (W traffic ~ 0); =3: both.
mkdir -p gpurun_out
for d in 0 1 2 3; do
  echo "=== CVAR_DEBUG_TRAFFIC=$d"
  CVAR_DEBUG_TRAFFIC=$d CVAR_GROUP_M=${G:-1} timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_ltcfabric.sum --clock-control none -k regex:tc_gemm2_kernel --csv --log-file gpurun_out/r02_attrib_$d.csv python tools/traffic_shapes.py > gpurun_out/r02_attrib_$d.txt 2>&1
  python - $d <<'PY'
import csv, sys
d_ = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(f"gpurun_out/r02_attrib_{d_}.csv") if not l.startswith("=="))]
h = rows[0]; ki, mi, vi = h.index("ID"), h.index("Metric Name"), h.index("Metric Value")
d = {}
for r in rows[1:]:
    if len(r) > vi: d.setdefault(int(r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
algs = [ln for ln in open(f"gpurun_out/r02_attrib_{d_}.txt") if "algorithmic" in ln]
for (i, m), ln in zip(sorted(d.items()), algs):
    rd, wr, t = m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0), m.get("gpu__time_duration.sum", 0)
    print(f"  {ln.split(':')[0]:34s} read {rd/1e9:6.3f} GB  write {wr/1e9:6.3f} GB  {t/1e6:.3f} ms  L2 hit {m.get('lts__t_sector_hit_rate.pct', 0):5.1f} %  "
          f"L2 read sectors from SMs {m.get('lts__t_sectors_srcunit_tex_op_read.sum', 0)*32/1e9:7.2f} GB  fabric sectors {m.get('lts__t_sectors_srcunit_ltcfabric.sum', 0)*32/1e9:7.2f} GB")
PY
done
