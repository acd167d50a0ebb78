for cfg in "96 32" "160 48" "256 64" "320 96" "512 128" "1024 256"; do
set -- $cfg
FPOHM_K2_SEARCH=$1 FPOHM_K2_WALK=$2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cp_ --csv --log-file gpurun_out/cp_launch_b.csv python scripts/cp_bench_only.py 4 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/cp_launch_b.csv")) if len(r)>5]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value")
t=[round(float(r[v])/1e6,3) for r in rows[-3:]]
print("$cfg", t, round(sum(t),3))
PY
done
