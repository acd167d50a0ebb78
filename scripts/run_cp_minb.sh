for b in 5 6 7 8; do
FPOHM_K1_MINB=$b timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cp_ --csv --log-file gpurun_out/cp_launch_$b.csv python scripts/cp_bench_only.py 4 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/cp_launch_$b.csv")) if len(r)>5]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value")
print($b, [ (r[k].split("(")[0][-16:], round(float(r[v])/1e6,3)) for r in rows[-3:]])
PY
done
