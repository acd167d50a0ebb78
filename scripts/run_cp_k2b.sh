for b in 5 6 7 8 10; do
FPOHM_K2_MINB=$b timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cp_ --csv --log-file gpurun_out/cp_launch_b.csv python scripts/cp_bench_only.py 4 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/cp_launch_b.csv")) if len(r)>5]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value")
t=[round(float(r[v])/1e6,3) for r in rows[-4:]]
print("minb $b", t, round(sum(t),3))
PY
done
