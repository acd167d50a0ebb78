timeout 300 python scripts/cp_stats_packet.py | tail -6
FPOHM_CP_MODE=0 timeout 300 python scripts/cp_ab.py /tmp/a.npz > /dev/null && FPOHM_CP_MODE=1 timeout 300 python scripts/cp_ab.py /tmp/b.npz && python scripts/cp_ab.py --compare /tmp/a.npz /tmp/b.npz
FPOHM_CP_MODE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cp_ --csv --log-file gpurun_out/cp_launch.csv python scripts/cp_ab.py /tmp/b.npz > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/cp_launch.csv")) if len(r)>5]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value")
for j in (0,8,16,24,32):
  print([ (r[k].split("(")[0][-16:], round(float(r[v])/1e6,3)) for r in rows[1+4*j:5+4*j]])
PY
