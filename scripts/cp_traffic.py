"""DRAM traffic of one C4 step from an `ncu --set full` capture of the step's kernels (profiles/r02_closest_point_traffic.json).
usage: python scripts/cp_traffic.py report.ncu-rep > profiles/r02_closest_point_traffic.json"""
import csv, json, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
ik, ir, iw, it = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
ur, uw = rows[1][ir], rows[1][iw]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = {}
tot = 0.0
for r in rows[2:]:
    b = float(r[ir]) * scale[ur] + float(r[iw]) * scale[uw]
    k = r[ik].split("(")[0].split("::")[-1]
    d = per.setdefault(k, {"launches": 0, "dram_bytes": 0.0, "ms": 0.0})
    d["launches"] += 1; d["dram_bytes"] += b; d["ms"] += float(r[it])
    tot += b
print(json.dumps({"source": rep, "dram_bytes_per_step": tot, "algorithmic_bytes_per_step": 84 * 11857634,
                  "ratio": tot / (84 * 11857634), "per_kernel": per, "time_unit": rows[1][it]}, indent=1))
