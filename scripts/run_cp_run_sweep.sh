# A/B of the seeded consecutive-packet runs: bit comparison against the per-lane igl-order kernel, then a sweep of the run length
FPOHM_CP_MODE=0 timeout 300 python scripts/cp_ab.py /tmp/a.npz | tail -5
for r in 1 4 8 16 32; do
  echo "== FPOHM_CP_RUN=$r"
  FPOHM_CP_RUN=$r FPOHM_CP_MODE=1 timeout 300 python scripts/cp_ab.py /tmp/b.npz && python scripts/cp_ab.py --compare /tmp/a.npz /tmp/b.npz
done
