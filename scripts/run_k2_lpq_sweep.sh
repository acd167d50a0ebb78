# search-kernel density sweep: lanes kept per listed query (x16); 16 = one lane per query, 4 = four queries per lane
for v in 64 16 8 5 4 3 2; do echo -n "LPQ=$v "; FPOHM_K2_LPQ=$v python scripts/cp_ab.py /tmp/b_$v.npz 2>&1 | grep "bench\|hexverts" | tr '\n' ' '; echo; done
FPOHM_CP_MODE=0 python scripts/cp_ab.py /tmp/a.npz > /dev/null; for v in 4 2; do python scripts/cp_ab.py --compare /tmp/a.npz /tmp/b_$v.npz; done
