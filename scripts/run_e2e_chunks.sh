for l in 2 3 4; do for c in 131072 262144 524288; do echo -n "lanes $l chunk $c: "; FPOHM_CP_LANES=$l FPOHM_CP_CHUNK=$c python scripts/e2e_chunks.py; done; done
