for l in 1 2; do for c in 262144 524288 1048576; do echo -n "lanes $l chunk $c: "; FPOHM_CP_LANES=$l FPOHM_CP_CHUNK=$c python scripts/e2e_chunks.py; done; done
