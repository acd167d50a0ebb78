for ch in 1 2 4 8; do for c in 8 16 32; do
echo -n "FILL_CH=$ch CTAS=$c  "; FPOHM_FILL_CH=$ch FPOHM_FILL_CTAS=$c python scripts/bench_kernels.py voxel 2>&1 | grep '"ms"' | tr -d '\n'; echo
done; done
