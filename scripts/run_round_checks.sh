# round-end evidence: GPU tests, bench line, ncu launch list of the bench command, full captures of the query kernels
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; tail -c 1200 gpurun_out/bench_r01b.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r01b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cp_|pseudonormal" -s 5 -c 5 -o gpurun_out/cp_final python scripts/cp_bench_only.py 3 > /dev/null 2>&1
ls gpurun_out | tail -5
