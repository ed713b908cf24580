set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; tail -c 2600 gpurun_out/bench_full.log
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pairs|k_combine|k_mutual|k_bitround|k_block" -c 100 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
python bench.py --time-steps 1024 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pairs_rows_tma -s 3 -c 2 -o gpurun_out/prof_rows_f32_final -f python bench.py --time-steps 1024 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out | tail -5
