# the measurement + profiling pass behind profiles/r02_* (run on a B200 box: gpurun -- bash scripts/profile_round2.sh)
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_full.log 2> gpurun_out/r02_bench_full.err; tail -c 600 gpurun_out/r02_bench_full.log
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02_plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 60 --csv --log-file gpurun_out/r02_launches_full.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02_ncu_a.log 2>&1
python bench.py --time-steps 1095 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02_plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_pairs_rows_tma|k_finish" -s 6 -c 4 -o gpurun_out/prof_r02_rows_f32 -f python bench.py --time-steps 1095 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02_ncu_b.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-700
ls -la gpurun_out | tail -8
