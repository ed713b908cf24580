# the measurement + profiling pass behind profiles/ (run on a B200 box: gpurun -- bash scripts/profile_round.sh)
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; tail -c 3500 gpurun_out/bench_full.log
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pairs|k_combine|k_mutual|k_bitround|k_block|k_byte|k_bit" -c 100 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
python bench.py --time-steps 1024 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pairs_rows_tma -s 3 -c 2 -o gpurun_out/prof_rows_f32_final -f python bench.py --time-steps 1024 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pairs_rows_tma -s 3 -c 1 -o gpurun_out/prof_rows_f32_nanmask -f python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pairs_rows_tma -s 3 -c 1 -o gpurun_out/prof_rows_f32_nanmask_land -f python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan --land 0.3 > gpurun_out/ncu_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pairs_cols -s 3 -c 1 -o gpurun_out/prof_cols_f64_nanmask_land -f python scripts/perf_k1.py --shape 8,75,1080,1440 --dtype float64 --axis 2 --mask nan --land 0.3 > gpurun_out/ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pairs_cols -s 3 -c 1 -o gpurun_out/prof_cols_f32 -f python scripts/perf_k1.py --shape 64,137,256,512 --dtype float32 --axis 1 --mask none > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none -k regex:"k_byte_shuffle|k_bit_shuffle" -s 6 -c 4 -o gpurun_out/prof_k4 -f python scripts/perf_k4.py 2e9 > gpurun_out/ncu_g.log 2>&1
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --bitround | tail -1
python scripts/perf_k4.py 4e9
python scripts/perf_k1.py --shape 20,75,1080,1440 --dtype float64 --axis 2 --mask nan --land 0.3
python scripts/perf_k1.py --shape 20,75,1080,1440 --dtype float64 --axis 2 --mask none
python scripts/perf_k1.py --shape 72,137,256,512 --dtype float32 --axis 0 --mask none
python scripts/perf_k1.py --shape 72,137,256,512 --dtype float32 --axis 1 --mask none
python scripts/perf_k1.py --shape 72,137,256,512 --dtype float32 --axis 2 --mask none
python scripts/perf_k1.py --shape 72,137,256,512 --dtype float32 --axis 3 --mask none
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float64 --mask none
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float64 --mask nan
python scripts/perf_k1.py --shape 2048,721,1440 --dtype float16 --mask none
python scripts/perf_k1.py --shape 2048,721,1440 --dtype float16 --mask nan
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-900
ls -la gpurun_out | tail -12
