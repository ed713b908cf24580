# second measurement + profiling pass of round 2 (gpurun -- bash scripts/profile_round2b.sh): profiles/r02b_*
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python -m pytest tests -m gpu -x -q) > gpurun_out/gputest.log 2>&1; tail -3 gpurun_out/gputest.log
python bench.py > gpurun_out/r02b_bench_full.log 2> gpurun_out/r02b_bench_full.err; tail -c 300 gpurun_out/r02b_bench_full.log
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02b_plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 60 --csv --log-file gpurun_out/r02b_launches_full.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02b_ncu_a.log 2>&1
python scripts/perf_k1.py --dtype float16 --mask nan --land 0.3 > gpurun_out/r02b_plain_c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_pairs_rows_tma" -s 4 -c 2 -o gpurun_out/prof_r02b_rows_f16_land -f python scripts/perf_k1.py --dtype float16 --mask nan --land 0.3 > gpurun_out/r02b_ncu_c.log 2>&1
python scripts/perf_k1.py --dtype float32 --mask nan --land 0.3 > gpurun_out/r02b_plain_d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_pairs_rows_tma" -s 4 -c 2 -o gpurun_out/prof_r02b_rows_f32_land -f python scripts/perf_k1.py --dtype float32 --mask nan --land 0.3 > gpurun_out/r02b_ncu_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pairs_chunked_walk2" -s 12 -c 2 -o gpurun_out/prof_r02b_chunk_walk2_lastax_f32 -f python scripts/perf_chunked.py --compare --chunk 744,103,180 > gpurun_out/r02b_ncu_e.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
ls -la gpurun_out | tail -12
