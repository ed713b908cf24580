mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pairs|k_combine|k_mutual|k_bitround" -c 100 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
tail -3 gpurun_out/ncu_a.log | cut -c1-200
