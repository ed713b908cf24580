mkdir -p gpurun_out
python scripts/perf_k1.py --shape 72,137,256,512 --dtype float32 --axis 1 > gpurun_out/plain_c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -s 40 -c 12 --csv --log-file gpurun_out/launches_cols.csv python scripts/perf_k1.py --shape 72,137,256,512 --dtype float32 --axis 1 > gpurun_out/ncu_c.log 2>&1
