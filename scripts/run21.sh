set -x
mkdir -p gpurun_out
for m in none nan; do
ncu --set full --clock-control none --import-source on -k regex:k_pairs_cols -s 3 -c 1 -o gpurun_out/cols_f32_$m -f python scripts/perf_k1.py --shape 64,137,256,512 --dtype float32 --axis 2 --mask $m > gpurun_out/ncu_cols_$m.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_pairs_cols -s 3 -c 1 -o gpurun_out/cols_f64_nan -f python scripts/perf_k1.py --shape 8,75,1080,1440 --dtype float64 --axis 2 --mask nan > gpurun_out/ncu_cols_f64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pairs_rows -s 3 -c 1 -o gpurun_out/rows_f32_nan_land -f python scripts/perf_k1.py --shape 256,721,1440 --dtype float32 --mask nan --land 0.3 > gpurun_out/ncu_rows_land.log 2>&1
ls -la gpurun_out
