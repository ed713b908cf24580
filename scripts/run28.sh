for m in none nan; do
XBI_COLS_TMA=1 ncu --set full --clock-control none --import-source on -k regex:k_pairs_cols -s 3 -c 1 -o gpurun_out/colstma_f32_$m -f python scripts/perf_k1.py --shape 512,721,1440 --dtype float32 --axis 1 --mask $m > gpurun_out/ncu_colstma_$m.log 2>&1
done
