mkdir -p gpurun_out
bash scripts/prof_cols.sh > gpurun_out/plain_cols.log 2>&1 &&
ncu --target-processes all --set full --clock-control none --import-source on -k regex:"k_pairs_cols" -s 2 -c 1 -o gpurun_out/prof_cols3 -f bash scripts/prof_cols.sh > gpurun_out/ncu_cols.log 2>&1
grep K1 gpurun_out/plain_cols.log; ls -la gpurun_out/prof_cols3*
