#!/bin/bash
# rows kernel: 32-bit shared atomics for the boundary corrections; ncu of the two-stream chunk walkers
mkdir -p gpurun_out
python -m pytest tests/test_counts_gpu.py tests/test_fused_gpu.py -x -q > gpurun_out/mask_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/mask_tests.log
for dt in float32 float64 float16; do
  python scripts/perf_k1.py --dtype $dt --mask nan --land 0.3 --check 2>&1 | grep -E "K1|fast-path|fast ==" 
  python scripts/perf_k1.py --dtype $dt --mask nan 2>&1 | grep -E "K1|fast-path"
done
python scripts/perf_k1.py --dtype float32 --mask -999 --land 0.3 --check 2>&1 | grep -E "K1|fast-path|fast =="
python scripts/perf_chunked.py --compare --chunk 744,103,180 > gpurun_out/pc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_pairs_chunked_walk2" -s 4 -c 2 -o gpurun_out/prof_r02_chunk_walk2_f32 -f python scripts/perf_chunked.py --compare --chunk 744,103,180 > gpurun_out/ncu_walk2.log 2>&1
ls -la gpurun_out | tail -4
