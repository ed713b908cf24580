#!/bin/bash
# packed-type dirty path of the rows kernel (float16 land masks), short-run rule of the chunk walkers
mkdir -p gpurun_out
python -m pytest tests/test_counts_gpu.py tests/test_fused_gpu.py tests/test_chunked_gpu.py -x -q > gpurun_out/mask_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/mask_tests.log
timeout 100 python scripts/fuzz_k1.py --seconds 40 --seed 11 > gpurun_out/fuzz_k1.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/fuzz_k1.log
for dt in float16 int8; do
  python scripts/perf_k1.py --dtype $dt --mask nan --land 0.3 --check 2>&1 | grep -E "K1|fast-path|fast ==" 
done
python scripts/perf_k1.py --dtype float16 --mask nan 2>&1 | grep -E "K1|fast-path"
python scripts/perf_k1.py --dtype float16 --mask -999 --land 0.3 --check 2>&1 | grep -E "K1|fast-path|fast =="
python scripts/perf_k1.py --dtype float16 --mask nan --randmask 0.0005 --check 2>&1 | grep -E "K1|fast-path|fast =="
python scripts/perf_chunked.py --compare --chunk 744,32,32 2>&1 | tail -2
