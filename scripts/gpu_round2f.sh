#!/bin/bash
# float32 NaN probe on three-operand packed FMAs
mkdir -p gpurun_out
python -m pytest tests/test_counts_gpu.py tests/test_fused_gpu.py tests/test_chunked_gpu.py tests/test_get_bitinformation_gpu.py -x -q > gpurun_out/mask_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/mask_tests.log
timeout 100 python scripts/fuzz_k1.py --seconds 45 --seed 31 > gpurun_out/fuzz_k1.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/fuzz_k1.log
python scripts/perf_k1.py --dtype float32 --mask nan 2>&1 | grep -E "fast-path"
python scripts/perf_k1.py --dtype float32 2>&1 | grep -E "fast-path"
python scripts/perf_k1.py --dtype float32 --mask nan --land 0.3 2>&1 | grep -E "fast-path"
python scripts/perf_k1.py --dtype float32 --mask nan --axis 1 2>&1 | grep -E "fast-path"
python scripts/perf_k1.py --dtype float32 --axis 1 2>&1 | grep -E "fast-path"
python scripts/perf_k1.py --dtype float32 --mask nan --shape 64,137,256,512 --axis 2 2>&1 | grep -E "fast-path|K1"
python scripts/perf_chunked.py --compare --chunk 744,103,180 2>&1 | tail -2
