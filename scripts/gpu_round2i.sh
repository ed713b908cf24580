#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_chunked_gpu.py -x -q > gpurun_out/chunk_tests.log 2>&1; echo "chunk tests rc=$?"; tail -2 gpurun_out/chunk_tests.log
timeout 120 python scripts/fuzz_gpu.py --seconds 30 --seed 43 > gpurun_out/fuzz_gpu.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/fuzz_gpu.log
python scripts/perf_chunked.py --compare 2>&1 | grep "axis 2"
python scripts/perf_chunked.py --compare --dtype float64 --shape 372,721,1440 2>&1 | grep "axis 2"
