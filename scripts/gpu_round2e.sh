#!/bin/bash
# cooperative copies of the chunk walkers along the innermost axis
mkdir -p gpurun_out
python -m pytest tests/test_chunked_gpu.py -x -q > gpurun_out/chunk_tests.log 2>&1; echo "chunk tests rc=$?"; tail -2 gpurun_out/chunk_tests.log
timeout 120 python scripts/fuzz_gpu.py --seconds 60 --seed 23 > gpurun_out/fuzz_gpu.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/fuzz_gpu.log
python scripts/perf_chunked.py --compare 2>&1 | grep "axis 2"
python scripts/perf_chunked.py --compare --dtype float64 --shape 372,721,1440 2>&1 | grep "axis 2"
python scripts/perf_chunked.py --compare --nan land --chunk 744,103,180 2>&1 | grep "axis 2"
python scripts/perf_chunked.py --compare --shape 744,720,1441 --chunk 744,103,131 2>&1 | grep "axis"
