#!/bin/bash
# swizzled lane cells + a sixth ring slot for the walkers along the innermost axis
mkdir -p gpurun_out
python -m pytest tests/test_chunked_gpu.py -x -q 2>&1 | tail -1
timeout 100 python scripts/fuzz_gpu.py --seconds 35 --seed 113 2>&1 | tail -1
python scripts/perf_chunked.py --compare 2>&1 | grep "axis 2" | head -2
python scripts/perf_chunked.py --compare --dtype float64 --shape 372,721,1440 2>&1 | grep "axis 2" | head -2
