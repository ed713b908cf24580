#!/bin/bash
# chunk-walker rewrite: correctness (tests + fuzz), then kernel-alone timings
mkdir -p gpurun_out
python -m pytest tests/test_chunked_gpu.py -x -q > gpurun_out/chunk_tests.log 2>&1; echo "chunk tests rc=$?"; tail -3 gpurun_out/chunk_tests.log
timeout 120 python scripts/fuzz_gpu.py --seconds 45 --seed 7 > gpurun_out/fuzz_gpu.log 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/fuzz_gpu.log
python scripts/perf_chunked.py --compare > gpurun_out/perf_chunk_f32.log 2>&1; cat gpurun_out/perf_chunk_f32.log
python scripts/perf_chunked.py --compare --dtype float64 --shape 372,721,1440 > gpurun_out/perf_chunk_f64.log 2>&1; cat gpurun_out/perf_chunk_f64.log
python scripts/perf_chunked.py --compare --nan land --chunk 744,103,180 > gpurun_out/perf_chunk_land.log 2>&1; cat gpurun_out/perf_chunk_land.log
python scripts/perf_chunked.py --compare --nan 0.01 --chunk 744,103,180 > gpurun_out/perf_chunk_s1.log 2>&1; cat gpurun_out/perf_chunk_s1.log
python scripts/perf_chunked.py --compare --nan 0.1 --chunk 744,103,180 > gpurun_out/perf_chunk_s10.log 2>&1; cat gpurun_out/perf_chunk_s10.log
