#!/bin/bash
# why do the two-stream walkers trail the three-stream ones on 10x10-column chunks?
mkdir -p gpurun_out
python scripts/perf_chunked.py --compare --chunk 744,10,10 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"k_pairs_chunked_walk" -s 3 -c 6 -o gpurun_out/prof_r02b_chunk_10x10 -f python scripts/perf_chunked.py --compare --chunk 744,10,10 > gpurun_out/r02b_ncu_g.log 2>&1
ls -la gpurun_out/prof_r02b_chunk_10x10.ncu-rep
