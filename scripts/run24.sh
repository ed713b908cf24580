set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask none
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan --land 0.3
python scripts/perf_k1.py --shape 64,137,256,512 --dtype float32 --axis 2 --mask nan
python scripts/perf_k1.py --shape 64,137,256,500 --dtype float32 --axis 2 --mask nan
python scripts/perf_k1.py --shape 64,137,256,500 --dtype float32 --axis 2 --mask none
python bench.py > gpurun_out/bench_r24.log 2> gpurun_out/bench_r24.err; tail -c 3000 gpurun_out/bench_r24.log
