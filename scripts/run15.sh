python -m pytest tests/test_counts_gpu.py -x -q -m gpu 2>&1 | tail -2
python scripts/perf_k1.py --shape 2048,721,1440 --dtype float16 --check 2>&1 | tail -4
python scripts/perf_k1.py --shape 72,137,256,512 --dtype float16 --axis 2 --check 2>&1 | tail -4
python scripts/perf_k1.py --shape 72,137,256,500 --dtype float16 --axis 1 --check 2>&1 | tail -4
python scripts/perf_k1.py --shape 2048,721,1440 --dtype float16 --mask nan --check 2>&1 | tail -4
