python -m pytest tests/test_counts_gpu.py -x -q -m gpu 2>&1 | tail -2
python scripts/perf_k1.py --shape 365,8,1080,1440 --dtype float64 --axis 2 --check 2>&1 | tail -3
python scripts/perf_k1.py --shape 72,137,256,512 --dtype float32 --axis 0 --check 2>&1 | tail -3
python scripts/perf_k1.py --shape 72,137,256,512 --dtype float32 --axis 1 --check 2>&1 | tail -3
python scripts/perf_k1.py --shape 72,137,256,512 --dtype float32 --axis 2 --check 2>&1 | tail -3
python scripts/perf_k1.py --shape 365,8,1080,1440 --dtype float64 --axis 2 --mask nan --check 2>&1 | tail -3
