python -m pytest tests/test_counts_gpu.py -x -q -m gpu 2>&1 | tail -2
for ax in 1 2; do python scripts/perf_k1.py --shape 72,137,256,512 --dtype float32 --axis $ax --check 2>&1 | tail -4; done
python scripts/perf_k1.py --shape 365,8,1080,1440 --dtype float64 --axis 2 --check 2>&1 | tail -4
