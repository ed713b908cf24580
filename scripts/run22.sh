set -x
python -m pytest tests/test_counts_gpu.py -m gpu -x -q 2>&1 | tail -4
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan --land 0.3
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask -999 --land 0.3
python scripts/perf_k1.py --shape 512,721,1440 --dtype float64 --mask nan --land 0.3
python scripts/perf_k1.py --shape 20,75,1080,1440 --dtype float64 --axis 2 --mask nan --land 0.3
python scripts/perf_k1.py --shape 64,137,256,512 --dtype float32 --axis 2 --mask nan
python scripts/perf_k1.py --shape 64,137,256,512 --dtype float32 --axis 1 --mask nan --land 0.3
