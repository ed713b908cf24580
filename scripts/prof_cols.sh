python scripts/perf_k1.py --shape 365,4,1080,1440 --dtype float64 --axis 2
python scripts/perf_k1.py --shape 36,137,256,512 --dtype float32 --axis 2
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float16
