python scripts/perf_k1.py --shape 36,137,256,512 --dtype float32 --axis 2
XBI_COLS_LDG=1 python scripts/perf_k1.py --shape 36,137,256,512 --dtype float32 --axis 2
