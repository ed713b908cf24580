python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for ax in 0 1 2 3; do python scripts/perf_k1.py --shape 72,137,256,512 --dtype float32 --axis $ax 2>&1 | tail -2; done
python scripts/perf_k1.py --shape 365,8,1080,1440 --dtype float64 --axis 2 2>&1 | tail -2
python scripts/perf_k1.py --shape 365,8,1080,1440 --dtype float64 --axis 2 --mask nan 2>&1 | tail -2
python scripts/perf_k1.py --shape 365,8,1080,1440 --dtype float64 --axis 2 --mask -999 2>&1 | tail -2
