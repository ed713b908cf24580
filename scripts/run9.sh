python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python scripts/perf_k1.py --shape 2048,721,1440 --dtype float32 --check 2>&1 | tail -3
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float64 --check 2>&1 | tail -3
python scripts/perf_k1.py --shape 2048,721,1440 --dtype float16 --check 2>&1 | tail -3
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan --check 2>&1 | tail -3
python bench.py --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
