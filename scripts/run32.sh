python -m pytest tests/test_counts_gpu.py -m gpu -x -q 2>&1 | tail -2
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask none | grep -E "alone|K1"
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan | grep -E "alone|K1"
python scripts/perf_k1.py --shape 72,137,256,512 --dtype float32 --axis 1 --mask none | grep -E "alone|K1"
python scripts/perf_k1.py --shape 20,75,1080,1440 --dtype float64 --axis 2 --mask none | grep -E "alone|K1"
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float64 --mask none | grep -E "alone|K1"
python scripts/perf_k1.py --shape 2048,721,1440 --dtype float16 --mask none | grep -E "alone|K1"
python scripts/perf_k1.py --shape 2048,721,1440 --dtype float16 --axis 1 --mask none | grep -E "alone|K1"
python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().splitlines()[-1]); print('bench value',d['value'],'kernel',d['roofline']['achieved'],'default',d['default_kwargs_step_GBps'],d['clocks'])"
