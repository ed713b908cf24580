set -x
python scripts/perf_k1.py --shape 2048,721,1440 --dtype float32 --check --bitround 2>&1 | tail -8
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float64 --check 2>&1 | tail -5
python scripts/perf_k1.py --shape 2048,721,1440 --dtype float16 --check 2>&1 | tail -5
python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan --check 2>&1 | tail -5
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
