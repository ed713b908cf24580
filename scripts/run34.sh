python -m pytest tests/test_counts_gpu.py tests/test_guard_page_gpu.py -m gpu -x -q 2>&1 | tail -2
run() { for m in none nan; do
echo "== $1 $2 axis=$3 mask=$m"
python scripts/perf_k1.py --shape $1 --dtype $2 --axis $3 --mask $m 2>&1 | grep -E "alone"
done; }
run 20,75,1080,1440 float64 2
run 72,137,256,512 float32 0
run 72,137,256,512 float32 1
run 72,137,256,512 float32 2
run 1024,721,1440 float32 1
run 1024,721,1440 float32 0
run 2048,721,1440 float16 1
python scripts/perf_k1.py --shape 20,75,1080,1440 --dtype float64 --axis 2 --mask nan --land 0.3 | grep alone
