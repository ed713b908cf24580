python -m pytest tests/test_counts_gpu.py tests/test_guard_page_gpu.py -m gpu -x -q 2>&1 | tail -2
for f in 0.001 0.01 0.1 0.5; do
echo "== rows f32 randmask $f"; python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan --randmask $f | grep alone
echo "== cols f32 axis1 randmask $f"; python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --axis 1 --mask nan --randmask $f | grep alone
done
echo "== rows f64 randmask 0.01"; python scripts/perf_k1.py --shape 512,721,1440 --dtype float64 --mask nan --randmask 0.01 | grep alone
echo "== rows f16 randmask 0.01"; python scripts/perf_k1.py --shape 2048,721,1440 --dtype float16 --mask nan --randmask 0.01 | grep alone
echo "== rows f32 clean / land"; python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan | grep alone; python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan --land 0.3 | grep alone
echo "== cols f64 land"; python scripts/perf_k1.py --shape 20,75,1080,1440 --dtype float64 --axis 2 --mask nan --land 0.3 | grep alone
