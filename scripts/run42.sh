python -m pytest tests/test_counts_gpu.py tests/test_guard_page_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -3
for o in "" "--land 0.3" "--randmask 0.001" "--randmask 0.01" "--randmask 0.1" "--randmask 0.5"; do
echo "== rows f32 $o"; python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan $o | grep alone
echo "== cols f32 axis1 $o"; python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --axis 1 --mask nan $o | grep alone
done
echo "== dense forced on clean rows/cols f32"; python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan --flavour dense | grep alone; python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --axis 1 --mask nan --flavour dense | grep alone
echo "== f64 rows rand 0.01 / cols f64 rand 0.01 / f16 rows rand 0.01"
python scripts/perf_k1.py --shape 512,721,1440 --dtype float64 --mask nan --randmask 0.01 | grep alone
python scripts/perf_k1.py --shape 20,75,1080,1440 --dtype float64 --axis 2 --mask nan --randmask 0.01 | grep alone
python scripts/perf_k1.py --shape 2048,721,1440 --dtype float16 --mask nan --randmask 0.01 | grep alone
