python -m pytest tests/test_counts_gpu.py tests/test_guard_page_gpu.py -m gpu -x -q 2>&1 | tail -2
echo "== rows f32 clean / land / rand 0.01 / 0.1"; for o in "" "--land 0.3" "--randmask 0.01" "--randmask 0.1"; do python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --mask nan $o | grep alone; done
echo "== rows f64 clean / land / rand 0.01 / 0.1"; for o in "" "--land 0.3" "--randmask 0.01" "--randmask 0.1"; do python scripts/perf_k1.py --shape 512,721,1440 --dtype float64 --mask nan $o | grep alone; done
echo "== rows f16 clean / rand 0.01"; for o in "" "--randmask 0.01"; do python scripts/perf_k1.py --shape 2048,721,1440 --dtype float16 --mask nan $o | grep alone; done
echo "== cols f64 clean / land / rand 0.01"; for o in "" "--land 0.3" "--randmask 0.01"; do python scripts/perf_k1.py --shape 20,75,1080,1440 --dtype float64 --axis 2 --mask nan $o | grep alone; done
echo "== cols f32 clean / land"; for o in "" "--land 0.3"; do python scripts/perf_k1.py --shape 1024,721,1440 --dtype float32 --axis 1 --mask nan $o | grep alone; done
