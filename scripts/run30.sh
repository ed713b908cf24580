python -m pytest tests/test_counts_gpu.py -m gpu -x -q 2>&1 | tail -6
for shp in 64,137,256,513 64,137,257,511; do for m in none nan; do
echo "== $shp f32 axis 2 mask=$m"; python scripts/perf_k1.py --shape $shp --dtype float32 --axis 2 --mask $m | grep -E "alone|K1"
done; done
echo "== f64 (20,75,1080,1441) axis 2"; python scripts/perf_k1.py --shape 10,75,1080,1441 --dtype float64 --axis 2 --mask nan | grep -E "alone|K1"
