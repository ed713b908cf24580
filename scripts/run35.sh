for shp in 10000000,100 40000000,25 4000000,250 2000000,500; do
echo "== $shp"; python scripts/perf_k1.py --shape $shp --dtype float32 --axis 1 --mask none 2>&1 | grep -E "alone|K1|launches"
done
echo "== f16 8000000,250"; python scripts/perf_k1.py --shape 8000000,250 --dtype float16 --axis 1 --mask none 2>&1 | grep -E "alone|K1"
