for shp in 64,137,256,500 64,137,256,512 32,137,256,1024 64,64,256,1440 16,137,256,2048 16,100,256,2880 8,137,256,4096 8,100,256,5000; do
for v in XBI_COLS_LDG XBI_COLS_TMA; do
echo "== $shp $v"
env $v=1 python scripts/perf_k1.py --shape $shp --dtype float32 --axis 2 --mask none 2>&1 | grep alone
done; done
