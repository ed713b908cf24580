cp gpurun_scratch/libold.so xbitinfo_b200/csrc/libxbitinfo_b200.so
for shp in 64,137,256,512 64,64,256,1440 8,100,256,5000; do
for v in XBI_COLS_LDG XBI_COLS_TMA; do
echo "== OLD f32 $shp $v"
env $v=1 python scripts/perf_k1.py --shape $shp --dtype float32 --axis 2 --mask none 2>&1 | grep alone
done; done
for v in XBI_COLS_LDG XBI_COLS_TMA; do
echo "== OLD f64 8,75,1080,1440 $v"
env $v=1 python scripts/perf_k1.py --shape 8,75,1080,1440 --dtype float64 --axis 2 --mask none 2>&1 | grep alone
done
