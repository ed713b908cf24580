run() { for v in XBI_COLS_LDG XBI_COLS_TMA; do for m in none nan; do
echo "== $1 $2 axis=$3 $v mask=$m"
env $v=1 python scripts/perf_k1.py --shape $1 --dtype $2 --axis $3 --mask $m 2>&1 | grep alone
done; done; }
run 20,75,1080,1440 float64 2
run 64,137,256,512 float32 1
run 64,137,256,512 float32 0
run 1024,721,1440 float32 1
run 1024,721,1440 float32 0
run 2048,721,1440 float16 1
