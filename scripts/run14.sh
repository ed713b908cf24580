python scripts/clock_probe.py 72,137,256,512 float32 2 1500
XBI_COLS_LDG=1 python scripts/clock_probe.py 72,137,256,512 float32 2 1500
python scripts/clock_probe.py 72,137,256,512 float32 0 1500
python scripts/clock_probe.py 2048,721,1440 float32 -1 1000
