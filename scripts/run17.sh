for vpt in 2 4 8; do for hint in 0 1; do for waves in 8 16 64; do echo "VPT=$vpt HINT=$hint WAVES=$waves"; XBI_BR_VPT=$vpt XBI_BR_HINT=$hint XBI_BR_WAVES=$waves python scripts/perf_k3.py 2>&1 | grep float32; done; done; done
echo default; python scripts/perf_k3.py
