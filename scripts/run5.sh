nproc; free -g | head -2
python bench.py --time-steps 1024 --steps 10 --warmup 3 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
