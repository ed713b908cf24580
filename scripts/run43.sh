python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; tail -c 3500 gpurun_out/bench_full.log
