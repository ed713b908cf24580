set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests/test_counts_gpu.py -x -q 2>&1 | tail -30
