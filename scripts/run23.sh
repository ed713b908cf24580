set -x
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
time python -m pytest tests/test_full_size_gpu.py -m gpu -x -q --durations=10 2>&1 | tail -25
