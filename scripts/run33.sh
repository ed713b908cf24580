set -x
which compute-sanitizer
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_counts_gpu.py -m gpu -x -q -k "element_aligned or mask_patterns_column or strided_fast_path or row_ends or every_axis" 2>&1 | tail -15
echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_shuffle_gpu.py -m gpu -x -q -k "not large" 2>&1 | tail -8
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_shuffle_gpu.py tests/test_counts_gpu.py -m gpu -x -q -k "(bit_shuffle and float32 and 4160) or (mask_patterns_rows and float32 and alternating) or (mask_patterns_column and float32 and blocks)" 2>&1 | tail -12
