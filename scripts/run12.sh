python -m pytest tests -x -q -m gpu 2>&1 | grep -E "Error|error|FAILED|assert" | head -20
