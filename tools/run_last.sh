timeout 120 python -m pytest tests -m gpu -q -x -k "golden" 2>&1 | tail -4
