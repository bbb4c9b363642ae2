python -m pytest tests -m gpu -q -x -k "eval" 2>&1 | tail -15
