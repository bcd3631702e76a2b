timeout 300 python -m pytest tests/test_runtime.py -m gpu -q -x --timeout 120 -k streamed 2>&1 | grep -v "^$" | tail -30
