python -m pytest tests/test_cli_gpu.py -x -q 2>&1 | tail -30
