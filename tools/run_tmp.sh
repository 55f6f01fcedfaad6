python -m pytest tests/test_net_gpu.py tests/test_train_samples.py -x -q 2>&1 | tail -15
