python -m pytest tests/test_gpu_sharded.py tests/test_gpu_ovo.py -x -q -m gpu 2>&1 | tail -25
