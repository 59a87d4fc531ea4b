set -x
python tools/fuse_bench.py
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fuse or dense or batch" 2>&1 | tail -3
