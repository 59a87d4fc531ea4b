#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ovo.py -m gpu -x -q -k "fuse or dense or growth" 2>&1 | tail -2
python tools/fuse_bench.py 2>&1 | tail -1
OVO_B200_FUSE8=0 python tools/fuse_bench.py 2>&1 | tail -1
python bench.py --only-value --steps 40 --warmup 5 2>&1 | tail -1
OVO_B200_FUSE8=0 python bench.py --only-value --steps 40 --warmup 5 2>&1 | tail -1
