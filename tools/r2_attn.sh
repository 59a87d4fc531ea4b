#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_crops.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3; do
echo -n "new   "; timeout 200 python bench.py --only-value --side none --steps 40 --warmup 5 2>&1 | tail -1
echo -n "prev  "; OVO_B200_LIB=/root/repo/tools/_ab/libovo_prev.so timeout 200 python bench.py --only-value --side none --steps 40 --warmup 5 2>&1 | tail -1
done
