#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python tools/attn_trace.py 2>&1 | head -24
timeout 200 python bench.py --only-value --steps 40 --warmup 5 2>&1 | tail -1
