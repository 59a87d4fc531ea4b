#!/bin/bash
# folded LayerNorm: A/B of the step (alternating, to see the noise)
cd /root/repo
python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -2
for f in 0 1 0 1; do
  OVO_B200_FOLD_LN=$f python bench.py --only-value --steps 40 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('fold=$f', d['value'], d['ms_per_step'])"
done
