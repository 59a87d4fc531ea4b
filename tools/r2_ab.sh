#!/bin/bash
# same-box A/B of tuning switches on the device-resident step
cd /root/repo
run() { # label, env...
  local l=$1; shift
  env "$@" python bench.py --only-value --steps 40 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$l', d['value'], d['ms_per_step'])"
}
python -m pytest tests/test_gpu_encoder.py tests/test_gpu_crops.py -m gpu -x -q 2>&1 | tail -2
run tail-on X=1
run tail-off OVO_B200_ATTN_TAIL=0
run tail-on X=1
run tail-off OVO_B200_ATTN_TAIL=0
