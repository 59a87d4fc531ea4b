#!/bin/bash
cd /root/repo

for i in 1 2 3 4; do
echo -n "new   "; timeout 200 python bench.py --only-value --side none --steps 40 --warmup 5 2>&1 | tail -1
echo -n "prev  "; OVO_B200_LIB=/root/repo/tools/_ab/libovo_prev.so timeout 200 python bench.py --only-value --side none --steps 40 --warmup 5 2>&1 | tail -1
done
