#!/bin/bash
# Same-box A/B (run under gpurun): the working tree's library against a previous build of it.
#   1. build the older commit in a git worktree and copy its ovo_b200/libovo_b200.so to tools/_ab/libovo_prev.so (git-ignored)
#   2. gpurun -- bash tools/r2_ab.sh [bench flags]
# Boxes of the pool differ by +-2 % on the same build, so only runs of ONE call are compared; the pairs alternate to expose drift.
cd /root/repo
FLAGS=${@:---only-value --side none --steps 40 --warmup 5}
for i in 1 2 3; do
  echo -n "new   "; timeout 200 python bench.py $FLAGS 2>&1 | tail -1
  echo -n "prev  "; OVO_B200_LIB=/root/repo/tools/_ab/libovo_prev.so timeout 200 python bench.py $FLAGS 2>&1 | tail -1
done
