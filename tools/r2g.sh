python tools/fuse_bench.py
for S in none fuse all; do python bench.py --steps 20 --warmup 3 --only-value --side $S; done
OVO_B200_ENC_PRIO=0 python bench.py --steps 20 --warmup 3 --only-value --side all
