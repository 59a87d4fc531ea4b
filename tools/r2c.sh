set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fuse_dense_batch|fuse_counts|batch_vote_scan" -s 12 -c 4 -o gpurun_out/r2c_fuse python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sam --no-stream --no-next-rows > gpurun_out/r2c_ncu1.log 2>&1
