set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python bench.py --steps 10 --warmup 3 --no-sam --no-stream --no-next-rows > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 1500 gpurun_out/r2a_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fuse_dense_batch|associate_pass1" -s 6 -c 4 -o gpurun_out/r2a_fuse python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sam --no-stream --no-next-rows > gpurun_out/r2a_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_fwd" -s 60 -c 2 -o gpurun_out/r2a_attn python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sam --no-stream --no-next-rows > gpurun_out/r2a_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
