python bench.py --steps 10 --warmup 3 --no-sam --no-stream --no-next-rows --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -20 gpurun_out/r2j_bench.err
