SECONDS=0
python bench.py --no-cpu-baseline --no-next-rows --no-gpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo "bench wall seconds: $SECONDS"
tail -12 gpurun_out/r2l_bench.err | cut -c1-300
