set -x
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ovo.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-sam --no-stream --no-next-rows --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 2500 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
