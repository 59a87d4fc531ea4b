python -m pytest tests/test_gpu_ovo.py tests/test_gpu_crops.py tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -8
python bench.py --no-cpu-baseline --no-next-rows --no-sam --no-configs --no-gpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
tail -5 gpurun_out/r2o_bench.err
