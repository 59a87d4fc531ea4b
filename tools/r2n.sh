python -m pytest tests/test_gpu_encoder.py -x -q -m gpu 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-next-rows --no-sam --no-configs --no-stream > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
tail -5 gpurun_out/r2n_bench.err
