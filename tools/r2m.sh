python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "batch or association" 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-next-rows --no-gpu-baseline --no-sam --no-configs --no-stream --profile-e2e > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
head -60 gpurun_out/r2m_bench.err | cut -c1-200
