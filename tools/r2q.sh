set -x
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "route or batch or association or depth" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-sam > gpurun_out/r2q_bench2.json 2> gpurun_out/r2q_bench2.err
tail -3 gpurun_out/r2q_bench2.err
