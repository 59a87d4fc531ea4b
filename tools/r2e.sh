set -x
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "batch or association or launch" 2>&1 | tail -3
for X in p2p; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-sam --exchange $X > gpurun_out/r2e_bench2_$X.json 2> gpurun_out/r2e_bench2_$X.err
tail -5 gpurun_out/r2e_bench2_$X.err
done
