set -x
nvidia-smi -L | wc -l
for X in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-sam --exchange $X > gpurun_out/r2f_bench8_$X.json 2> gpurun_out/r2f_bench8_$X.err
tail -5 gpurun_out/r2f_bench8_$X.err
done
