set -x
python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
for V in persistent launches; do
OVO_B200_VOTE=$V timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-sam > gpurun_out/r2p_bench2_$V.json 2> gpurun_out/r2p_bench2_$V.err
tail -3 gpurun_out/r2p_bench2_$V.err
done
