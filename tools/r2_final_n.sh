#!/bin/bash
# round-2 multi-GPU verification: N = $1 (sharded-map parity suite at N=2, then the bench line under torchrun)
N=$1
cd /root/repo; mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
fi
S=$SECONDS
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 400 gpurun_out/r2_bench_n$N.err; head -c 700 gpurun_out/r2_bench_n$N.json; echo
echo "bench N=$N: $((SECONDS-S)) s"
