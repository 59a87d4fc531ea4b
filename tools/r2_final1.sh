#!/bin/bash
# round-2 final verification on ONE GPU: the whole gpu suite, smoke(), the default bench line and the reference arm
cd /root/repo; mkdir -p gpurun_out
S=$SECONDS
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "pytest: $((SECONDS-S)) s"; S=$SECONDS
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "smoke: $((SECONDS-S)) s"; S=$SECONDS
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 600 gpurun_out/r2_bench_n1.err; head -c 1500 gpurun_out/r2_bench_n1.json
echo "bench: $((SECONDS-S)) s"; S=$SECONDS
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cat gpurun_out/r2_bench_ref.json | head -c 800
echo "reference arm: $((SECONDS-S)) s"
