set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"associate_batch_pass|batch_vote_scan" -s 27 -c 9 -o gpurun_out/r2_map -f python tools/fuse_bench.py > gpurun_out/r2_cap_map.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fuse_dense_batch" -s 4 -c 2 -o gpurun_out/r2_fuse -f python tools/fuse_bench.py > gpurun_out/r2_cap_fuse.log 2>&1
ls -la gpurun_out/*.ncu-rep
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 300 gpurun_out/r2_bench_n1.err
