# Round-2 evidence captures (run under gpurun, ONE GPU): launch list of a bench step + ncu --set full of the dominant kernels.
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 1400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --only-value > gpurun_out/r2_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tn" -s 10 -c 8 -o gpurun_out/r2_gemm python tools/profile_target.py gemm > gpurun_out/r2_cap_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attention_fwd" -s 4 -c 2 -o gpurun_out/r2_attn python tools/profile_target.py attn > gpurun_out/r2_cap_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tn" -s 2 -c 1 -o gpurun_out/r2_query python tools/profile_target.py query > gpurun_out/r2_cap_query.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"associate_batch_pass|batch_vote_scan" -s 27 -c 9 -o gpurun_out/r2_map python tools/fuse_bench.py > gpurun_out/r2_cap_map.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fuse_dense_batch" -s 4 -c 2 -o gpurun_out/r2_fuse python tools/fuse_bench.py > gpurun_out/r2_cap_fuse.log 2>&1
ls -la gpurun_out/*.ncu-rep
