#!/bin/bash
# round-end evidence: tests, bench, launch list with DRAM traffic, full capture of the top kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SRK_TEST_ENGINES=tcgen05,mma_sync
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_final.log; tail -2 gpurun_out/pytest_final.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
K='regex:gemm_|mlp_|qkv_attn|window_attention|layernorm|conv_in|conv_out|metrics_|tail_border'
SRK_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_traffic.csv python scripts/one_step.py > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
# full captures: the 4 GEMM kernels of one Swin block + one RSTB conv (launches 2..6 of the second step), then the tail
SRK_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 1 -c 4 -f -o gpurun_out/prof_block_final python scripts/one_step.py > gpurun_out/ncu_block_final.log 2>&1
SRK_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc5' -s 144 -c 10 -f -o gpurun_out/prof_tail_final python scripts/one_step.py > /dev/null 2>&1
SRK_STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'metrics_fast_kernel|conv_in_ln|tail_border' -c 3 -f -o gpurun_out/prof_misc_final python scripts/one_step.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
