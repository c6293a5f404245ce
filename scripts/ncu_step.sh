#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
# launch list of ONE step (second step: skip the 1st step's launches)
SRK_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 340 --csv --log-file gpurun_out/launches.csv python scripts/one_step.py > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
# full captures of one instance of the non-GEMM kernels and the qkv GEMM
SRK_STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'window_attention|metrics_tile|conv_out|layernorm' -s 2 -c 4 -f -o gpurun_out/prof_misc python scripts/one_step.py > gpurun_out/ncu_misc.log 2>&1
tail -2 gpurun_out/ncu_misc.log
SRK_PROFILE_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -c 5 -f -o gpurun_out/prof_gemm2 python scripts/gemm_bench.py tcgen05 > gpurun_out/ncu_gemm2.log 2>&1
tail -2 gpurun_out/ncu_gemm2.log
ls -la gpurun_out/*.ncu-rep
