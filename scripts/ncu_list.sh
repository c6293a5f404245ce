#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SRK_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_|window_attention|layernorm|conv_in|conv_out|metrics_" -s 204 -c 204 --csv --log-file gpurun_out/launches.csv python scripts/one_step.py ${1:-cfg3} > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
