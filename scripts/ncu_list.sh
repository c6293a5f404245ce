#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
K='regex:gemm_|mlp_|qkv_attn|window_attention|layernorm|conv_in|conv_out|metrics_'
SRK_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_traffic.csv python scripts/one_step.py ${1:-cfg3} > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
