#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qkv_attn -c 1 -f -o gpurun_out/prof_qa python scripts/attn_bench.py > gpurun_out/ncu_qa.log 2>&1
tail -2 gpurun_out/ncu_qa.log
