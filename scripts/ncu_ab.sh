#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SRK_PROFILE_ONCE=1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_block -c 1 -f -o gpurun_out/prof_ab python scripts/gemm_bench.py tcgen05 > gpurun_out/ncu_ab.log 2>&1
tail -2 gpurun_out/ncu_ab.log
