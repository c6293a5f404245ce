#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python scripts/gemm_bench.py tcgen05 2>&1 | grep -E "fc1|fc2|MLP"
export SRK_PROFILE_ONCE=1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:mlp_tc5 -c 1 -f -o gpurun_out/prof_mlp python scripts/gemm_bench.py tcgen05 > gpurun_out/ncu_mlp.log 2>&1
tail -2 gpurun_out/ncu_mlp.log
