#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SRK_PROFILE_ONCE=1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -c ${1:-4} -f -o gpurun_out/prof_gemm3 python scripts/gemm_bench.py tcgen05 > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
