#!/bin/bash
# full ncu captures of the kernels of one Swin block + the RSTB conv (per-shape microbench, one launch each)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SRK_PROFILE_ONCE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_block|mlp_tc5|gemm_tc5_kernel<192, 3" -c 3 -f -o gpurun_out/prof_block_r2 python scripts/gemm_bench.py tcgen05 > gpurun_out/ncu_block_r2.log 2>&1
tail -2 gpurun_out/ncu_block_r2.log
