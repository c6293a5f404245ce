#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SRK_STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'metrics_fast_kernel|conv_in_ln' -c 2 -f -o gpurun_out/prof_met python scripts/one_step.py > /dev/null 2>&1
ls -la gpurun_out/prof_met.ncu-rep
