#!/bin/bash
# first GPU trip: legacy-engine parity + a first timing
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
export SRK_TEST_ENGINES=mma_sync
timeout 1200 python -m pytest tests/test_gpu.py -m gpu -q 2>&1 | tail -150 > gpurun_out/pytest_mma.log
tail -15 gpurun_out/pytest_mma.log
true
true
