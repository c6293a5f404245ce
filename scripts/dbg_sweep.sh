#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for d in 0 1 2 3 4 7; do echo "== SRK_TC5_DBG=$d"; SRK_TC5_DBG=$d python scripts/gemm_bench.py tcgen05 2>&1 | grep -E "qkv|bf16 only|fc1|up2"; done
