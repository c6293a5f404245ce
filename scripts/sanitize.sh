#!/bin/bash
# compute-sanitizer passes over the round-2 kernels (small parity tests): memcheck, then racecheck of the shared-memory protocols
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SRK_TEST_ENGINES=tcgen05
K='test_attn_block_kernel_equals_unfused_sequence and (geom3 or geom4) or test_metrics_streaming_kernel_geometries and (shape0 or shape3) or test_conv_halo_tiles_equal_per_tap_boxes and (geom4 or geom3) or test_fused_mlp_kernel and dims1'
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu.py -m gpu -q -x -k "$K" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/sanitize_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 1 python -m pytest tests/test_gpu.py -m gpu -q -x -k "$K" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed|Error" gpurun_out/sanitize_racecheck.log | head -20
