#!/bin/bash
# launch list (device time + DRAM bytes per launch) of ONE cfg3 step, plain launches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SRK_STEPS=2 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_tc5|attn_block|mlp_tc5|metrics_|conv_in|conv_out|tail_border|layernorm|window_attention|bicubic" -c 200 --csv --log-file gpurun_out/r02_launches.csv python scripts/one_step.py > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
python scripts/launch_summary.py gpurun_out/r02_launches.csv gpurun_out/r02_step_traffic.json | tee gpurun_out/r02_launches.txt
