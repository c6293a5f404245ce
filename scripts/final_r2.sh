#!/bin/bash
# round-2 closing evidence on one GPU: launch list of a step, ncu --set full of the block kernels, the default bench
# line (with the eager baselines) and the other BASELINE configs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
bash scripts/ncu_step_r2.sh > gpurun_out/final_step.log 2>&1
bash scripts/ncu_block_r2.sh > gpurun_out/final_block.log 2>&1
timeout 400 python bench.py > gpurun_out/final_bench_cfg3.json 2>gpurun_out/final_bench.err
for w in cfg1 cfg2 cfg4 cfg5; do
  timeout 200 python bench.py --workload $w --no-cpu-baseline --no-eager-baseline > gpurun_out/final_bench_$w.json 2>>gpurun_out/final_bench.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_bench_*.json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 3), round(d.get("e2e", {}).get("value", 0), 1), d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
head -12 gpurun_out/r02_launches.txt
