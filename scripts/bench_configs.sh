#!/bin/bash
# bench lines of the other BASELINE configs, the eval.py geometry and the legacy engine (1 GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); r = d["roofline"]
print(sys.argv[1], round(d["value"], 1), "patches/s", round(d["ms_per_step"], 3), "ms/step", r["bound"], round(r["frac"], 3),
      "e2e", round(d["e2e"]["value"], 1), d["config"]["net_input"])
PY
}
for wl in cfg1 cfg2 cfg4 cfg5; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  show gpurun_out/bench_$wl.json || tail -3 gpurun_out/bench_$wl.err
done
timeout 600 python bench.py --workload cfg3 --geometry evalpy --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_evalpy.json 2>/dev/null
show gpurun_out/bench_cfg3_evalpy.json
timeout 600 python bench.py --engine mma_sync --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mma_sync.json 2>/dev/null
show gpurun_out/bench_mma_sync.json
