#!/bin/bash
# closing multi-GPU lines of round 2 at N GPUs (default 8): cfg3 / cfg5 weak scaling, cfg4 1471-patch sweep (strong)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
n=${1:-8}
run() { local port=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" 2>gpurun_out/scale.err
}
run 29611 --workload cfg3 --steps 10 --no-cpu-baseline --no-eager-baseline > gpurun_out/fin_cfg3_n$n.json
run 29612 --workload cfg4 --sweep 1471 --steps 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/fin_cfg4_sweep_n$n.json
run 29613 --workload cfg5 --steps 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/fin_cfg5_n$n.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/fin_cfg*_n*.json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], d["n_gpus"], d["scaling"], round(d["value"], 1), round(d.get("e2e", {}).get("value", 0), 1), d["config"].get("shard_sizes"), d.get("sharded_vs_single_gpu_means_max_rel_diff"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 300 gpurun_out/scale.err
