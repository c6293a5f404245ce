#!/bin/bash
# round-2 multi-GPU lines: strong-scaling cfg4 sweep (1471 patches) and weak-scaling cfg3 / cfg5 at N = 4, 8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { # N port args...
  local n=$1 port=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" 2>gpurun_out/scale.err
}
for n in 4 8; do
  run $n 2951$n --workload cfg4 --sweep 1471 --steps 3 > gpurun_out/bench_r2_cfg4_sweep_n$n.json
  run $n 2952$n --workload cfg3 --steps 10 --no-cpu-baseline > gpurun_out/bench_r2_cfg3_n$n.json
  run $n 2953$n --workload cfg5 --steps 5 --no-cpu-baseline > gpurun_out/bench_r2_cfg5_n$n.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r2_cfg*_n[48].json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], d["n_gpus"], d["scaling"], round(d["value"], 1), round(d.get("e2e", {}).get("value", 0), 1), d["config"].get("shard_sizes"), d.get("sharded_vs_single_gpu_means_max_rel_diff"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 300 gpurun_out/scale.err
