#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for wl in cfg1 cfg2 cfg4 cfg5; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  python -c "import json;d=json.load(open('gpurun_out/bench_$wl.json'));print('$wl',round(d['value'],1),'patches/s',round(d['ms_per_step'],3),'ms/step frac',round(d['roofline']['whole_step_frac'],4), d['config']['net_input'])" || tail -3 gpurun_out/bench_$wl.err
done
timeout 600 python bench.py --workload cfg3 --geometry evalpy --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_evalpy.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_cfg3_evalpy.json'));print('cfg3 eval.py geometry',round(d['value'],1),round(d['ms_per_step'],3),round(d['roofline']['whole_step_frac'],4))"
timeout 600 python bench.py --engine mma_sync --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mma_sync.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_mma_sync.json'));print('cfg3 legacy mma.sync engine',round(d['value'],1),round(d['ms_per_step'],3))"
