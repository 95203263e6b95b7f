#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "rc=$?" >> gpurun_out/r2g_bench.err
tail -3 gpurun_out/r2g_bench.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
print({k: l[k] for k in ("value", "ms_per_step", "gpu_launches")}, l["e2e"]["value"], l["roofline"]["frac"])
print(json.dumps(l.get("train_step"), indent=1)[:3000])
print(l.get("latency_b1", {}).get("ms_per_net_eval"))
PY
