#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2z_tests.log 2>&1; tail -n 2 gpurun_out/r2z_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1; tail -n 2 gpurun_out/r2z_smoke.log
timeout 1500 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
echo "rc=$?" >> gpurun_out/r2z_bench.err
tail -n 3 gpurun_out/r2z_bench.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2z_bench.json").read().strip().splitlines()[-1])
print({k: l[k] for k in ("value", "ms_per_step", "gpu_launches")}, l["e2e"]["value"], l["roofline"]["frac"], l.get("cpu_baseline",{}).get("value"))
PY
