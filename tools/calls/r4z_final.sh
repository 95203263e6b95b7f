#!/bin/bash
# Round-2 last session: full GPU suite, smoke, the two new kernels' stand-alone timings, the full bench line, the launch
# list of one 64-frame step and --set full captures of the boundary / conv_in kernels (each after its plain run).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r4z_tests.log 2>&1; tail -n 2 gpurun_out/r4z_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r4z_smoke.log 2>&1; tail -n 1 gpurun_out/r4z_smoke.log
timeout 100 python tools/gpu_time_boundary.py > gpurun_out/r4z_boundary_time.log 2>&1; cat gpurun_out/r4z_boundary_time.log
timeout 100 python tools/gpu_time_conv_in.py > gpurun_out/r4z_convin_time.log 2>&1; cat gpurun_out/r4z_convin_time.log
timeout 900 python bench.py > gpurun_out/r4z_bench.json 2> gpurun_out/r4z_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r4z_bench.json").read().strip().splitlines()[-1])
    print({k: l[k] for k in ("value", "ms_per_step", "gpu_launches")}, l["e2e"]["value"], l["roofline"]["frac"])
    print([(r["family"], round(r["achieved"], 1), round(r["frac"], 3)) for r in l.get("rooflines", [])])
    print("b1", l.get("latency_b1", {}).get("ms_per_frame"))
except Exception as e:
    print("bench line unreadable:", e)
PY
CMD="python tools/profile_step.py --frames 64 --queries 500000"
timeout 200 $CMD > gpurun_out/r4z_plain64.log 2>&1 && \
timeout 400 ncu --nvtx --nvtx-include "profiled_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r04_launches_f64.csv $CMD > gpurun_out/r4z_ncu64a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:"boundary_kernel|conv_in_mma_kernel" -c 4 -f -o gpurun_out/r04_boundary_convin $CMD > gpurun_out/r4z_ncu64b.log 2>&1
tail -n 1 gpurun_out/r4z_plain64.log gpurun_out/r4z_ncu64a.log gpurun_out/r4z_ncu64b.log
gzip -f -k gpurun_out/r04_launches_f64.csv
