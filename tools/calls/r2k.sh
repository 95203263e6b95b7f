python -m pytest tests -m gpu -q > gpurun_out/r2k_tests.log 2>&1; tail -3 gpurun_out/r2k_tests.log
python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -2 gpurun_out/r2k_bench.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2k_ref.json 2> gpurun_out/r2k_ref.err; tail -2 gpurun_out/r2k_ref.err
python -c "
import json; d=json.load(open('gpurun_out/r2k_ref.json')); print({k:d[k] for k in ('value','ms_per_step','timed_s','frames_timed','wall_s')})"
for F in 64 8; do
CMD="python tools/profile_step.py --frames $F --queries 500000"
$CMD > gpurun_out/r2k_plain$F.log 2>&1 && ncu --nvtx --nvtx-include "profiled_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_f$F.csv $CMD > gpurun_out/r2k_ncu$F.log 2>&1
tail -1 gpurun_out/r2k_plain$F.log gpurun_out/r2k_ncu$F.log
done
python __graft_entry__.py --smoke 2>&1 | tail -2
