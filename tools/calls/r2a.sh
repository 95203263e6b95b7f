python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; tail -3 gpurun_out/r2a_tests.log
for F in 8 16 32 64; do
python bench.py --frames-per-gpu $F --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2a_bench_f$F.json 2> gpurun_out/r2a_bench_f$F.err
done
CMD="python tools/profile_step.py --frames 8 --queries 500000"
ncu --nvtx --nvtx-include "profiled_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_f8.csv $CMD > gpurun_out/r2a_ncu8.log 2>&1
tail -2 gpurun_out/r2a_ncu8.log; ls -la gpurun_out | tail -8
