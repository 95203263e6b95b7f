"""First-tile phase stamps of the pair GEMMs at the bench shapes (M = 32768): python tools/gemm_phases_big.py"""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import runpy
mod = runpy.run_path(os.path.join(os.path.dirname(__file__), "gemm_phases.py"), run_name="gemm_phases_lib")
run = mod["run"]
for (M, N, K, mode) in [(32768, 4096, 512, 2), (32768, 1536, 512, 0), (32768, 512, 2048, 1), (32768, 512, 512, 1)]:
    run(M, N, K, mode, 0, reps=5)
