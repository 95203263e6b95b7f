"""Key per-kernel metrics of an ncu --set full report as a markdown table:
   python tools/ncu_summary.py gpurun_out/r01_gemm.ncu-rep > profiles/r01_gemm_ncu.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
print("| " + " | ".join(f"{n} [{units[i]}]" if units[i] else n for i, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    cells = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = "`" + v.split("(")[0][-48:] + "`"
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
