"""Compact per-kernel summary of an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}
base = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_lsu.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
stall = [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("==", r[col["Kernel Name"]][:60])
    for k in base:
        if k in col:
            print(f"   {k} = {r[col[k]]} {units[col[k]]}")
    st = sorted(((float(r[col[k]] or 0), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k in stall), reverse=True)[:6]
    print("   stalls(warps/issue):", ", ".join(f"{n}={v:.2f}" for v, n in st))
