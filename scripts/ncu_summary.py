"""Summarise an .ncu-rep (read here, no GPU): per-launch key metrics + warp-stall breakdown.  Usage: ncu_summary.py REP [ID]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
for r in rows[2:]:
    if len(sys.argv) > 2 and r[idx["ID"]] != sys.argv[2]:
        continue
    print("=" * 100)
    print(r[idx["Kernel Name"]][:160])
    for k in KEYS:
        if k in idx:
            print(f"  {k:72s} {r[idx[k]]:>18s} {units[idx[k]]}")
    stalls = [(h, float(r[i] or 0)) for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    stalls.sort(key=lambda t: -t[1])
    print("  warp stalls per issue-active (top):", ", ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}={v:.2f}" for h, v in stalls[:7]))
    pipes = [(h, float(r[i] or 0)) for h, i in idx.items() if h.startswith("sm__inst_executed_pipe_") and h.endswith(".avg.pct_of_peak_sustained_active")]
    pipes.sort(key=lambda t: -t[1])
    print("  busiest pipes (% of peak):", ", ".join(f"{h[len('sm__inst_executed_pipe_'):-len('.avg.pct_of_peak_sustained_active')]}={v:.1f}" for h, v in pipes[:8]))
