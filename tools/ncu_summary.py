"""Compact per-launch summary of an `ncu --set full` capture (read here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rN_ncu_full_<kernel>_summary.csv"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [f"launch {i}: {r[idx['Kernel Name']][:70]}" for i, r in enumerate(rows[2:])])
for m in WANT:
    if m in idx:
        w.writerow([m, units[idx[m]]] + [r[idx[m]] for r in rows[2:]])
