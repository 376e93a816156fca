"""Summarise an ncu report (raw page) into the handful of counters this project is judged on.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_write.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
]
out = []
for li, vals in enumerate(rows[2:]):
    out.append(f"--- launch {li}: {vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out.append(f"{k:90s} {vals[i]:>18s} {units[i]}")
    stalls = [(float(vals[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and vals[i]]
    tot = sum(s for s, _ in stalls) or 1.0
    for s, h in sorted(stalls, reverse=True)[:8]:
        out.append(f"stall {h[len('smsp__pcsamp_warps_issue_stalled_'):]:40s} {s:12.0f}  {100 * s / tot:5.1f} %")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
