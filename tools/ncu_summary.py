"""Condensed text summary of an .ncu-rep (raw page): duration, DRAM bytes, tensor pipe, occupancy, stall mix."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__t_bytes.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
for k in want:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:70s} {vals[i]:>20s} {units[i]}")
stalls = [(hdr[i], float(vals[i])) for i in range(len(hdr))
          if hdr[i].startswith("smsp__average_warp") and "issue_stalled" in hdr[i] and hdr[i].endswith("_per_issue_active.ratio") or
          (hdr[i].startswith("smsp__average_warps_issue_stalled") and hdr[i].endswith(".ratio"))]
for k, v in sorted(stalls, key=lambda kv: -kv[1])[:8]:
    print(f"  stall {k:90s} {v:8.2f}")
