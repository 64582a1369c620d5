"""One-screen summary of an .ncu-rep (first kernel matching a regex): duration, DRAM bytes, issue rate, occupancy, stalls.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel regex]
"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "lts__t_bytes.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if pat and not pat.search(d.get("Kernel Name", "")):
        continue
    print("kernel:", d["Kernel Name"][:100])
    u = dict(zip(hdr, units))
    for k in KEYS:
        if k in d and d[k] != "":
            print("  %-62s %s %s" % (k, d[k], u.get(k, "")))
    stalls = sorted(((float(d[k]), k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and d[k]), reverse=True)
    print("  stalls per issue:", ", ".join("%s %.2f" % (k.split("stalled_")[1].split("_per_issue")[0], v) for v, k in stalls[:6]))
    break
