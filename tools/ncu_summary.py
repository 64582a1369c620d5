"""Key numbers of an ncu report (run where ncu is installed, no GPU needed):

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--source N] > profiles/r02_ncu_<kernel>_summary.txt

Prints, per captured launch: duration, DRAM bytes, throughput percentages, issue activity, occupancy, registers, the
warp-stall breakdown, and (with --source N) the N source lines with the most sampled stalls / executed instructions."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_average_branch_targets_threads_uniform.pct",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def main():
    rep = sys.argv[1]
    nsrc = int(sys.argv[sys.argv.index("--source") + 1]) if "--source" in sys.argv else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for r in data:
        print("=== launch %s: %s  grid %s block %s" % (r[hdr.index("ID")], r[hdr.index("Kernel Name")][:90],
                                                        r[hdr.index("Grid Size")] if "Grid Size" in hdr else "?", r[hdr.index("Block Size")] if "Block Size" in hdr else "?"))
        for k in KEYS:
            if k in hdr:
                print("  %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        stalls = [(float(r[i].replace(",", "") or 0), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") or h.startswith("smsp__average_warps_issue_stalled")]
        stalls = [(v, h) for v, h in stalls if v > 0]
        for v, h in sorted(stalls, reverse=True)[:10]:
            print("  stall %-64s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")[:64], v))
    if nsrc:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        hi = None
        for i, r in enumerate(rows):
            if "Source" in r and any("Samples" in c or "Executed" in c for c in r):
                hi = i
                break
        if hi is not None:
            h = rows[hi]
            si = h.index("Source")
            cand = [c for c in h if "Sampling" in c or "Samples" in c]
            ei = [i for i, c in enumerate(h) if c.startswith("# Warp Instructions Executed") or c == "Warp Instructions Executed"]
            ci = h.index(cand[0]) if cand else None
            body = [r for r in rows[hi + 1:] if len(r) == len(h)]
            def num(x):
                try:
                    return float(x.replace(",", ""))
                except ValueError:
                    return 0.0
            if ci is not None:
                tot = sum(num(r[ci]) for r in body) or 1.0
                print("--- top %d source lines by %s (share of samples)" % (nsrc, h[ci]))
                for r in sorted(body, key=lambda r: -num(r[ci]))[:nsrc]:
                    print("  %5.1f%%  %s%s" % (100 * num(r[ci]) / tot, (("instr %s  " % r[ei[0]]) if ei else ""), r[si].strip()[:150]))


if __name__ == "__main__":
    main()
