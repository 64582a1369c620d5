"""Per CUDA-source-line totals of an ncu report (needs -lineinfo and --import-source on):
    python tools/ncu_lines.py report.ncu-rep [units] [top]
prints instructions executed (per `units` work items) and the share of stall samples of the hottest lines."""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    inst, samp, text = defaultdict(float), defaultdict(float), {}
    fname, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ie, isa = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].isdigit():
            continue
        key = (fname, int(r[0]))
        text[key] = r[1].strip()
        try:
            inst[key] += float(r[ie].replace(",", "") or 0)
            samp[key] += float(r[isa].replace(",", "") or 0)
        except ValueError:
            pass
    ti, ts = sum(inst.values()), sum(samp.values()) or 1.0
    print("total warp instructions %.0f = %.1f per unit; %d stall samples" % (ti, ti / units, ts))
    for k in sorted(inst, key=lambda k: -inst[k])[:top]:
        print("%8.1f instr/unit %5.1f%% of samples  %s:%d  %s" % (inst[k] / units, 100 * samp[k] / ts, k[0], k[1], text[k][:100]))


if __name__ == "__main__":
    main()
