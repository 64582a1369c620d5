"""Aggregate `ncu --page source --print-source cuda,sass --csv` output by CUDA source line.

    ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv --kernel-id ::regex:<kernel>:<n> > src.csv
    python tools/ncu_hotspots.py src.csv <units per launch> [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fname, out, hdr = "", [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and r[0].isdigit() and len(r) > ie and r[2] == "-":     # a CUDA line row (SASS rows carry an address)
        v, s = float(r[ie] or 0), float(r[isamp] or 0)
        if v > 0 or s > 0:
            out.append((v, s, fname, r[0], r[1]))
tot, tots = sum(o[0] for o in out), sum(o[1] for o in out) or 1.0
print("total warp-instructions %d, per unit %.1f" % (tot, tot / units))
for v, s, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%-14s %5s i/u=%8.1f (%4.1f%%) samp=%4.1f%%  %s" % (f[:14], ln, v / units, 100 * v / tot, 100 * s / tots, src.strip()[:105]))
