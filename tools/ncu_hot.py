#!/usr/bin/env python
"""Top stalled SASS instructions of each kernel in an .ncu-rep (source page).  Usage: ncu_hot.py rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for kb in blocks[1:]:
    lines = kb.splitlines()
    print("==", lines[0][:100])
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    h = rows[0]
    iS, iSamp, iEx = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    body = [r for r in rows[1:] if len(r) == len(h)]
    tot = sum(int(r[iSamp] or 0) for r in body)
    print("total samples", tot, "instructions", len(body))
    ranked = sorted(enumerate(body), key=lambda x: -int(x[1][iSamp] or 0))[:topn]
    for idx, r in ranked:
        st = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"  #{idx:5d} {int(r[iSamp]):7d} {100*int(r[iSamp])/max(tot,1):5.1f}%  ex={r[iEx]:>9s} {r[iS].strip()[:70]:70s} {st}")
