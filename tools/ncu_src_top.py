#!/usr/bin/env python
"""Top-stalled SASS instructions of an `ncu --page source --csv` export:  python tools/ncu_src_top.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def f(r, k):
    try:
        return int(r[ix[k]] or 0)
    except ValueError:
        return 0


data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(f(r, "# Samples") for r in data)
print("samples", tot, " instructions", sum(f(r, "Instructions Executed") for r in data))
agg = {s: sum(f(r, s) for r in data) for s in stall}
print("  ".join(f"{s[6:]} {100 * v / max(tot, 1):.1f}%" for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:n]:
    st = sorted(((f(r, s), s[6:]) for s in stall), reverse=True)[:2]
    print(str(f(r, "# Samples")).rjust(6), str(f(r, "Instructions Executed")).rjust(9), r[ix["Source"]].strip()[:70].ljust(70), st)
