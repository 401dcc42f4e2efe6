#!/usr/bin/env python
"""Per source line: share of one stall reason (default stall_wait) and of all stall samples.

usage: ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass --kernel-name K > k.csv
       python tools/ncu_stalls.py k.csv [stall_long_sb] [top]"""
import csv
import sys
from collections import defaultdict


def main(path, reason="stall_wait", top=25):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    i_r, i_s, i_i = hdr.index(reason), hdr.index("# Samples"), hdr.index("Instructions Executed")
    agg = defaultdict(lambda: [0, 0, 0, ""])
    fpath = cur = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] in ("Line No", "Function Name") or len(r) <= i_r:
            continue
        if r[0] != "":
            cur = (fpath, int(r[0]))
            agg[cur][3] = r[1].strip()
        try:
            agg[cur][0] += int(r[i_r] or 0)
            agg[cur][1] += int(r[i_s] or 0)
            agg[cur][2] += int(r[i_i] or 0)
        except ValueError:
            pass
    tot = max(sum(v[0] for v in agg.values()), 1)
    tots = max(sum(v[1] for v in agg.values()), 1)
    toti = max(sum(v[2] for v in agg.values()), 1)
    print(f"{reason}: {tot} of {tots} samples")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*v[0]/tot:5.1f}% {reason[6:]} {100*v[1]/tots:5.1f}% samp {100*v[2]/toti:5.1f}% inst  "
              f"{k[0]}:{k[1]}: {v[3][:90]}")


if __name__ == "__main__":
    a = sys.argv
    main(a[1], a[2] if len(a) > 2 else "stall_wait", int(a[3]) if len(a) > 3 else 25)
