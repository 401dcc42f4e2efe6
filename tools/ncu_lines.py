#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line.

usage: ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass --kernel-name K > k.csv
       python tools/ncu_lines.py k.csv [top]
Prints the source lines with the most executed warp instructions and their stall samples."""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    agg = defaultdict(lambda: [0, 0, 0, ""])   # (file,line) -> inst, thread inst, samples, text
    fpath = None
    hdr = None
    cur_line = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            hdr = None
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst = hdr.index("Instructions Executed")
            i_tinst = hdr.index("Thread Instructions Executed")
            i_samp = hdr.index("# Samples")
            continue
        if hdr is None:
            continue
        if r[0] != "":
            cur_line = (fpath, int(r[0]))
            agg[cur_line][3] = r[1].strip()
        if len(r) > i_inst and r[i_inst] not in ("", "-"):
            try:
                agg[cur_line][0] += int(r[i_inst])
                agg[cur_line][1] += int(r[i_tinst])
                agg[cur_line][2] += int(r[i_samp])
            except ValueError:
                pass
    tot = sum(v[0] for v in agg.values())
    tots = sum(v[2] for v in agg.values())
    print(f"total warp instructions {tot}, stall samples {tots}")
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*v[0]/tot:5.1f}% inst {100*v[2]/max(tots,1):5.1f}% samp  thr/inst {v[1]/max(v[0],1):4.1f}  {f}:{ln}: {v[3][:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
