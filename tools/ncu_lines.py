#!/usr/bin/env python
"""Aggregate ncu warp-stall samples per CUDA source line.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_lines.py src.csv [top_n]
"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr, cur, cur_file = None, None, None
    per_line, text = {}, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
            stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[0] != "":
            cur = (cur_file, int(r[0]))
            text[cur] = r[1]
            continue
        if r[2] in ("...", ""):
            continue
        try:
            s = int(r[si])
        except ValueError:
            continue
        if cur is None:
            cur = ("?", 0)
            text[cur] = "(no line info)"
        d = per_line.setdefault(cur, [0, {}, 0])
        d[0] += s
        d[2] += int(r[ii])
        for i in stall_cols:
            if r[i] not in ("", "0", "-"):
                d[1][hdr[i]] = d[1].get(hdr[i], 0) + int(r[i])
    tot = sum(d[0] for d in per_line.values())
    print("total samples", tot)
    for ln, d in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        st = sorted(d[1].items(), key=lambda kv: -kv[1])[:3]
        print(f"{ln[0]}:{ln[1]:<5d} {d[0]:>8d} {100 * d[0] / tot:5.1f}%  inst={d[2]:<11d} {text[ln].strip()[:80]:<80s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
