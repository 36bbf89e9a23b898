#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass` export (gzipped or not): instruction and stall-sample shares
by opcode and by 60-instruction region.  Usage: python tools/sass_profile.py gpurun_out/x_source_sass.csv.gz"""
import collections
import csv
import gzip
import re
import sys


def main(path, chunk=60):
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(f))
    hdr, data = rows[1], rows[2:]
    ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[ia]) for r in data) or 1
    tots = sum(int(r[ist]) for r in data) or 1
    print(rows[0][1][:120])
    print("warp instructions", tot, " stall samples", tots, " sass lines", len(data))
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        o = m.group(2).split(".")[0] if m else "?"
        op[o] += int(r[ia]); ops[o] += int(r[ist])
    for o, c in op.most_common(25):
        print(f"  {o:10s} {c:10d} {100 * c / tot:5.1f}%   stall samples {100 * ops[o] / tots:5.1f}%")
    print()
    for i in range(0, len(data), chunk):
        c = sum(int(r[ia]) for r in data[i:i + chunk]); s = sum(int(r[ist]) for r in data[i:i + chunk])
        print(f"  sass {i:5d}-{i + chunk - 1:5d}: inst {100 * c / tot:5.1f}%  stalls {100 * s / tots:5.1f}%   {data[i][isrc].strip()[:70]}")


if __name__ == "__main__":
    main(sys.argv[1])
