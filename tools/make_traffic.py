#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` raw-page CSV exports: per-launch dram__bytes_read.sum + dram__bytes_write.sum of the
kernels bench.py's roofline objects name (bench.py reads the file; nothing is typed into bench.py).
Usage: python tools/make_traffic.py key=csv[:kernel-regex[:grid]] ...   (median over the matching launches)"""
import csv, io, json, os, re, statistics, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    rows = list(csv.reader(io.StringIO(open(path).read())))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    return col, units, rows[2:]


def main(argv):
    out_path = os.path.join(ROOT, "profiles", "traffic.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for spec in argv:
        key, rest = spec.split("=", 1)
        parts = rest.split(":")
        path, rx, grid = parts[0], (parts[1] if len(parts) > 1 else "."), (parts[2] if len(parts) > 2 else None)
        col, units, rows = load(path)
        vals, durs = [], []
        for r in rows:
            if not re.search(rx, r[col["Kernel Name"]]):
                continue
            if grid and r[col["Grid Size"]].replace(" ", "") != grid.replace(" ", ""):
                continue
            b = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                b += float(r[col[m]].replace(",", "")) * UNIT[units[col[m]]]
            vals.append(b)
            durs.append(float(r[col["gpu__time_duration.sum"]].replace(",", "")))
        if not vals:
            raise SystemExit(f"{key}: no launch matches {rx} {grid} in {path}")
        out[key] = {"bytes": statistics.median(vals), "launches": len(vals), "kernel": rx, "grid": grid, "duration_us_under_ncu": statistics.median(durs),
                    "source": os.path.relpath(path, ROOT), "how": "ncu --set full --clock-control none, cold-cache replay of each launch: median of dram__bytes_read.sum + dram__bytes_write.sum"}
        print(key, out[key])
    json.dump(out, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1:])
