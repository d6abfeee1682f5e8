#!/usr/bin/env python
"""Stall samples of an `ncu --set full --import-source on` report aggregated per CUDA source line."""
import csv
import subprocess
import sys


def main(path, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source",
                          "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr, agg = "?", None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-":  # a source line (its SASS rows follow)
            try:
                v = float(r[hdr.index("# Samples")])
            except ValueError:
                continue
            key = (fname, r[0], r[1].strip()[:100])
            agg[key] = agg.get(key, 0.0) + v
    tot = sum(agg.values()) or 1.0
    for (f, ln, s), v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        print("%5.1f%%  %s:%s  %s" % (100 * v / tot, f, ln, s))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
