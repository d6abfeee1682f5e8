#!/usr/bin/env python
"""Top stall locations of an `ncu --set full --import-source on` report (SASS view)."""
import csv
import subprocess
import sys


def main(path, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    si = hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for n, r in enumerate(rows[2:]):
        try:
            v = float(r[si])
        except (ValueError, IndexError):
            continue
        stalls = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        data.append((v, n, r[1].strip()[:70], stalls))
    tot = sum(d[0] for d in data) or 1.0
    print("total samples %d, %d instructions" % (tot, len(data)))
    for v, n, s, st in sorted(data, reverse=True)[:top]:
        print("%5.1f%%  #%-5d %-70s %s" % (100 * v / tot, n, s,
                                          " ".join("%s:%d" % (b, a) for a, b in st if a)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
