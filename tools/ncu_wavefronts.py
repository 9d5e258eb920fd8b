#!/usr/bin/env python
"""Shared-memory wavefronts and global L1 tag requests per source line (the L1TEX data pipe's load).

    python tools/ncu_wavefronts.py sass.csv k.sass <kernel-substring> source.cu [matches]
"""
import csv
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as nl


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main():
    sass_csv, disasm, kernel, source = sys.argv[1:5]
    matches = float(sys.argv[5]) if len(sys.argv) > 5 else 262144.0
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {n: i for i, n in enumerate(hdr)}
    body = []
    for r in rows[hi + 1:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr):
            body.append(r)
    dis = nl.sass_lines(disasm, kernel)
    src = open(source).read().split("\n")
    agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    tot = [0.0, 0.0, 0.0, 0.0]
    for (off, line, text), r in zip(dis, body):
        v = (num(r[col["L1 Wavefronts Shared"]]), num(r[col["L1 Wavefronts Shared Excessive"]]), num(r[col["L1 Tag Requests Global"]]),
             num(r[col["L2 Theoretical Sectors Global"]]))
        for k in range(4):
            agg[line][k] += v[k]
            tot[k] += v[k]
    print("per match: shared wavefronts %.1f (excessive %.1f), global tag requests %.1f, global sectors %.1f" % tuple(t / matches for t in tot))
    for line, a in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][2]))[:int(os.environ.get("TOPN", "30"))]:
        print("%5s sh=%6.2f exc=%6.2f gtag=%6.2f gsec=%6.2f  %s" % (line, a[0] / matches, a[1] / matches, a[2] / matches, a[3] / matches,
                                                                     src[line - 1].strip()[:105] if line else ""))


if __name__ == "__main__":
    main()
