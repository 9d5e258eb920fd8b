#!/usr/bin/env python
"""Aggregate an ncu SASS source page per CUDA source line.

    ncu -i prof.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all libevgsim.so ; nvdisasm -g -c evg_kernels.sm_100a.cubin > k.sass
    python tools/ncu_lines.py sass.csv k.sass evg_step_kernel [source.cu]

Joins on the instruction offset inside the kernel (ncu lists every SASS instruction of the kernel in
order; nvdisasm -g interleaves `//## File ..., line N` markers).  Prints, per source line, the warp
instructions executed, their share, average active threads and stall samples, for one launch.
"""
import csv
import re
import sys
from collections import defaultdict


def sass_lines(path, kernel):
    """[(offset, line_no, text)] for the function whose mangled name contains `kernel`."""
    out, cur_line, inside = [], None, False
    for ln in open(path):
        if ln.startswith("\t.section\t.text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m:
            f = m.group(1).rsplit("/", 1)[-1]
            if f.endswith(".cu"):
                cur_line = int(m.group(2))  # instructions inlined from toolkit headers keep the last .cu line
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), cur_line, m.group(2).strip()))
    return out


def main():
    sass_csv, disasm, kernel = sys.argv[1:4]
    src = open(sys.argv[4]).read().split("\n") if len(sys.argv) > 4 else None
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {n: i for i, n in enumerate(hdr)}
    body = []
    for r in rows[hi + 1:]:
        if r and r[0] == "Kernel Name":
            break  # first launch only
        if len(r) == len(hdr):
            body.append(r)
    dis = sass_lines(disasm, kernel)
    assert len(dis) == len(body), (len(dis), len(body))
    per = defaultdict(lambda: [0, 0, 0, 0])  # inst, thread inst, samples, n sass
    total = 0
    for (off, line, text), r in zip(dis, body):
        inst = int(r[col["Instructions Executed"]] or 0)
        thr = int(r[col["Thread Instructions Executed"]] or 0)
        smp = int(r[col["# Samples"]] or 0)
        p = per[line]
        p[0] += inst; p[1] += thr; p[2] += smp; p[3] += 1
        total += inst
    tot_smp = sum(p[2] for p in per.values())
    print("total warp instructions: %d, samples: %d" % (total, tot_smp))
    if len(sys.argv) > 5:
        import json
        phases(per, total, json.load(open(sys.argv[5])))
    print("%5s %12s %6s %6s %8s %6s  %s" % ("line", "warp_inst", "share", "thr/in", "samples", "smp%", "source"))
    for line, p in sorted(per.items(), key=lambda kv: -kv[1][0])[:70]:
        s = src[line - 1].strip()[:90] if src and isinstance(line, int) and line <= len(src) else ""
        print("%5s %12d %5.1f%% %6.1f %8d %5.1f%%  %s" % (line, p[0], 100.0 * p[0] / max(total, 1), p[1] / max(p[0], 1), p[2],
                                                          100.0 * p[2] / max(tot_smp, 1), s))


def phases(per, total, bounds):
    """bounds: [(name, first_line, last_line)] -> instruction share per phase."""
    print("--- per phase")
    for name, a, b in bounds:
        v = sum(p[0] for ln, p in per.items() if isinstance(ln, int) and a <= ln <= b)
        print("%-28s %12d %5.1f%%" % (name, v, 100.0 * v / max(total, 1)))


if __name__ == "__main__":
    main()
