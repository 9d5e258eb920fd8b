#!/usr/bin/env python
"""Per-phase share of executed warp instructions and of stall samples (= time) for one kernel launch.

    python tools/ncu_phases.py sass.csv disasm.sass <kernel-name-substring> source.cu

Phases are the `// ---- title` comment markers of the source (plus the helper functions above the
kernel); instructions inlined from toolkit headers count towards the last .cu line before them.
"""
import csv
import importlib.util
import os
import re
import sys

spec = importlib.util.spec_from_file_location("ncu_lines", os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_lines.py"))
nl = importlib.util.module_from_spec(spec)
spec.loader.exec_module(nl)


def main():
    sass_csv, disasm, kernel, source = sys.argv[1:5]
    src = open(source).read().split("\n")
    marks = []
    for i, ln in enumerate(src, 1):
        m = re.match(r"\s*// ---- (.*)", ln)
        if m:
            marks.append((i, m.group(1)[:48]))
        elif re.match(r"(__device__|__global__|template <)", ln) and not (marks and marks[-1][0] == i - 1):
            marks.append((i, "fn: " + src[i].strip()[:44] if ln.startswith("template") else "fn: " + ln.strip()[:44]))
    marks.sort()
    bounds = [(name, a, (marks[k + 1][0] - 1 if k + 1 < len(marks) else len(src))) for k, (a, name) in enumerate(marks)]
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
    assert len(dis) == len(body), (len(dis), len(body))
    stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    agg = {name: [0, 0, {}, 0] for name, _, _ in bounds}
    ti = ts = 0
    for (off, line, text), r in zip(dis, body):
        inst = int(r[col["Instructions Executed"]] or 0)
        smp = int(r[col["# Samples"]] or 0)
        ti += inst
        ts += smp
        for name, a, b in bounds:
            if isinstance(line, int) and a <= line <= b:
                g = agg[name]
                g[0] += inst
                g[1] += smp
                g[3] += 1
                for c in stalls:
                    v = int(r[col[c]] or 0)
                    if v:
                        g[2][c] = g[2].get(c, 0) + v
                break
    print("total warp instructions %d, samples %d, static SASS %d" % (ti, ts, len(dis)))
    print("%-50s %6s %6s %6s  top stalls" % ("phase", "inst%", "time%", "sass"))
    for name, a, b in bounds:
        i, s, st, n = agg[name]
        if not n:
            continue
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print("%-50s %5.1f%% %5.1f%% %6d  %s" % (name, 100.0 * i / max(ti, 1), 100.0 * s / max(ts, 1), n,
                                                 " ".join("%s=%d" % (k[6:], v) for k, v in top)))


if __name__ == "__main__":
    main()
