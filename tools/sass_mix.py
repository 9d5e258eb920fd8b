#!/usr/bin/env python
"""Static SASS opcode mix of libevgsim's kernels (cuobjdump -sass) -> profiles/<tag>_sass_mix.txt.
Shows which instruction families each kernel uses: no tensor-core / TMA / TMEM opcodes in the step kernels (HBM-bound
integer/byte work), UTCHMMA (tcgen05.mma), UBLKCP (cp.async.bulk), LDTM (tcgen05.ld), UTCBAR / SYNCS (mbarriers) in the
policy forward."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "everglades-ai-wargame_b200/libevgsim.so"
out = sys.argv[2] if len(sys.argv) > 2 else "profiles/r2_final_sass_mix.txt"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
fn = None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        kernels[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and fn:
        kernels[fn][m.group(1)] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
special = ("UTCHMMA", "UTCBAR", "UBLKCP", "UTMALDG", "LDTM", "STTM", "SYNCS", "HMMA", "ATOMS", "RED", "REDG", "LDG.E.EF.ENL2.256", "STG.E.ENL2.256", "BAR")
with open(out, "w") as f:
    f.write(__doc__.split("\n", 1)[1] + "\n")
    for (name, c), dm in zip(kernels.items(), demangle):
        total = sum(c.values())
        fam = collections.Counter()
        for op, n in c.items():
            fam[op.split(".")[0]] += n
        f.write("%s\n  %d instructions\n  by opcode: %s\n" % (dm[:200], total, ", ".join("%s %d" % kv for kv in fam.most_common(22))))
        mem = {op: n for op, n in c.items() if re.match(r"(LD|ST|ATOM|RED|SHFL|BAR|UTC|UBLK|UTMA|SYNCS|CCTL|LDTM|STTM|HMMA|REDUX)", op)}
        f.write("  memory / warp-level / tensor forms: %s\n\n" % ", ".join("%s %d" % kv for kv in sorted(mem.items(), key=lambda kv: -kv[1])[:32]))
print("wrote", out)
