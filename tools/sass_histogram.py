#!/usr/bin/env python
"""SASS opcode histogram per kernel of m3pc_b200/libm3pc.so (cuobjdump -sass): the Blackwell-specific opcodes that prove which
hardware path a kernel uses (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store /
reduce, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = mma.sync, FFMA = fp32 FMA), registers and spills from the ptxas logs.
    python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt"""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "m3pc_b200", "libm3pc.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "HMMA", "FFMA", "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "BAR"]
kern, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::|m3pc::", "", kern)
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                hist[kern][o] += 1
regs = {}
for log in glob.glob(os.path.join(ROOT, "m3pc_b200", "csrc", "build", "*.ptxas.log")):
    txt = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?(\d+) bytes spill stores, (\d+) bytes spill loads.*?Used (\d+) registers", txt, re.S):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|m3pc::", "", name)
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        regs[name] = (int(m.group(4)), int(m.group(2)) + int(m.group(3)))
print("# SASS opcode counts per kernel of m3pc_b200/libm3pc.so (sm_100a), static instruction counts; regs / spill bytes from ptxas -v")
print(f"{'kernel':<50} {'instr':>6} {'regs':>4} {'spill':>5} " + " ".join(f"{o:>8}" for o in OPS))
for k, c in hist.items():
    r = regs.get(k, ("", ""))
    print(f"{k[:50]:<50} {c['_total']:>6} {r[0]:>4} {r[1]:>5} " + " ".join(f"{c[o]:>8}" for o in OPS))
