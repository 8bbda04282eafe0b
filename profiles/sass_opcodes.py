#!/usr/bin/env python
"""Per-kernel SASS opcode counts of libdlrm_b200.so (cuobjdump -sass), the evidence for which hardware
paths the kernels use: HMMA = legacy warp-level tensor-core MMA (mma.sync), UTC*MMA / LDTM / STTM =
tcgen05 + TMEM, UBLKCP / UTMALDG / UTMASTG = TMA bulk copies, FFMA2 = packed fp32 FMA, MATCH = match.any,
ATOMS = shared-memory atomics, REDUX/SHFL = warp reductions, LDG/STG .128 = 16-byte vectorised global access.

    python profiles/sass_opcodes.py [path/to/lib.so] > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dlrm_jl_b200", "lib", "libdlrm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["HMMA", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FFMA2", "FFMA",
         "FADD", "MATCH", "ATOMS", "ATOMG", "RED", "SHFL", "VOTE", "LDG.E.128", "STG.E.128", "LDG", "STG", "LDS", "STS",
         "LDGSTS", "BAR", "MEMBAR", "NANOSLEEP", "ACQBULK", "CCTL"]
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                kernels[cur][w] += 1


def demangle(names):
    try:
        r = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return r if len(r) == len(names) else names
    except Exception:
        return names


names = list(kernels)
pretty = demangle(names)
print(f"# SASS opcode counts per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a)")
print("# kernel | instructions | watched opcodes (count)")
tot = collections.Counter()
for n, p in zip(names, pretty):
    c = kernels[n]
    short = re.sub(r"\((?:int|bool|unsigned int)\)", "", p)      # template-argument casts
    short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", short)
    short = re.sub(r"\(.*", "", short).replace("void ", "")
    cols = ", ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{short} | {c['_total']} | {cols}")
    tot.update(c)
print("# totals: " + ", ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))
print(f"# tcgen05 (UTC*MMA/LDTM/STTM): {tot['UTCHMMA'] + tot['UTCQMMA'] + tot['UTCIMMA'] + tot['LDTM'] + tot['STTM']}; "
      f"legacy MMA (HMMA): {tot['HMMA']}; TMA bulk copies (UBLKCP): {tot['UBLKCP']}; FFMA2: {tot['FFMA2']}")
