#!/usr/bin/env python
"""Opcode histogram per kernel of libamps_b200.so (cuobjdump -sass): what proves which hardware paths the kernels use
(UBLKCP = TMA bulk copy, SYNCS = mbarrier, FFMA2/FMUL2/FADD2 = packed fp32x2, UTC*MMA / HMMA = tensor cores (none expected:
no dense contraction on this path)).  usage: python tools/sass_digest.py > profiles/r2_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gr_amps_b200", "libamps_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, hist = None, collections.OrderedDict()
arch = set(re.findall(r"arch = (sm_\w+)", txt))
for line in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("# SASS digest of gr_amps_b200/libamps_b200.so -- architectures:", ", ".join(sorted(arch)))
print("# marker opcodes: UBLKCP (TMA bulk copy), SYNCS (mbarrier), FFMA2/FMUL2/FADD2 (packed fp32x2), LDS/STS, REDUX/SHFL/VOTE (warp), UTC*MMA/HMMA (tensor cores)")
mark = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "LDS", "STS", "LDG", "STG", "LDCU", "SHFL", "VOTE", "ATOMG", "RED", "BAR", "HMMA", "I2F", "MUFU"]
for k, c in hist.items():
    total = sum(c.values())
    tens = sum(v for o, v in c.items() if o.startswith("UTC") or o in ("HMMA", "IMMA", "DMMA", "HGMMA"))
    print("\n%s\n  %d instructions; tensor-core ops: %d" % (demangle(k), total, tens))
    print("  " + "  ".join("%s %d" % (o, c[o]) for o in mark if c.get(o)))
    rest = [(o, v) for o, v in c.most_common(12)]
    print("  top: " + ", ".join("%s %d" % kv for kv in rest))
