"""SASS mnemonic counts per kernel of the built libdqa.so (cuobjdump -sass): which pipes each kernel uses
(DMMA = FP64 tensor tiles, DFMA/DMUL/DADD = FP64 vector, SHFL, LDS/STS, LDG/STG, BAR, UBLKCP/UTMALDG = bulk/TMA copies).
usage: python tools/sass_mix.py > profiles/r02_sass_instruction_mix.csv   (CPU only)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "perceptron_dqa_b200", "libdqa.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["DMMA", "DFMA", "DMUL", "DADD", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "ATOM", "RED", "UBLKCP", "UTMALDG", "IMAD.MOV", "total"]
counts, name = collections.OrderedDict(), None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and name:
        op = m.group(1)
        counts[name]["total"] += 1
        for k in keys[:-1]:
            if op.startswith(k):
                counts[name][k] += 1
print("# final build of round 2 (sm_100a), cuobjdump -sass perceptron_dqa_b200/libdqa.so; static instruction counts per kernel")
print("kernel," + ",".join(keys))
for n, c in counts.items():
    print(n + "," + ",".join(str(c[k]) for k in keys))
