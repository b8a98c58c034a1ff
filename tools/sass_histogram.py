"""SASS opcode histogram of every kernel in the built library (cuobjdump -sass; no GPU needed).

    python tools/sass_histogram.py [lib.so] > profiles/r02_sass_opcodes.csv

Columns: total instructions and the mnemonics that identify the Blackwell paths (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, UTMALDG/UTMASTG = tensor-map TMA, SYNCS = mbarrier) and the hot integer
ops of the penetration kernels (VABSDIFF4, IDP.4A, REDUX, ATOMS, POPC/FLO).
"""
import collections
import re
import subprocess
import sys
import os

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ihmr_b200", "_lib", "libihmr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FFMA", "FMUL", "FADD", "MUFU", "VABSDIFF4", "IDP", "REDUX",
        "ATOMS", "ATOMG", "RED", "POPC", "FLO", "VOTE", "SHFL", "BAR", "LDG", "STG", "LDS", "STS", "LDL", "STL"]
name, per = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "").replace("ihmr::", "")
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        op = m.group(1)
        per[name]["total"] += 1
        per[name][op] += 1
print("kernel,total," + ",".join(KEYS))
for k, c in per.items():
    print(k.replace(",", ";") + "," + str(c["total"]) + "," + ",".join(str(c[x]) for x in KEYS))
