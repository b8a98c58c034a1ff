"""SASS instructions per source line of one kernel (from -lineinfo): where the code size goes.

    python tools/sass_lines.py file.o k_sdf_dirILb0 [top]
"""
import collections
import re
import subprocess
import sys
import tempfile
import os

obj, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
cubin = os.path.join(d, [f for f in os.listdir(d) if f.endswith(".cubin")][0])
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur, line, active = None, None, False
per = collections.Counter()
total = 0
for l in out.splitlines():
    m = re.match(r"\s*//-+ \.text\.(\S+)", l)
    if m:
        active = kern in m.group(1)
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l):
        per[line] += 1
        total += 1
print("total", total)
for k, v in per.most_common(top):
    print(v, k)
