"""Per-source-line view of `ncu --page source --print-source cuda,sass --csv`: executed warp instructions, stall samples.

    python tools/ncu_src_lines.py gpurun_out/r02b_sdf_src.csv [top]
"""
import csv
import sys
import collections

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# the file is a sequence of blocks: "File Path",..; "Function Name",..; header; rows
per = collections.OrderedDict()
hdr = None
fpath = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or len(r) < 40:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    if r[ix["Address"]] != "-":   # the sass rows under a source line (empty line number) are skipped above
        continue
    key = (fpath, ln)
    d = per.setdefault(key, collections.Counter())
    d["src"] = r[1]
    for name in ("# Samples", "Instructions Executed", "stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst",
                 "stall_branch_resolving", "stall_math", "stall_mio", "stall_lg", "stall_not_selected", "stall_selected"):
        try:
            d[name] += int(r[ix[name]])
        except (ValueError, KeyError):
            pass
tot_s = sum(d["# Samples"] for d in per.values())
tot_i = sum(d["Instructions Executed"] for d in per.values())
print("total samples", tot_s, "warp instructions", tot_i)
for name in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_branch_resolving", "stall_math",
             "stall_mio", "stall_lg", "stall_not_selected", "stall_selected"):
    print(f"  {name:24s} {100.0 * sum(d[name] for d in per.values()) / max(tot_s, 1):5.1f}%")
print("-- by line (file order), lines with >= 0.4% of samples or instructions")
for (f, ln), d in per.items():
    ps, pi = 100.0 * d["# Samples"] / tot_s, 100.0 * d["Instructions Executed"] / tot_i
    if ps >= 0.4 or pi >= 0.4:
        print(f"{f}:{ln:5d} samp {ps:5.1f}% inst {pi:5.1f}%  bar {100.0*d['stall_barrier']/tot_s:4.1f} lsb {100.0*d['stall_long_sb']/tot_s:4.1f} "
              f"ssb {100.0*d['stall_short_sb']/tot_s:4.1f} wait {100.0*d['stall_wait']/tot_s:4.1f} noi {100.0*d['stall_no_inst']/tot_s:4.1f} | {str(d['src']).strip()[:90]}")
