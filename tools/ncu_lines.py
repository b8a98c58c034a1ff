"""Summarise an ncu report per CUDA source line / per region.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]
Regions are delimited by the '// ----' comment markers and function headers of the source file."""
import csv, subprocess, sys, re
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
path = rows[0][1]; hdr = rows[2]
ci, li, ti = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
lines = [r for r in rows[3:] if r and r[0].isdigit() and r[ci].replace('.', '').isdigit()]
tot = sum(int(r[ci]) for r in lines); tots = sum(int(r[li]) for r in lines)
print(f"{path}: total warp-inst {tot:.3e}, samples {tots}")
src = [r[1] for r in lines]
# regions: a new region starts at lines that look like markers
marks = []
for r in lines:
    t = r[1].strip()
    if t.startswith('// ----') or re.match(r'^(__device__|__global__|static|template)', t):
        marks.append((int(r[0]), t[:70]))
marks.append((10**9, 'end'))
agg = []
for (a, name), (b, _) in zip(marks[:-1], marks[1:]):
    sel = [r for r in lines if a <= int(r[0]) < b]
    i = sum(int(r[ci]) for r in sel); s_ = sum(int(r[li]) for r in sel); t = sum(int(r[ti]) for r in sel)
    if i * 1000 > tot or s_ * 1000 > tots:
        agg.append((i, s_, t, a, name))
for i, s_, t, a, name in agg:
    print(f"  L{a:<4} inst {i*100/tot:5.1f}%  samples {s_*100/tots:5.1f}%  lanes {t/max(i,1):4.1f} | {name}")
print("top lines by samples:")
for r in sorted(lines, key=lambda r: -int(r[li]))[:topn]:
    print(f"  {r[0]:>4} inst {int(r[ci])*100/tot:5.1f}% samp {int(r[li])*100/tots:5.1f}% lanes {int(r[ti])/max(1,int(r[ci])):4.1f} | {r[1].strip()[:95]}")
