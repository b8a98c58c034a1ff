"""Summarise an ncu report (captured with --import-source on, code built with -lineinfo) per region of
one CUDA source file.  usage: python tools/ncu_lines.py report.ncu-rep kernel_regex source.cu [top_n]
Regions start at every comment line beginning with '// ----', '// (' or at function headers."""
import csv, re, subprocess, sys
rep, kre, srcpath = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[start]
ci, li, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
lines = []
for r in rows[start + 1:]:
    if r and r[0] in ("File Path", "Line No"):
        break
    if r and r[0].isdigit() and r[ci].replace(".", "").isdigit():
        num = lambda x: int(x) if x.replace('.', '').isdigit() else 0
        lines.append((int(r[0]), num(r[ci]), num(r[li]), num(r[ti]), r[1]))
tot = sum(l[1] for l in lines); tots = sum(l[2] for l in lines)
print(f"total warp-inst {tot:.3e}, samples {tots}")
src = open(srcpath).read().split("\n")
marks = [(i + 1, t.strip()[:80]) for i, t in enumerate(src)
         if t.strip().startswith("// ----") or t.strip().startswith("// (") or t.strip().startswith("//@") or re.match(r"^(__device__|__global__|template|static|int launch)", t)]
marks.append((10 ** 9, "end"))
for (a, name), (b, _) in zip(marks[:-1], marks[1:]):
    sel = [l for l in lines if a <= l[0] < b]
    i = sum(l[1] for l in sel); s_ = sum(l[2] for l in sel); t = sum(l[3] for l in sel)
    if i * 200 > tot or s_ * 200 > tots:
        print(f"  L{a:<4} inst {i*100/tot:5.1f}%  samples {s_*100/tots:5.1f}%  lanes {t/max(i,1):4.1f} | {name}")
print("top lines by instructions:")
for l in sorted(lines, key=lambda l: -l[1])[:topn]:
    print(f"  {l[0]:>4} inst {l[1]*100/tot:5.1f}% samp {l[2]*100/tots:5.1f}% lanes {l[3]/max(1,l[1]):4.1f} | {l[4].strip()[:95]}")
