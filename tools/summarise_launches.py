"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1; d[1] += float(r[-1]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {ms:.2f} | {ms/tot*100:.1f}% | {ms/n:.3f} |")
print(f"| total | {sum(v[0] for v in agg.values())} | {tot:.2f} | 100% | |")
