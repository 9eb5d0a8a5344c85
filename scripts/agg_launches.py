"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict(); seq = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:70]
    seq.append((name, row["Grid Size"], v))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.1f} | {100*t/tot:.1f}% |")
print(f"| total | {len(seq)} | {tot:.1f} | |")
if len(sys.argv) > 2:
    for s in seq: print(s)
