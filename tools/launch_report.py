"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name (markdown table)."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
acc = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    a = acc.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in acc.values())
print(f"Total of the {sum(a[0] for a in acc.values())} captured launches: {tot/1000:.2f} ms\n")
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for name, (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:100]}` | {n} | {us:.1f} | {100*us/tot:.1f}% | {us/n:.1f} |")
