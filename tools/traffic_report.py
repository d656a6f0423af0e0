"""DRAM traffic per kernel family from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` launch list: writes the JSON bench.py reads for `roofline.traffic` and prints a markdown table.
    python tools/traffic_report.py launches.csv profiles/r01_traffic.json"""
import csv, sys, json, collections, re

path, out = sys.argv[1], sys.argv[2]
import gzip

opener = gzip.open if path.endswith(".gz") else open
lines = [l for l in opener(path, "rt") if not l.startswith("==")]
acc = collections.OrderedDict()
cur = {}
rows_all = list(csv.DictReader(lines))
# one training step = from the fourth-from-last pack_fe_weights_kernel (one per extractor, first kernel of a forward) on
starts = [r["ID"] for r in rows_all if "pack_fe_weights_kernel" in r["Kernel Name"] and r["Metric Name"] == "gpu__time_duration.sum"]
first_id = int(starts[-4]) if len(starts) >= 4 else -1
for r in rows_all:
    if int(r["ID"]) < first_id:
        continue
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", r["Kernel Name"])
    fam = re.sub(r"[<(].*", "", name).replace("void ", "").strip()
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        v = v / 1000.0 if unit.startswith("n") else (v if unit.startswith("u") else v * 1000.0)  # us
    else:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    a = acc.setdefault(fam, {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    if m == "gpu__time_duration.sum":
        a["launches"] += 1; a["us"] += v
    elif m == "dram__bytes_read.sum":
        a["rd"] += v
    elif m == "dram__bytes_write.sum":
        a["wr"] += v
tot = sum(a["us"] for a in acc.values())
print(f"Total of the {sum(a['launches'] for a in acc.values())} captured launches: {tot/1000:.2f} ms (ncu: serialised, cold caches)\n")
print("| kernel family | launches | total us | share | avg us | DRAM read GB | DRAM write GB | GB/s |\n|---|---|---|---|---|---|---|---|")
for fam, a in sorted(acc.items(), key=lambda kv: -kv[1]["us"]):
    gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0
    print(f"| `{fam}` | {a['launches']} | {a['us']:.0f} | {100*a['us']/tot:.1f}% | {a['us']/a['launches']:.1f} | {a['rd']/1e9:.2f} | {a['wr']/1e9:.2f} | {gbs:.0f} |")
gemm = [a for f, a in acc.items() if f.startswith("koa::gemm_conv_kernel") or f.startswith("koa::gemm_kmajor_kernel")]
n = sum(a["launches"] for a in gemm)
js = {"source": path, "kernel_family": "gemm_conv_kernel + gemm_kmajor_kernel (fwd + dgrad)", "launches": n,
      "dram_bytes_per_launch": (sum(a["rd"] + a["wr"] for a in gemm) / n) if n else None,
      "us_per_launch_under_ncu": (sum(a["us"] for a in gemm) / n) if n else None,
      "share_of_captured_time": sum(a["us"] for a in gemm) / tot if tot else None}
json.dump(js, open(out, "w"), indent=1)
print("\n", json.dumps(js))
