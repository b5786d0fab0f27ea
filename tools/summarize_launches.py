"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    m = re.search(r"GemmCfg<(\d+)", r["Kernel Name"])
    if m:
        name = f"k_gemm_tiles<{m.group(1)}>"
    m = re.search(r"(k_factor_small|k_update_small)<\(?int\)?(\d+)>", r["Kernel Name"])
    if m:
        name = f"{m.group(1)}<{m.group(2)}>"
    v = float(r["Metric Value"].replace(",", ""))
    v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
    a = agg[name]
    a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us serialised (cold cache; compare shares)")
print(f"{'kernel':32s} {'launches':>8s} {'total_us':>10s} {'max_us':>9s} {'share':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:32s} {a[0]:8d} {a[1]:10.1f} {a[2]:9.1f} {100 * a[1] / tot:6.1f}%")
