"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, rows = r, rows[i + 1:]
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
    d[r[ki][:70]][0] += 1
    d[r[ki][:70]][1] += v
tot = sum(v[1] for v in d.values())
print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'avg us':>9s} {'share':>6s}")
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:9.2f} {100 * v[1] / tot:5.1f}%")
print(f"{'total':70s} {sum(v[0] for v in d.values()):8d} {tot:12.1f}")
