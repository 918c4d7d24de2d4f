"""Summarise an ncu --csv launch list (gpu__time_duration.sum) per kernel name."""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
ui = hdr.index("Metric Unit")
agg = collections.OrderedDict(); total = 0.0
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum": continue
    v = float(r[vi].replace(",", "")); u = r[ui]
    us = v / 1000 if u in ("nsecond", "ns") else (v if u in ("usecond", "us") else v * 1000)
    name = re.sub(r"\(.*", "", r[ki]); name = re.sub(r"<.*", "", name)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; total += us
print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'share':>7s}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {c:8d} {t:12.1f} {100*t/total:6.1f}%")
print(f"{'TOTAL':60s} {sum(c for c,_ in agg.values()):8d} {total:12.1f}")
