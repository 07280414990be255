"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time per kernel family."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(lambda: [0, 0.0])
for row in r:
    if len(row) <= vi:
        continue
    name = re.sub(r"\(.*", "", row[ki])
    name = re.sub(r"^void ", "", name)
    v = float(row[vi].replace(",", ""))
    u = row[ui]
    us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    tot[name][0] += 1
    tot[name][1] += us
all_us = sum(v[1] for v in tot.values())
print(f"total {all_us/1000:.2f} ms over {sum(v[0] for v in tot.values())} launches")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{us/1000:9.3f} ms {100*us/all_us:5.1f}%  n={n:5d}  avg={us/n:9.1f} us  {k[:110]}")
