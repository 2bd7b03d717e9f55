"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel-name count / total / share."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
idx = {h: i for i, h in enumerate(hdr)}
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for row in r:
    if row[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(row[idx["Metric Value"]].replace(",", ""))
    unit = row[idx["Metric Unit"]]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*", "", row[idx["Kernel Name"]])
    rows.append((int(row[idx["ID"]]), name, row[idx["Grid Size"]], row[idx["Block Size"]], us))
rows = rows[skip:]
tot = sum(r[4] for r in rows)
agg = collections.OrderedDict()
for _, name, grid, blk, us in rows:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
print("launches %d  total %.1f us" % (len(rows), tot))
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%7.1f us  %5.1f%%  x%-4d avg %7.2f us  %s" % (us, 100 * us / tot, c, us / c, name))
if len(sys.argv) > 3:
    print("---- top individual launches")
    for r_ in sorted(rows, key=lambda r: -r[4])[:int(sys.argv[3])]:
        print("%8.1f us  id %-5d grid %-18s %s" % (r_[4], r_[0], r_[2], r_[1]))
