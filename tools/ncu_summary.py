"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per-kernel-name count / total time / share (/ DRAM bytes per launch).   python tools/ncu_summary.py list.csv [skip] [top] [last_n]"""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
idx = {h: i for i, h in enumerate(hdr)}
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
launches = collections.OrderedDict()      # id -> [name, grid, block, us, dram bytes]
for row in r:
    m = row[idx["Metric Name"]]
    v = float(row[idx["Metric Value"]].replace(",", ""))
    unit = row[idx["Metric Unit"]]
    rec = launches.setdefault(int(row[idx["ID"]]), [re.sub(r"\(.*", "", row[idx["Kernel Name"]]), row[idx["Grid Size"]], row[idx["Block Size"]], 0.0, 0.0])
    if m == "gpu__time_duration.sum":
        rec[3] = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        rec[4] += v * mult
rows = list(launches.items())[skip:]
if len(sys.argv) > 4:
    rows = rows[-int(sys.argv[4]):]
tot = sum(rec[3] for _, rec in rows)
agg = collections.OrderedDict()
for _, (name, grid, blk, us, by) in rows:
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += us
    a[2] += by
print("launches %d  total %.1f us  DRAM bytes %.1f MB" % (len(rows), tot, sum(rec[4] for _, rec in rows) / 1e6))
for name, (c, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%9.1f us  %5.1f%%  x%-4d avg %8.2f us  %9.2f MB/launch  %s" % (us, 100 * us / tot, c, us / c, by / c / 1e6, name))
if len(sys.argv) > 3 and int(sys.argv[3]) > 0:
    print("---- top individual launches")
    for i, rec in sorted(rows, key=lambda kv: -kv[1][3])[:int(sys.argv[3])]:
        print("%8.1f us  %8.2f MB  id %-5d grid %-18s %s" % (rec[3], rec[4] / 1e6, i, rec[1], rec[0]))
