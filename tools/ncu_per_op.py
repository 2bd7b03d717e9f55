"""Join an ncu launch list (gpu__time_duration) of tools/unet_step.py with the plan's ordered op names -> per-op table."""
import collections
import csv
import re
import sys

ops = [l.rstrip("\n").split("\t") for l in open(sys.argv[1])]
lines = [l for l in open(sys.argv[2]) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
idx = {h: i for i, h in enumerate(hdr)}
rows = []
for row in r:
    if row[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(row[idx["Metric Value"]].replace(",", ""))
    rows.append((re.sub(r"\(.*", "", row[idx["Kernel Name"]]).replace("void ", "").replace("sdod::", ""), row[idx["Grid Size"]], v / 1000.0))
starts = [i for i, rw in enumerate(rows) if rw[0].startswith("silu_f32")]
n_launch = 0
for ms, name in ops:
    n = 2 if name.startswith("gn ") else 1
    m = re.search(r"split(\d+)", name)
    if m and int(m.group(1)) > 1:
        n = 2
    n_launch += n
s0 = [s for s in starts if s + n_launch <= len(rows)][-1]
rows = rows[s0:s0 + n_launch]
i = 0
per = collections.OrderedDict()
for ms, name in ops:
    n = 2 if name.startswith("gn ") else 1
    m = re.search(r"split(\d+)", name)
    if m and int(m.group(1)) > 1:
        n = 2
    us = sum(rows[i + k][2] for k in range(n))
    kn = "+".join(rows[i + k][0][:26] for k in range(n))
    i += n
    a = per.setdefault(name, [0, 0.0, kn])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in per.values())
print("launches %d, total %.1f us (ncu gpu__time_duration per launch, --cache-control none)" % (n_launch, tot))
for name, (c, us, kn) in sorted(per.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print("%8.1f us %4.1f%% x%-3d avg %6.1f  %-54s %s" % (us, 100 * us / tot, c, us / c, name, kn))
