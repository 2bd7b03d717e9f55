"""Per-op hot timings of the batch-B UNet forward plan (CUDA events between ops, eager). Development aid; output -> profiles/."""
import collections
import os
import re
import sys

import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import model as M  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
net = M.UNet(None, seed=0, latent_hw=64, max_batch=B)
net.set_context(torch.randn(B, 77, 768, device="cuda"))
x, emb = torch.randn(B, 64, 64, 4, device="cuda"), torch.randn(B, 1280, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        net.forward_nhwc(x, emb)
    rows = net.profile(B, 5)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "plan_ops_b%d.txt" % B), "w").write("\n".join("%.4f\t%s" % r for r in rows))
tot = sum(r[0] for r in rows)
print("ops %d  total %.3f ms (eager, per-op events)" % (len(rows), tot))
agg = collections.OrderedDict()
for ms, name in rows:
    key = re.sub(r"^(\w+).*", r"\1", name)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += ms
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.3f ms %5.1f%%  x%-4d avg %7.1f us  %s" % (ms, 100 * ms / tot, c, 1000 * ms / c, k))
print("---- by op (sorted)")
byname = collections.OrderedDict()
for ms, name in rows:
    a = byname.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
for k, (c, ms) in sorted(byname.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%8.3f ms  x%-3d avg %7.1f us  %s" % (ms, c, 1000 * ms / c, k))
