"""UNet step latency (CUDA-graph replay, L2 flushed between replays — the number bench.py reports as unet_step_p50_ms) plus the hot per-op table.
Usage: python tools/step_time.py [B=2] [tag]   -> gpurun_out/step_time_<tag>.txt"""
import collections
import os
import re
import statistics
import sys

import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import model as M  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tag = sys.argv[2] if len(sys.argv) > 2 else "b%d" % B
HW = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda")
net = M.UNet(None, seed=0, latent_hw=HW, max_batch=B)
net.set_context(torch.randn(B, 77, 768, device=dev))
x, emb = torch.randn(B, HW, HW, 4, device=dev), torch.randn(B, 1280, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream()
cold, warm = [], []
with torch.cuda.stream(s):
    for _ in range(4):
        net.forward_nhwc(x, emb, use_graph=True)
    for it in range(40):
        if it < 20:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        net.forward_nhwc(x, emb, use_graph=True)
        b.record(s)
        s.synchronize()
        (cold if it < 20 else warm).append(a.elapsed_time(b))
    rows = net.profile(B, 5)
out = []
out.append("B=%d hw=%d launches/forward %d  ops %d  env %s" % (B, HW, net.launches_per_forward(B), len(rows),
                                                      {k: v for k, v in os.environ.items() if k.startswith("SDOD_")}))
out.append("graph replay: p50 %.3f ms (L2 flushed before each replay), %.3f ms back-to-back; min %.3f" % (statistics.median(cold), statistics.median(warm), min(cold + warm)))
tot = sum(r[0] for r in rows)
out.append("eager per-op events: total %.3f ms" % tot)
agg = collections.OrderedDict()
for ms, name in rows:
    a = agg.setdefault(re.sub(r"^(\S+).*", r"\1", name), [0, 0.0])
    a[0] += 1
    a[1] += ms
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%8.3f ms %5.1f%%  x%-4d avg %7.1f us  %s" % (ms, 100 * ms / tot, c, 1000 * ms / c, k))
out.append("---- by op")
byname = collections.OrderedDict()
for ms, name in rows:
    a = byname.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
for k, (c, ms) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    out.append("%8.3f ms  x%-3d avg %7.1f us  %s" % (ms, c, 1000 * ms / c, k))
txt = "\n".join(out)
print("\n".join(out[:16]))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "step_time_%s.txt" % tag), "w").write(txt + "\n")
