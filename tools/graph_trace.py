"""In-graph kernel timeline of the UNet step: CUPTI activity records (torch.profiler) of CUDA-graph replays -> per-kernel start/duration,
so kernel time and the gaps between dependent kernels can be told apart (ncu serialises and cold-starts every launch; eager per-op events
include host launch cost).   python tools/graph_trace.py [B=2] [hw=64] [tag]  -> gpurun_out/graph_trace_<tag>.txt"""
import collections
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import model as M  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 64
tag = sys.argv[3] if len(sys.argv) > 3 else "b%d_hw%d" % (B, HW)
dev = torch.device("cuda")
net = M.UNet(None, seed=0, latent_hw=HW, max_batch=B)
net.set_context(torch.randn(B, 77, 768, device=dev))
x, emb = torch.randn(B, HW, HW, 4, device=dev), torch.randn(B, 1280, device=dev)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(5):
        net.forward_nhwc(x, emb, use_graph=True)
    s.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            net.forward_nhwc(x, emb, use_graph=True)
        s.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
ev.sort(key=lambda e: e.time_range.start)
n = len(ev) // 3
last = ev[2 * n:]                       # the third replay
t0 = last[0].time_range.start
span = last[-1].time_range.end - t0
busy = sum(e.time_range.end - e.time_range.start for e in last)
gaps = []
for a, b in zip(last[:-1], last[1:]):
    gaps.append(b.time_range.start - a.time_range.end)
out = ["B=%d hw=%d kernels/replay %d  span %.1f us  sum of kernel durations %.1f us  sum of positive gaps %.1f us  overlapped (negative gaps) %.1f us" % (
    B, HW, len(last), span, busy, sum(g for g in gaps if g > 0), -sum(g for g in gaps if g < 0))]
agg = collections.OrderedDict()
for i, e in enumerate(last):
    k = re.sub(r"\(.*", "", e.name).replace("void ", "").replace("sdod::", "")[:48]
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
    if i + 1 < len(last):
        a[2] += max(0.0, last[i + 1].time_range.start - e.time_range.end)
out.append("%10s %6s %5s %8s %10s  %s" % ("dur us", "share", "n", "avg us", "gap-after", "kernel"))
for k, (c, d, g) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%10.1f %5.1f%% %5d %8.2f %10.1f  %s" % (d, 100 * d / busy, c, d / c, g, k))
# label every kernel with its plan op (ops are recorded in launch order; a split-K op without the in-cluster reduction and the two-kernel
# GroupNorm are two launches each) and attribute the span by end-to-end deltas: the time an op adds to the step once overlap is accounted for
with torch.cuda.stream(s):
    names = [n for _, n in net.profile(B, 1)]
labels = []
for n in names:
    m = re.search(r" split(\d+)", n)
    k = 2 if ((m and int(m.group(1)) > 1) or n.startswith("gn ")) else 1
    labels += [n] * k
if len(labels) == len(last):
    byop = collections.OrderedDict()
    prev_end = last[0].time_range.start
    for e, n in zip(last, labels):
        a = byop.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += e.time_range.end - prev_end
        prev_end = e.time_range.end
    out.append("---- in-graph cost by plan op (end-to-end deltas; sums to the span)")
    for n, (c, d) in sorted(byop.items(), key=lambda kv: -kv[1][1]):
        out.append("%9.1f us %5.1f%%  x%-3d avg %7.2f us  %s" % (d, 100 * d / span, c, d / c, n))
else:
    out.append("---- (op labels not aligned: %d labels vs %d kernels)" % (len(labels), len(last)))
    labels = [""] * len(last)
out.append("---- timeline (start us, dur us, gap to next us, kernel)")
for i, e in enumerate(last):
    g = last[i + 1].time_range.start - e.time_range.end if i + 1 < len(last) else 0.0
    out.append("%9.1f %7.2f %6.2f  %-52s %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, g, re.sub(r"\(.*", "", e.name).replace("void ", "").replace("sdod::", "")[:52], labels[i]))
cut = [i for i, l in enumerate(out) if l.startswith("---- timeline")][0]
print("\n".join(out[:min(cut, 90)]))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "graph_trace_%s.txt" % tag), "w").write("\n".join(out) + "\n")
