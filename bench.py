#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path on B200: SD v1.5 512x512, 20-step DPM-Solver++(2M) txt2img with
classifier-free guidance (UNet batch 2 per image: cond + uncond) + VAE decode, random-init weights, synthetic prompts.

  python bench.py --gpus N --steps K --warmup W [--images-per-step I] [--impl reference]

One "step" = one call of the generate loop for I images per GPU (20 UNet steps + decode each).  Sample-parallel over
GPUs (one process per GPU, no data-path collective; weak scaling).  Prints ONE JSON line (rank 0).
  value    : images/s, device-event timed, inputs/outputs resident in HBM (libsdod_b200_generate_device)
  e2e      : images/s through the reference-facing C API with HOST buffers (libsdod_b200_generate: pinned H2D of
             conditioning + latents, D2H of the uint8 images inside the timed region)
  roofline : the dominant kernel (gemm_tcgen05_kernel: every linear layer and 3x3 conv) over one UNet pass at the
             benchmarked batch (2 x images per step), per-launch CUDA events, against the measured sustained bf16 peak
  step_roofline : BASELINE.json's second metric, "UNet step p50 ms" = config C2 (batch 2: one image's cond + uncond),
             one CUDA-graph replay of the whole step, 1.6065 TFLOP, against the same peak
  parity   : the GPU UNet on the CPU leg's own inputs and weights (seeds 0/1/2) at batch 2 AND inside the benchmarked
             batch — eps relative L2 vs the fp32 oracle, asserted <= 1e-2 (BASELINE.json north star)
  cpu_baseline : the oracle (fp32 PyTorch restatement + C sampler) on the host cores, bounded sample
--impl reference times that CPU path as its own arm: there one "step" = one CFG UNet step (cond + uncond) on the
host, ms_per_step is its mean time, and images/s = 1 / (20 steps + 1 VAE decode).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))

METRIC = "SD1.5 512x512 20-step CFG txt2img throughput"
UNIT = "images/s"
UNET_STEP_TFLOP = 1.6065            # batch-2 step, SURVEY §8(d) / Appendix B (803.3 GMAC)
WORKLOAD = "SD v1.5 txt2img 512x512: 20-step DPM-Solver++(2M), CFG 7.5 (UNet batch 2 per image), VAE decode; random-init weights"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_oracle_times(n_unet_steps=1, keep=None):
    """Bounded CPU sample of the same workload: one CFG UNet step (cond + uncond, as context.cpp:352,366) and one VAE
    decode at full size, fp32, all host threads -> extrapolated images/s = 1 / (20 * t_step + t_vae).
    keep (dict, optional) receives the first step's inputs, outputs and the oracle module for the GPU parity check."""
    import torch
    from oracle import ldm_oracle as L
    from oracle import sampler as S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    unet, vae = L.make_unet(0), L.make_vae(0)
    x = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(1))
    g = torch.Generator().manual_seed(2)
    cond, uncond = torch.randn(1, 77, 768, generator=g), torch.randn(1, 77, 768, generator=g)
    solver = S.OracleSolver()
    solver.prepare(20)
    ts = []
    with torch.no_grad():
        emb = unet.embed_time(torch.tensor(solver.table("model_ts")[:1]))
        xs = x.numpy().copy().ravel()
        for i in range(n_unet_steps):
            t0 = time.perf_counter()
            e_c_t, e_u_t = unet(x, emb, cond), unet(x, emb, uncond)
            e_c, e_u = e_c_t.numpy().ravel(), e_u_t.numpy().ravel()
            if keep is not None and i == 0:
                keep.update(unet=unet, x=x, emb=emb, cond=cond, uncond=uncond, eps_c=e_c_t.clone(), eps_u=e_u_t.clone())
            e = S.cfg_combine(e_c, e_u, 7.5)
            solver.update(0 if i == 0 else 1, xs, e)
            ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        vae(x * 0.18215)
        t_vae = time.perf_counter() - t0
    return ts, t_vae, cores


def run_reference(args, rank):
    if rank != 0:
        return
    ts, t_vae, cores = cpu_oracle_times(n_unet_steps=args.warmup + args.steps)
    ts = ts[args.warmup:]
    t_step = statistics.mean(ts)
    v = 1.0 / (20 * t_step + t_vae)
    sample = "%d timed CFG UNet steps (cond+uncond, 64x64 latent) + 1 VAE decode, fp32 PyTorch oracle; images/s = 1/(20*t_step + t_vae)" % len(ts)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t_step, "ms_per_image_extrapolated": 1000.0 * (20 * t_step + t_vae),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sampled": sample,
                                        "step": "one CFG UNet step (cond + uncond) on the host; steps x ms_per_step = the timed region"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "unet_cfg_step_s": t_step, "vae_decode_s": t_vae},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from sdod import libsdod as A
    from sdod import model as M
    from sdod import ops
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.images_per_step
    pk = peaks()

    ctx = A.Context("random-init:0", latent_spatial=64, steps=20, log_level=A.LOG_ERROR, max_images=n, device=local_rank)
    g = torch.Generator().manual_seed(2)
    cond_h, uncond_h = torch.randn(n, 77, 768, generator=g), torch.randn(n, 77, 768, generator=g)
    lat_h = torch.randn(n, 4, 64, 64, generator=torch.Generator().manual_seed(1))
    cond_d, uncond_d = cond_h.to(dev), uncond_h.to(dev)
    lat_d = lat_h.permute(0, 2, 3, 1).contiguous().to(dev)
    img_d = torch.empty(n, 512, 512, 3, dtype=torch.uint8, device=dev)
    cond_np, uncond_np, lat_np = cond_h.numpy(), uncond_h.numpy(), lat_h.numpy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > L2 (126 MB)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- warm-up (builds plans, captures graphs)
    for _ in range(max(args.warmup, 3)):
        ctx.generate_device(cond_d, uncond_d, lat_d, 7.5, img_d)
    ctx.generate(cond_np, uncond_np, lat_np, 7.5)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed: HBM-resident leg (device events inside the library bracket each call)
    launches0 = ops.launch_count()
    dev_ms, iter_ms = [], []
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        ctx.generate_device(cond_d, uncond_d, lat_d, 7.5, img_d)
        t = ctx.last_timings()
        dev_ms.append(t["total_ms"])
        iter_ms.append(t["iteration_ms"])
    barrier()
    wall_dev = time.perf_counter() - t0
    launches = ops.launch_count() - launches0
    # ---- timed: end-to-end leg through the host-pointer C API
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.generate(cond_np, uncond_np, lat_np, 7.5)
    barrier()
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---- UNet step p50 (batch 2, CUDA-graph replay) on its own stream, L2 flushed between replays
    unet = M.UNet(None, seed=0, latent_hw=64, max_batch=2)
    s = torch.cuda.Stream(device=dev)
    x2 = torch.randn(2, 64, 64, 4, device=dev)
    emb2 = torch.randn(2, 1280, device=dev)
    unet.set_context(torch.randn(2, 77, 768, device=dev))
    step_ms = []
    with torch.cuda.stream(s):
        for _ in range(4):
            unet.forward_nhwc(x2, emb2, use_graph=True)
        for _ in range(max(20, args.steps)):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s)
            unet.forward_nhwc(x2, emb2, use_graph=True)
            b.record(s)
            s.synchronize()
            step_ms.append(a.elapsed_time(b))
    unet_p50 = statistics.median(step_ms)
    unet_launches = unet.launches_per_forward(2)
    del unet

    # ---- dominant kernel (gemm_tcgen05_kernel: every linear layer and 3x3 conv of the step): device time of each of its
    #      launches in one UNet pass at this run's batch (2 x images per step), CUDA events between the ops of the plan on
    #      the launching stream; algorithmic FLOPs from the layer shapes.
    Bk = 2 * n
    unet_b = M.UNet(None, seed=0, latent_hw=64, max_batch=Bk)
    unet_b.set_context(torch.randn(Bk, 77, 768, device=dev))
    xb, embb = torch.randn(Bk, 64, 64, 4, device=dev), torch.randn(Bk, 1280, device=dev)
    with torch.cuda.stream(s):
        for _ in range(3):
            unet_b.forward_nhwc(xb, embb)
        rows = unet_b.profile(Bk, 5)
    gemm_ms, gemm_flop, gemm_flop_exec, gemm_n, all_ms = 0.0, 0.0, 0.0, 0, 0.0
    for ms, name in rows:
        all_ms += ms
        f = op_flops(name, Bk)
        if f:
            gemm_ms += ms
            gemm_flop += f
            gemm_flop_exec += op_flops_executed(name, Bk)
            gemm_n += 1
    del unet_b

    ctx.release()
    ctx = None
    multi = multi_gpu_records(args, rank, world, local_rank, dev, barrier) if (world >= 2 and not args.no_c5) else {}

    tot_dev_s = sum(dev_ms) / 1000.0
    t_dev = torch.tensor([tot_dev_s, wall_e2e, wall_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    tot_dev_s, wall_e2e, wall_dev = t_dev.tolist()
    images = args.steps * n * world
    if rank == 0:
        achieved = UNET_STEP_TFLOP / (unet_p50 / 1000.0)
        line = {
            "metric": METRIC, "value": images / tot_dev_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1000.0 * tot_dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_step_per_gpu": n, "latent": "4x64x64", "parallelism": "sample-parallel x%d" % world,
                       "l2": "flushed between timed iterations (256 MiB write)", "timing": "CUDA events inside the library around each generate call; max over ranks"},
            "unet_step_p50_ms": unet_p50, "unet_step_batch": 2, "iteration_ms_in_loop": statistics.median(iter_ms),
            "wall_s_device_leg": wall_dev,
            "e2e": {"value": images / wall_e2e, "unit": UNIT, "h2d_bytes_per_step": int(n * (2 * 77 * 768 * 4 + 4 * 64 * 64 * 4)),
                    "d2h_bytes_per_step": int(n * 512 * 512 * 3), "timing": "wall clock around libsdod_b200_generate with host buffers; max over ranks"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel: the %d linear / conv3x3 launches of one batch-%d UNet pass" % (gemm_n, Bk),
                         "achieved": gemm_flop / (gemm_ms * 1e-3) * 1e-12, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": gemm_flop / (gemm_ms * 1e-3) * 1e-12 / pk["bf16_tflops_sustained"],
                         "frac_of_executed_flops": gemm_flop_exec / (gemm_ms * 1e-3) * 1e-12 / pk["bf16_tflops_sustained"],
                         "flops_note": "algorithmic = the layer's definition (SURVEY App. B); the 3 Upsample convs run in sub-pixel form and execute 4/9 of theirs",
                         "peak_source": pk["source"] + " (sustained: the launches are timed inside a long step)",
                         "algorithmic_gflop_per_launch": gemm_flop / gemm_n * 1e-9, "avg_launch_us": 1e3 * gemm_ms / gemm_n,
                         "share_of_unet_pass": gemm_ms / all_ms, "traffic": traffic_from_profiles(),
                         "traffic_note": "mean dram__bytes_read+write per gemm_tcgen05 launch, ncu capture of the same pass (profiles/r02_gemm_traffic.json)"},
            "step_roofline": {"bound": "tensor", "kernel": "whole UNet denoising step, batch 2 (one CUDA-graph replay of %d kernels)" % unet_launches,
                              "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"],
                              "frac_of_burst_peak": achieved / pk["bf16_tflops"], "algorithmic_tflop_per_launch": UNET_STEP_TFLOP},
            "clocks": clocks,
        }
        line.update(multi)
        if not args.no_cpu_baseline:
            keep = {}
            ts, t_vae, cores = cpu_oracle_times(2, keep)             # the first step is the warm-up (and the parity vector), the second is timed
            v = 1.0 / (20 * ts[1] + t_vae)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "1 warm-up + 1 timed CFG UNet step (cond+uncond) + 1 VAE decode, fp32 PyTorch oracle on host; images/s = 1/(20*t_step + t_vae)",
                                    "unet_cfg_step_s": ts[1], "vae_decode_s": t_vae}
            line["parity"] = gpu_parity(keep, dev, Bk)
        print(json.dumps(line))
        if line.get("parity") and not line["parity"]["ok"]:
            raise SystemExit("bench.py: GPU eps differs from the oracle beyond tolerance: %s" % line["parity"])
    if world > 1:
        dist.destroy_process_group()


def multi_gpu_records(args, rank, world, local_rank, dev, barrier):
    """N >= 2 only: (1) `cfg_split` — the second partitioning of SURVEY §8e through the library (libsdod_b200_generate_pair: rank 2k runs the
    conditional half, rank 2k+1 the unconditional half, per-step eps exchange = peer stores inside the fused sampler kernel), at the same
    images per GPU as the headline, plus the single-image latency of a pair; (2) `c5` — BASELINE config C5: total batches of 8 / 32 / 128
    images over the N GPUs in both partitionings.  Wall clock around the calls (host buffers), barrier on both sides, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from sdod import libsdod as A
    from sdod import parallel as P
    n = args.images_per_step
    pairs = world // 2
    totals = [t for t in (8, 32, 128) if t % world == 0 and t // world <= 64 and t // pairs <= 64]
    cap = max([2 * n] + [t // pairs for t in totals])
    pair, role = P.pair_layout(rank, world)
    groups = P.make_pair_groups(world)
    ctx = A.Context("random-init:0", latent_spatial=64, steps=20, log_level=A.LOG_ERROR, max_images=cap, device=local_rank)
    ctx.set_seed(1234 + pair)                                # both ranks of a pair draw the same x_T
    P.connect_pair(ctx, groups[pair], role)
    g = torch.Generator().manual_seed(2 + pair)
    cond = torch.randn(cap, 77, 768, generator=g).numpy()
    uncond = torch.randn(cap, 77, 768, generator=g).numpy()
    lat = torch.randn(cap, 4, 64, 64, generator=torch.Generator().manual_seed(1 + pair)).numpy()
    half = cond if role == 0 else uncond

    def timed(fn, reps):
        fn()                                                 # builds the plan / graph of this batch size
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / reps

    out = {}
    m = 2 * n                                                # images per pair per call = the headline's images per GPU
    t_pair = timed(lambda: ctx.generate_pair(half[:m], lat[:m], 7.5), max(2, args.steps // 2))
    it_ms = ctx.last_timings()["iteration_ms"]
    t_one = timed(lambda: ctx.generate_pair(half[:1], lat[:1], 7.5), 3)
    out["cfg_split"] = {"images_per_s": pairs * m / t_pair, "unit": UNIT, "pairs": pairs, "images_per_pair_per_call": m, "unet_batch_per_gpu": m,
                        "iteration_ms_in_loop": it_ms, "single_image_latency_ms_on_a_pair": 1000.0 * t_one,
                        "eps_bytes_per_step_per_rank": m * 64 * 64 * 4 * 4,
                        "comm": "peer stores into the other GPU's IPC-mapped exchange buffer + flag, fused with the CFG + DPM update (cfg_dpm_step_pair_kernel); no NCCL on the data path",
                        "timing": "wall clock around libsdod_b200_generate_pair with host buffers; barrier both sides; max over ranks"}
    rows = []
    for total in totals:
        k = total // world                                   # sample-parallel: k images per GPU per call
        t_sp = timed(lambda: ctx.generate(cond[:k], uncond[:k], lat[:k], 7.5), 2)
        rows.append({"total_batch": total, "partition": "sample-parallel", "images_per_gpu_per_call": k, "images_per_s": total / t_sp})
        kp = total // pairs                                  # CFG split: kp images per pair per call
        t_cs = timed(lambda: ctx.generate_pair(half[:kp], lat[:kp], 7.5), 2)
        rows.append({"total_batch": total, "partition": "cfg-split", "images_per_pair_per_call": kp, "images_per_s": total / t_cs})
    out["c5"] = {"n_gpus": world, "rows": rows, "skipped_totals": [t for t in (8, 32, 128) if t not in totals],
                 "note": "BASELINE.json config C5; one generate call per measurement, 2 timed calls after 1 warm-up, host buffers"}
    ctx.release()
    return out


def gpu_parity(keep, dev, big_batch):
    """The product UNet on the CPU leg's inputs and weights: eps relative L2 vs the fp32 oracle at batch 2 (config C2) and inside
    the benchmarked batch (every slot of the batch carries one of the two oracle-checked samples, so every tile of the
    large-batch plan — CTA-pair and persistent GEMM variants included — has a known answer)."""
    import torch
    from sdod import model as M
    tol = 1e-2
    want = torch.cat([keep["eps_c"], keep["eps_u"]], 0).to(dev)
    weights = M.Weights(keep["unet"].state_dict())
    x2 = torch.cat([keep["x"], keep["x"]], 0)
    emb2 = torch.cat([keep["emb"], keep["emb"]], 0)
    ctx2 = torch.cat([keep["cond"], keep["uncond"]], 0)

    def rel(got, ref):
        return ((got.double() - ref.double()).norm() / ref.double().norm()).item()

    net = M.UNet(weights, latent_hw=64, max_batch=2)
    e2 = rel(net(x2, emb2, ctx2), want)
    del net
    reps = big_batch // 2
    net = M.UNet(weights, latent_hw=64, max_batch=big_batch)
    got = net(x2.repeat(reps, 1, 1, 1), emb2.repeat(reps, 1), ctx2.repeat(reps, 1, 1))
    per_slot = [rel(got[i:i + 1], want[i % 2:i % 2 + 1]) for i in range(big_batch)]
    del net
    torch.cuda.synchronize()
    worst = max(per_slot)
    return {"eps_rel_l2_batch2": e2, "eps_rel_l2_bench_batch_worst_slot": worst, "bench_batch": big_batch, "tol": tol,
            "ok": bool(e2 <= tol and worst <= tol), "oracle": "oracle/ldm_oracle.py fp32 on the host (seeds 0/1/2: weights / latent / prompts)"}


def op_flops(name, batch):
    """Algorithmic FLOPs of one op of the UNet plan from its profile name (None for non-GEMM ops)."""
    import re
    m = re.match(r"gemm M(\d+) N(\d+) K(\d+)", name)
    if m:
        return 2.0 * int(m.group(1)) * int(m.group(2)) * int(m.group(3))
    m = re.match(r"conv3 HW(\d+) Cin(\d+) Cout(\d+)", name)
    if m:
        return 2.0 * batch * int(m.group(1)) * int(m.group(3)) * 9 * int(m.group(2))
    m = re.match(r"conv3\+skip HW(\d+) Cin(\d+)\+(\d+) Cout(\d+)", name)     # 3x3 conv + the ResBlock's 1x1 skip projection in one launch
    if m:
        return 2.0 * batch * int(m.group(1)) * int(m.group(4)) * (9 * int(m.group(2)) + int(m.group(3)))
    m = re.match(r"conv3up2 HW(\d+) Cin(\d+) Cout(\d+)", name)               # conv3x3 on the 2x-upsampled image (HW = source pixels): 9 taps per output pixel
    if m:                                                                      # are the ALGORITHMIC FLOPs; the sub-pixel form executes 4/9 of them
        return 2.0 * batch * 4 * int(m.group(1)) * int(m.group(3)) * 9 * int(m.group(2))
    m = re.match(r"conv3s2 HW(\d+) Cin(\d+) Cout(\d+)", name)                # stride-2 conv: HW = output pixels
    if m:
        return 2.0 * batch * int(m.group(1)) * int(m.group(3)) * 9 * int(m.group(2))
    return None


def op_flops_executed(name, batch):
    """Multiply-adds the kernel really issues: equal to op_flops except for the sub-pixel Upsample convs (4 of 9 taps)."""
    import re
    f = op_flops(name, batch)
    return f * 4.0 / 9.0 if (f and re.match(r"conv3up2 ", name)) else f


def traffic_from_profiles():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (None if absent)."""
    path = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    try:
        return json.load(open(path))["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def main():
    knobs = sorted(k for k in os.environ if k.startswith("SDOD_"))
    if knobs:
        print("bench.py: note — A/B knobs set in the environment: %s" % knobs, file=sys.stderr)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--images-per-step", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="N >= 2: skip the cfg_split record and the C5 sweep")
    args = ap.parse_args()
    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
