"""CFG split over a GPU pair through the libsdod C API (libsdod_b200_pair_export / _connect / _generate_pair), launched as
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tools/pair_check.py [--S 16] [--images 2] [--perf]
Each rank owns one GPU and one context; rank 0 = conditional half, rank 1 = unconditional half.  Checks (rank 0 prints one JSON line):
  * both ranks end with bit-identical sampler state (the update is replicated, not exchanged),
  * final latents / images vs the ORACLE generate loop (oracle/pipeline.py) on the same weights, latents and prompts,
  * --perf: single-image 512x512 latency on the pair vs on one GPU (unsplit libsdod_b200_generate), and images/s at n images per pair."""
import argparse
import json
import math
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import libsdod as A  # noqa: E402
from sdod import model as M  # noqa: E402
from sdod import parallel as P  # noqa: E402


def psnr_u8(a, b):
    mse = ((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean()
    return 10 * math.log10(255.0 ** 2 / max(mse, 1e-12))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--S", type=int, default=16)
    ap.add_argument("--images", dest="n", type=int, default=2)   # (not "--n": torchrun's argparse rejects it as an ambiguous abbreviation of its own options)
    ap.add_argument("--perf", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    assert world == 2, "one pair"
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo")
    role = rank
    res = {}
    if not args.perf:
        from oracle import ldm_oracle as L
        from oracle import pipeline as OP
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        unet, vae = L.make_unet(0), L.make_vae(0)
        d = tempfile.mkdtemp(prefix="sdod_pair_%d_" % rank)
        M.save_weight_file(os.path.join(d, "unet.sdodw"), unet.state_dict())
        M.save_weight_file(os.path.join(d, "vae_decoder.sdodw"), vae.state_dict())
        models = d
    else:
        models = "random-init:0"
    S, n = args.S, args.n
    lat = torch.randn(n, 4, S, S, generator=torch.Generator().manual_seed(1))
    g2 = torch.Generator().manual_seed(2)
    cond, uncond = torch.randn(n, 77, 768, generator=g2), torch.randn(n, 77, 768, generator=g2)
    half = (cond if role == 0 else uncond).numpy()
    with A.Context(models, latent_spatial=S, steps=20, max_images=n, device=rank) as ctx:
        P.connect_pair(ctx, dist.group.WORLD, role)
        imgs, first, cnt, lat_out = ctx.generate_pair(half, lat.numpy(), 7.5, return_latents=True)
        imgs2, _, _, lat2 = ctx.generate_pair(half, lat.numpy(), 7.5, return_latents=True)         # sequence numbers keep counting across calls
        assert np.array_equal(lat_out, lat2) and np.array_equal(imgs, imgs2)
        t_lat = torch.from_numpy(lat_out)
        both = [torch.empty_like(t_lat), torch.empty_like(t_lat)]
        dist.all_gather(both, t_lat)
        t_img = torch.from_numpy(imgs)
        dist.all_reduce(t_img, op=dist.ReduceOp.MAX)            # each rank filled its own share, the rest is zero
        res["replicated_bitwise"] = bool(torch.equal(both[0], both[1]))
        res["decode_split"] = [int(first), int(cnt)]
        if args.perf:
            times = []
            for _ in range(5):
                dist.barrier()
                t0 = time.perf_counter()
                ctx.generate_pair(half, lat.numpy(), 7.5)
                dist.barrier()
                times.append(time.perf_counter() - t0)
            res["pair_call_s"] = min(times)
            res["pair_iteration_ms"] = ctx.last_timings()["iteration_ms"]
            res["images_per_s_per_pair"] = n / min(times)
    if rank == 0:
        if not args.perf:
            want_u8, _, want_lat = OP.generate(unet, vae, cond, uncond, lat, 7.5, 20, device="cuda")
            res["psnr_vs_oracle_db"] = min(psnr_u8(t_img.numpy()[i], want_u8[i]) for i in range(n))
            res["final_latent_rel_l2_vs_oracle"] = float(np.linalg.norm(lat_out - want_lat) / np.linalg.norm(want_lat))
            with A.Context(models, latent_spatial=S, steps=20, max_images=n, device=0) as one:
                u8, lat_one = one.generate(cond.numpy(), uncond.numpy(), lat.numpy(), 7.5, return_latents=True)
            res["split_vs_unsplit_latent_rel_l2"] = float(np.linalg.norm(lat_out - lat_one) / np.linalg.norm(lat_one))
        else:
            with A.Context(models, latent_spatial=S, steps=20, max_images=n, device=0) as one:
                one.generate(cond.numpy(), uncond.numpy(), lat.numpy(), 7.5)
                ts = []
                for _ in range(5):
                    t0 = time.perf_counter()
                    one.generate(cond.numpy(), uncond.numpy(), lat.numpy(), 7.5)
                    ts.append(time.perf_counter() - t0)
                res["one_gpu_call_s"] = min(ts)
                res["one_gpu_iteration_ms"] = one.last_timings()["iteration_ms"]
        res.update(S=S, n=n, eps_bytes_per_step_per_rank=n * S * S * 4 * 4, comm="peer stores (CUDA IPC mapped buffer) inside cfg_dpm_step_pair_kernel")
        line = json.dumps(res)
        print(line)
        if args.out:
            open(args.out, "w").write(line + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
