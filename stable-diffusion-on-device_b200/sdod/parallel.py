"""Multi-GPU partitioning of the hot path on one 8xB200 box (one process per GPU, torch.distributed / NCCL).

The path shards in exactly two ways, both present in the reference's own loop structure:
  * sample-parallel: images are independent (the reference loop carries only per-image x / y_prev,
    context.cpp:333-378) -> split the batch over ranks; no data-path collective at all.
  * CFG split over GPU pairs: the reference already evaluates cond and uncond as two separate UNet calls
    (context.cpp:352 and :366) -> rank 2k runs the cond half, rank 2k+1 the uncond half on the same latents;
    ONE exchange per denoising step (a 2-rank all-gather of eps, 64 KB fp32 per image) after which both
    ranks apply the identical fused CFG+DPM update, so x / y_prev stay replicated and no second exchange
    is needed.  The collective exists only because the path has this real exchange step.
The loop below is backend-agnostic (NCCL on GPUs; gloo in the CPU tests with injected eps/step functions).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced [start, end) of `total` independent images for `rank` (sample-parallel)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pair_layout(rank, world):
    """(pair index, role) with role 0 = cond half, 1 = uncond half."""
    if world % 2 != 0:
        raise ValueError("CFG split needs an even number of ranks, got %d" % world)
    return rank // 2, rank % 2


def make_pair_groups(world):
    """Every rank must create every pair group (torch.distributed rule); returns the list indexed by pair."""
    if world % 2 != 0:
        raise ValueError("CFG split needs an even number of ranks, got %d" % world)
    return [dist.new_group([2 * k, 2 * k + 1]) for k in range(world // 2)]


class CfgSplitLoop:
    """Denoising loop of one GPU pair.

    eps_fn(x, step) -> this rank's half of the noise prediction (cond for role 0, uncond for role 1)
    step_fn(step, x, eps_cond, eps_uncond) -> updates x (and its own y_prev state) in place or returns the new x
    """

    def __init__(self, eps_fn, step_fn, group, role):
        self.eps_fn, self.step_fn, self.group, self.role = eps_fn, step_fn, group, role
        self.bytes_exchanged = 0

    def run(self, x, steps):
        for s in range(steps):
            e_local = self.eps_fn(x, s).contiguous()
            both = [torch.empty_like(e_local), torch.empty_like(e_local)]
            dist.all_gather(both, e_local, group=self.group)           # slot r = rank r of the pair: [cond, uncond]
            self.bytes_exchanged += e_local.numel() * e_local.element_size()
            out = self.step_fn(s, x, both[0], both[1])
            if out is not None:
                x = out
        return x


def connect_pair(ctx, group, role):
    """Wire two libsdod contexts (one per process / GPU) into a CFG-split pair: exchange the 64-byte CUDA IPC handles of their eps
    exchange buffers over `group` (any backend: this is set-up plumbing, the per-step exchange itself is peer stores inside the fused
    sampler kernel — csrc/kernels/sampler.cu: cfg_dpm_step_pair_kernel) and map the peer's buffer."""
    mine = ctx.pair_export()
    both = [None, None]
    dist.all_gather_object(both, mine, group=group)
    ctx.pair_connect(both[1 - role], role)


def gpu_cfg_split_generate(unet, vae, emb_all, latents_nhwc, guidance, group, role, steps=20):
    """CFG-split generation on GPUs: `unet` already holds this rank's context (cond or uncond), batch = images."""
    from . import ops
    x = latents_nhwc.clone()
    y_prev = torch.zeros_like(x)
    n = x.shape[0]

    def eps_fn(xc, s):
        return unet.forward_nhwc(xc, emb_all[s:s + 1].expand(n, -1).contiguous(), use_graph=True)

    def step_fn(s, xc, e_c, e_u):
        k = ops.dpm_coeffs(s, steps)
        torch.ops.sdod.cfg_dpm_step(xc.view(-1), y_prev.view(-1), e_c.view(-1), e_u.view(-1), guidance, k["sigma_s"], k["alpha_s"], k["c_x"],
                                    k["c_prev"], k["c_y0"], k["order"])

    loop = CfgSplitLoop(eps_fn, step_fn, group, role)
    x = loop.run(x, steps)
    # decode: split the images of the pair between its two ranks
    lo, hi = shard_range(n, role, 2)
    imgs = None
    if hi > lo:
        imgs, _ = vae(x[lo:hi].permute(0, 3, 1, 2).contiguous(), use_graph=False)
    return x, imgs, loop.bytes_exchanged
