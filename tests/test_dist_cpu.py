"""CPU, world_size 2, gloo: the host logic of the multi-GPU partitioning (sample sharding + the CFG-split exchange loop).
eps / step functions are injected: the oracle's CFG combine + DPM update stand in for the CUDA sampler, a fixed
nonlinear map stands in for the UNet halves — what is tested is the exchange protocol and that both ranks stay replicated
and agree with the single-process loop."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sampler as S
from sdod import parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _eps_half(x, step, role):
    w = 0.3 + 0.05 * step
    return torch.tanh(x * w) + (0.1 if role == 0 else -0.2) * torch.cos(x + step)


def _single_process(x0, steps, g):
    solver = S.OracleSolver()
    solver.prepare(steps)
    x = x0.numpy().copy().ravel()
    for s in range(steps):
        xt = torch.from_numpy(x.copy())
        e = S.cfg_combine(_eps_half(xt, s, 0).numpy(), _eps_half(xt, s, 1).numpy(), g)
        solver.update(s, x, e)
    return x


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        groups = P.make_pair_groups(world)
        pair, role = P.pair_layout(rank, world)
        steps, g = 20, 7.5
        x0 = torch.randn(4 * 8 * 8, generator=torch.Generator().manual_seed(100 + pair))
        solver = S.OracleSolver()
        solver.prepare(steps)

        def step_fn(s, x, e_c, e_u):
            xs = x.numpy()                       # shares memory: in-place update
            e = S.cfg_combine(e_c.numpy(), e_u.numpy(), g)
            solver.update(s, xs, e)

        loop = P.CfgSplitLoop(lambda x, s: _eps_half(x, s, role), step_fn, groups[pair], role)
        x = loop.run(x0.clone(), steps)
        q.put((rank, pair, role, x.numpy().copy(), loop.bytes_exchanged, x0.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_cfg_split_pair_exchange_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=100) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, pair0, role0, x_a, bytes_a, x0), (r1, pair1, role1, x_b, bytes_b, _) = out
    assert (pair0, role0, pair1, role1) == (0, 0, 0, 1)
    assert np.array_equal(x_a.view(np.uint32), x_b.view(np.uint32))                     # replicated state, bit for bit
    want = _single_process(torch.from_numpy(x0), 20, 7.5)
    assert np.array_equal(x_a.view(np.uint32), want.view(np.uint32))                    # == the unsplit loop
    assert bytes_a == bytes_b == 20 * 4 * 8 * 8 * 4                                     # one eps exchange per step


def test_sample_parallel_sharding():
    for total in (1, 7, 8, 32, 128, 129):
        for world in (1, 2, 4, 8):
            spans = [P.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.pair_layout(0, 3)
    assert [P.pair_layout(r, 8) for r in range(8)] == [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1)]
