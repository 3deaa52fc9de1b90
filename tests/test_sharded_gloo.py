"""CPU test of the multi-GPU host logic: world_size 2 over gloo, with a stand-in shard that
computes with the oracle (tests may use the oracle; the product never does).  Checks that the
partition + in-place all-gather + double-buffer sequencing of ShardedSimulation reproduces the
single-rank trajectory bit for bit."""
import contextlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShard(object):
    """Same interface as gravhopper_b200.sharded.CudaShard, on CPU tensors, fp64."""

    def __init__(self, n, begin, count):
        from oracle import oracle as O
        self.O = O
        self.n, self.begin, self.count = n, begin, count
        self.bufs = [torch.zeros((n, 3), dtype=torch.float64) for _ in range(2)]
        self.cur = 0

    def upload(self, pos, vel, mass_all, origin):
        self.x, self.v, self.m = np.array(pos), np.array(vel), np.array(mass_all)

    def prepare(self, dt):
        xh = self.O.half_drift(self.x, self.v, dt)
        self.bufs[self.cur][self.begin:self.begin + self.count] = torch.from_numpy(xh)

    def source_index(self):
        return self.cur

    def stream_context(self):
        return contextlib.nullcontext()

    def step(self, dt, eps, theta, alg):
        src = self.bufs[self.cur].numpy()
        xh = src[self.begin:self.begin + self.count]
        if alg == 0:
            a = self.O.direct_summation_position(src, self.m, xh, eps)
        else:
            a = self.O.tree_force_position(src, self.m, xh, eps, theta)
        K = self.O.KPC_PER_KMS_MYR
        self.v = self.v + (a * self.O.C_ACC) * dt
        self.x = xh + ((0.5 * self.v) * dt) * K
        nxt = self.x + ((0.5 * self.v) * dt) * K
        self.cur ^= 1
        self.bufs[self.cur][self.begin:self.begin + self.count] = torch.from_numpy(nxt)

    def download(self):
        return self.x.copy(), self.v.copy()


def _worker(rank, world, port, n, alg, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gravhopper_b200 import ic_raw
    from gravhopper_b200.sharded import ShardedSimulation
    x, v, m = ic_raw.Plummer(n, 1e-3, 1e6, seed=9)
    sim = ShardedSimulation(x, v, m, 0.005, 5e-5, algorithm=alg, rank=rank, world=world,
                            shard_factory=lambda nn, b, c: OracleShard(nn, b, c))
    sim.run(3)
    pos, vel = sim.gather_state()
    if rank == 0:
        q.put((pos, vel))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,alg", [(64, "direct"), (65, "direct"), (96, "tree")])
def test_world2_matches_single_rank(n, alg):
    from oracle import oracle as O
    from gravhopper_b200 import ic_raw
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, alg, q)) for r in range(2)]
    for p in procs:
        p.start()
    pos, vel = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, v, m = ic_raw.Plummer(n, 1e-3, 1e6, seed=9)
    for _ in range(3):
        x, v, _ = O.leapfrog_step(x, v, m, 0.005, 5e-5, alg, theta=0.7)
    # direct: the sharded run sums the same sources in the same order -> bitwise
    if alg == "direct":
        assert np.allclose(pos, x, rtol=0, atol=1e-15 * np.abs(x).max())
    else:
        assert np.allclose(pos, x, rtol=0, atol=1e-14 * np.abs(x).max())
    assert np.allclose(vel, v, rtol=1e-12)
