"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the sharded run over NCCL
must reproduce the single-GPU run.  fp64 direct sums the same sources in the same tile order on
every rank, so it is compared tightly; the tree accepts the same nodes."""
import os
import subprocess
import sys

import numpy as np
import pytest

from gravhopper_b200 import _lib, ic_raw

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# comm "native": NCCL communicator inside the library (fp32 tree: DISTRIBUTED build from the second
# step on); "torch": torch.distributed all-gather + redundant build (the round-1 path)
@pytest.mark.parametrize("world,n,alg,prec,steps,comm", [
    (2, 4096, "direct", "fp64", 3, "native"), (2, 4099, "direct", "fp64", 3, "native"),
    (2, 4096, "tree", "fp64", 3, "native"), (2, 8192, "direct", "fp32", 3, "native"),
    (2, 8192, "tree", "fp32", 6, "native"), (2, 8195, "tree", "fp32", 4, "native"),
    (2, 50000, "tree", "fp32", 5, "native"), (2, 8192, "tree", "fp32", 3, "torch"),
    (2, 4096, "direct", "fp64", 3, "torch")])
def test_sharded_matches_single_gpu(tmp_path, world, n, alg, prec, steps, comm):
    if _lib.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from gravhopper_b200.sharded import ShardedSimulation
    out = str(tmp_path / "multi.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "_sharded_worker.py"), out, str(n), alg, prec, str(steps)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, GH_COMM=comm))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = np.load(out)
    x, v, m = ic_raw.Plummer(n, 1e-3, 1e6, seed=21)
    single = ShardedSimulation(x, v, m, 0.005, 5e-5, algorithm=alg, precision=prec)
    single.run(steps)
    pos, vel = single.gather_state()
    # fp32 tree: the ranks' key ranges do not start on the single-GPU run's group boundaries (and the
    # torch path groups each rank's OWN targets), so the groups of 32 -- and with them the
    # interaction lists -- differ: the runs agree to the tree's own approximation error (~1e-2 of
    # the force, i.e. ~1e-5 of the positions after a few of these short steps), not to rounding
    if prec == "fp32" and alg == "tree":
        # Measured (tests/_sharded_debug.py, N = 8192, 6 steps, 2 GPUs): positions median 1.7e-6 / p99
        # 1.4e-5 / max 4.4e-5 of max|x|, velocities median 9.2e-5 / p99 7.2e-4 / max 1.9e-3 of max|v| --
        # already 4e-5 after ONE step: the force differs by the group walk's own error (~3e-3) and the
        # kick is ~1 % of max|v|.  The redundant-build path (GH_TREE_DIST=0) differs by the same amount.
        for (a, b), (med, p99, mx) in zip(((got["pos"], pos), (got["vel"], vel)),
                                          ((2e-5, 2e-4, 2e-3), (1e-3, 1e-2, 5e-2))):
            d = np.abs(a - b).max(axis=1) / np.abs(b).max()
            assert np.median(d) <= med and np.percentile(d, 99) <= p99 and d.max() <= mx
        return
    tol = 1e-13 if prec == "fp64" else 1e-6
    assert np.abs(got["pos"] - pos).max() <= tol * np.abs(pos).max()
    assert np.abs(got["vel"] - vel).max() <= tol * np.abs(vel).max()


def test_simulation_devices_argument_matches_one_gpu():
    """Simulation(devices=2): single-process multi-GPU inside the drop-in class (ncclCommInitAll
    behind the C ABI); same trajectories as one GPU."""
    if _lib.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from gravhopper_b200 import Simulation
    x, v, m = ic_raw.Plummer(6000, 1e-3, 1e6, seed=8)
    res = {}
    for alg, prec in (("direct", "fp64"), ("tree", "fp32")):
        for ndev in (1, 2):
            sim = Simulation(dt=0.005, eps=5e-5, algorithm=alg, precision=prec, devices=ndev)
            sim.add_IC({"pos": x, "vel": v, "mass": m})
            sim.run(5)
            res[ndev] = (np.asarray(sim.positions.value)[-1].copy(), np.asarray(sim.velocities.value)[-1].copy())
        for k in (0, 1):
            d = np.abs(res[2][k] - res[1][k]).max(axis=1) / np.abs(res[1][k]).max()
            if prec == "fp64":
                assert d.max() <= 1e-13, (alg, prec)
            else:  # fp32 tree: see test_sharded_matches_single_gpu
                med, p99, mx = ((2e-5, 2e-4, 2e-3), (1e-3, 1e-2, 5e-2))[k]
                assert np.median(d) <= med and np.percentile(d, 99) <= p99 and d.max() <= mx, (alg, prec)
