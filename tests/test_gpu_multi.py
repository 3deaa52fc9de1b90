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


@pytest.mark.parametrize("world,n,alg,prec", [(2, 4096, "direct", "fp64"), (2, 4099, "direct", "fp64"),
                                              (2, 4096, "tree", "fp64"), (2, 8192, "direct", "fp32"),
                                              (2, 8192, "tree", "fp32")])
def test_sharded_matches_single_gpu(tmp_path, world, n, alg, prec):
    if _lib.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from gravhopper_b200.sharded import ShardedSimulation
    steps = 3
    out = str(tmp_path / "multi.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "_sharded_worker.py"), out, str(n), alg, prec, str(steps)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = np.load(out)
    x, v, m = ic_raw.Plummer(n, 1e-3, 1e6, seed=21)
    single = ShardedSimulation(x, v, m, 0.005, 5e-5, algorithm=alg, precision=prec)
    single.run(steps)
    pos, vel = single.gather_state()
    tol = 1e-13 if prec == "fp64" else 1e-6
    assert np.abs(got["pos"] - pos).max() <= tol * np.abs(pos).max()
    assert np.abs(got["vel"] - vel).max() <= tol * np.abs(vel).max()
