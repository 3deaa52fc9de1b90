"""Bitwise check of the running simulation's sort / emit forms at a given size (scratch tool for a
GPU box): runs `steps` fp32 tree steps of a Hernquist sphere under each GH_SORT / GH_EMIT
combination in its own process and compares the final positions with the classic sort + thread
emit.  usage: python scripts/gpu_sort_check.py [N] [steps] [mode,mode,...]   (mode = sort+emit)"""
import os, subprocess, sys, tempfile, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
from gravhopper_b200 import Simulation, ic_raw
n, steps = int(sys.argv[2]), int(sys.argv[3])
x, v, m = ic_raw.Hernquist(n, 1.0, 1e10, seed=17)
sim = Simulation(dt=1.0, eps=0.05, algorithm="tree", precision="fp32")
sim.add_IC({"pos": x, "vel": v, "mass": m})
sim.run(steps)
np.save(sys.argv[1], np.asarray(sim.positions.value)[-1])
""" % ROOT

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["place2+warp"]
out = {}
with tempfile.TemporaryDirectory() as tmp:
    for mode in ["classic+thread"] + modes:
        sort, emit = mode.split("+")
        f = os.path.join(tmp, mode.replace("+", "_") + ".npy")
        t0 = time.time()
        r = subprocess.run([sys.executable, "-c", RUN, f, str(n), str(steps)],
                           env=dict(os.environ, GH_SORT=sort, GH_EMIT=emit), capture_output=True, text=True)
        if r.returncode != 0:
            print(mode, "FAILED", r.stderr[-1500:])
            sys.exit(1)
        out[mode] = np.load(f)
        print("%-16s ran in %.1f s" % (mode, time.time() - t0), flush=True)
ref = out["classic+thread"]
ok = bool(np.isfinite(ref).all())
for mode in modes:
    same = bool(np.array_equal(out[mode], ref))
    ok = ok and same
    print("N=%d steps=%d %-16s %s" % (n, steps, mode, "bit-identical to classic+thread" if same else
                                     "DIFFERS: max |dx| = %.3e" % float(np.abs(out[mode] - ref).max())))
sys.exit(0 if ok else 2)
