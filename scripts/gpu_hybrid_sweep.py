"""Hybrid-rule sweep of the fp32 group walk over ALL particles (scratch tool, not a test).

Hernquist N (default 4,194,304), theta = 0.7: per-particle relative acceleration error against
fp64 direct summation for (i) the fp64 per-target walk = the reference's node set
(_jbgrav.c:487-541), (ii) the fp32 group walk with kappa in a list.  Prints mean / median / p99 /
p99.9 / p99.99 / max, the re-evaluated fraction and the time of one evaluation (CUDA events,
device-resident inputs, build included).  usage: gpu_hybrid_sweep.py [N] [kappa ...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gravhopper_b200 import _jbgrav as J, ic_raw  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
kappas = [float(a) for a in sys.argv[2:]] or [0.0, 0.1, 0.15, 0.2, 0.3]
x, v, m = ic_raw.Hernquist(n, 1.0, 1e10, seed=42)
eps, theta = 0.05, 0.7
tx = torch.from_numpy(np.ascontiguousarray(x)).cuda()
tm = torch.from_numpy(m).cuda()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps


def stats(a, d):
    e = (torch.linalg.norm(a - d, dim=1) / torch.linalg.norm(d, dim=1)).cpu().numpy()
    return {"mean": float(e.mean()), "median": float(np.median(e)), "p99": float(np.percentile(e, 99)),
            "p99.9": float(np.percentile(e, 99.9)), "p99.99": float(np.percentile(e, 99.99)), "max": float(e.max())}


d, t_direct = timed(lambda: J.direct_summation(tx, tm, eps), reps=1)
res = {"n": n, "direct_fp64_ms": t_direct}
ref, t_ref = timed(lambda: J.tree_force(tx, tm, eps, theta))
res["reference_criterion_fp64_walk"] = dict(stats(ref, d), ms=t_ref)
print("direct fp64 %.0f ms; fp64 walk %s" % (t_direct, json.dumps(res["reference_criterion_fp64_walk"])), flush=True)
J.tree_walk("target")
a, t = timed(lambda: J.tree_force(tx, tm, eps, theta, precision="fp32"))
res["fp32_target_walk"] = dict(stats(a, d), ms=t)
print("fp32 target walk %s" % json.dumps(res["fp32_target_walk"]), flush=True)
J.tree_walk("group")
for k in kappas:
    J.tree_walk_hybrid(k)
    a, t = timed(lambda: J.tree_force(tx, tm, eps, theta, precision="fp32"))
    J.tree_stats(True)
    J.tree_force(tx, tm, eps, theta, precision="fp32")
    torch.cuda.synchronize()
    st = J.tree_stats()
    J.tree_stats(False)
    r = dict(stats(a, d), ms=t, redo_fraction=st["hybrid_targets"] / float(n), fallback_groups=st["warp_entries_max"])
    res["group_kappa_%g" % k] = r
    print("group kappa %.2f %s" % (k, json.dumps(r)), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/hybrid_sweep_N%d.json" % n, "w"), indent=1)
