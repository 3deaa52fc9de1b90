"""Cost and accuracy of the opt-in quadrupole extension at N = 4,194,304 (scratch tool): one
tree_force evaluation (device-resident inputs, build included, CUDA events) and the error against
fp64 direct summation on 4096 sampled targets, for the monopole walks and the quadrupole walk."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gravhopper_b200 import _jbgrav as J, ic_raw

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
x, v, m = ic_raw.Hernquist(n, 1.0, 1e10, seed=42)
tx, tm = torch.from_numpy(np.ascontiguousarray(x)).cuda(), torch.from_numpy(m).cuda()
eps, theta = 0.05, 0.7
sel = torch.from_numpy(np.random.default_rng(0).choice(n, 4096, replace=False)).cuda()
d = J.direct_summation_position(tx, tm, tx[sel].contiguous(), eps)


def run(prec):
    J.tree_force(tx, tm, eps, theta, precision=prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        a = J.tree_force(tx, tm, eps, theta, precision=prec)
    e1.record()
    torch.cuda.synchronize()
    e = (torch.linalg.norm(a[sel] - d, dim=1) / torch.linalg.norm(d, dim=1)).cpu().numpy()
    return {"ms": e0.elapsed_time(e1) / 3, "mean": float(e.mean()), "median": float(np.median(e)),
            "p99": float(np.percentile(e, 99)), "max": float(e.max())}


out = {"n": n, "theta": theta}
out["fp32_group_walk_monopole"] = run("fp32")
J.tree_walk("target")
out["fp32_target_walk_monopole"] = run("fp32")
out["fp64_target_walk_monopole"] = run("fp64")
J.tree_quadrupoles(True)
out["fp32_target_walk_quadrupole"] = run("fp32")
out["fp64_target_walk_quadrupole"] = run("fp64")
J.tree_quadrupoles(False)
J.tree_walk("group")
for k, v in out.items():
    print(k, v)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/quadrupoles_N%d.json" % n, "w"), indent=1)
