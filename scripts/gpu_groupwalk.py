"""GPU check of the fp32 group walk (scratch tool, not a test): accuracy against direct summation and
against the per-target walk, edge sizes, eps = 0, statistics, and walk-kernel timing at N = 4M."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravhopper_b200 import _jbgrav as J, ic_raw  # noqa: E402
from oracle import oracle as O  # noqa: E402
import torch  # noqa: E402

out = {}


def relerr(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


def both(fn):
    r = {}
    for mode in ("target", "group"):
        J.tree_walk(mode)
        r[mode] = fn()
    J.tree_walk("group")
    return r


# ---- N = 2000 README Plummer --------------------------------------------------------------
p, v, m = ic_raw.Plummer(2000, 1e-3, 1e6, seed=42)
eps = 5e-5
direct = O.direct_summation(p, m, eps)
for th in (0.0, 0.3, 0.5, 0.7, 1.0):
    ref = O.tree_force(p, m, eps, th)
    eref = relerr(ref, direct)
    r = both(lambda: J.tree_force(p, m, eps, th, precision="fp32"))
    for mode, a in r.items():
        e = relerr(a, direct)
        print("N=2000 th=%.1f %-6s mean %.3e p99 %.3e max %.3e | reference tree mean %.3e p99 %.3e max %.3e"
              % (th, mode, e.mean(), np.percentile(e, 99), e.max(), eref.mean(), np.percentile(eref, 99), eref.max()),
              flush=True)
        out["n2000_th%.1f_%s" % (th, mode)] = [float(e.mean()), float(np.percentile(e, 99)), float(e.max())]
J.tree_stats(True)
for mode in ("target", "group"):
    J.tree_walk(mode)
    J.tree_force(p, m, eps, 0.7, precision="fp32")
    print("stats", mode, J.tree_stats(), flush=True)
J.tree_stats(False)
J.tree_walk("group")

# ---- edge sizes, separate targets, eps = 0 ----------------------------------------------------
for n in (1, 2, 3, 31, 32, 33, 255, 257, 1000, 5000):
    rng = np.random.default_rng(n)
    x = rng.normal(size=(n, 3))
    mm = rng.uniform(0.5, 2, n)
    t = rng.normal(size=(n + 3, 3)) * 2
    d = O.direct_summation(x, mm, 0.05)
    dp = O.direct_summation_position(x, mm, t, 0.05)
    r = both(lambda: (J.tree_force(x, mm, 0.05, 0.6, precision="fp32"),
                      J.tree_force_position(x, mm, t, 0.05, 0.6, precision="fp32"),
                      J.tree_force(x, mm, 0.0, 0.6, precision="fp32"),
                      J.tree_force(x, mm, 0.05, 0.0, precision="fp32")))
    d0 = O.direct_summation(x, mm, 0.0) if n > 1 else d
    for mode, (a, ap, a0, at0) in r.items():
        ok = np.isfinite(a).all() and np.isfinite(ap).all() and np.isfinite(a0).all()
        if n > 1:
            print("n=%d %-6s finite=%s self %.2e pos %.2e eps0 %.2e theta0 %.2e" %
                  (n, mode, ok, relerr(a, d).max(), relerr(ap, dp).max(), relerr(a0, d0).max(), relerr(at0, d).max()),
                  flush=True)
        else:
            print("n=1", mode, ok, a, relerr(ap, dp).max(), flush=True)

# ---- N = 200k Hernquist: error distributions vs sampled direct ---------------------------------
p, v, m = ic_raw.Hernquist(200000, 1.0, 1e10, seed=42)
p = np.ascontiguousarray(p)
eps = 0.05
sel = np.random.default_rng(0).choice(200000, 4096, replace=False)
rd = O.direct_summation_position(p, m, p[sel], eps, nthreads=0)
rt = O.tree_force_position(p, m, p[sel], eps, 0.7, nthreads=0)
eref = relerr(rt, rd)
r = both(lambda: J.tree_force(p, m, eps, 0.7, precision="fp32"))
for mode, a in r.items():
    e = relerr(a[sel], rd)
    print("N=200k %-6s mean %.3e p99 %.3e max %.3e | reference tree mean %.3e p99 %.3e max %.3e"
          % (mode, e.mean(), np.percentile(e, 99), e.max(), eref.mean(), np.percentile(eref, 99), eref.max()), flush=True)
    out["n200k_" + mode] = [float(e.mean()), float(np.percentile(e, 99)), float(e.max())]
# determinism
J.tree_walk("group")
a1 = J.tree_force(p, m, eps, 0.7, precision="fp32")
a2 = J.tree_force(p, m, eps, 0.7, precision="fp32")
print("deterministic:", np.array_equal(a1, a2), flush=True)

# ---- N = 4M: statistics and timing -----------------------------------------------------------
N = int(os.environ.get("GW_N", 1 << 22))
p, v, m = ic_raw.Hernquist(N, 1.0, 1e10, seed=42)
tp = torch.from_numpy(np.ascontiguousarray(p)).cuda()
tm = torch.from_numpy(m).cuda()
for mode in ("target", "group"):
    J.tree_walk(mode)
    J.tree_stats(True)
    J.tree_force(tp, tm, eps, 0.7, precision="fp32")
    torch.cuda.synchronize()
    st = J.tree_stats()
    J.tree_stats(False)
    print("N=%d %s stats: accepted/target %.1f tested/target %.1f iters/warp %.1f fallback-or-max %d warps %d"
          % (N, mode, st["accepted"] / N, st["visited"] / N, st["warp_entries"] / st["warps"], st["warp_entries_max"],
             st["warps"]), flush=True)
    out["n4m_stats_" + mode] = st
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        J.tree_force(tp, tm, eps, 0.7, precision="fp32")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("N=%d %s tree_force (build + walk, device resident): %.3f ms" % (N, mode, min(ts)), flush=True)
    out["n4m_ms_" + mode] = min(ts)
for lim in (1200, 1600, 2000, 3200, 4800):
    pass  # list-limit sweep is done with GH_WALK_LIST_LIMIT in separate processes (static in the library)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/groupwalk.json", "w"), indent=1, default=int)
