"""First-light GPU probe (scratch tool, not a test): parity vs the oracle + raw kernel timings."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravhopper_b200 import _jbgrav as J, _lib, ic_raw
from oracle import oracle as O
import torch

out = {}
def relerr(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)

# ---- parity at N=2000 ------------------------------------------------------------------
p, v, m = ic_raw.Plummer(2000, 1e-3, 1e6, seed=42)
eps = 5e-5
ref = O.direct_summation(p, m, eps)
for prec in ("fp64", "fp32"):
    a = J.direct_summation(p, m, eps, precision=prec)
    e = relerr(a, ref)
    out["direct_%s_N2000" % prec] = dict(max=float(e.max()), median=float(np.median(e)))
    print("direct", prec, e.max(), np.median(e), flush=True)
for th in (0.0, 0.3, 0.7, 1.0):
    rt = O.tree_force(p, m, eps, th)
    for prec in ("fp64", "fp32"):
        a = J.tree_force(p, m, eps, th, precision=prec)
        e = relerr(a, rt)
        out["tree_%s_th%.1f_N2000" % (prec, th)] = dict(max=float(e.max()), median=float(np.median(e)))
        print("tree", prec, th, e.max(), np.median(e), flush=True)
J.tree_stats(True)
a = J.tree_force(p, m, eps, 0.7)
_, st = O.tree_force(p, m, eps, 0.7, return_stats=True)
print("stats gpu", J.tree_stats(), "oracle", st, flush=True)
out["tree_stats_N2000"] = dict(gpu=J.tree_stats(), oracle=st)
J.tree_stats(False)

# ---- parity at N=200k (tree, sampled direct) -----------------------------------------------
p, v, m = ic_raw.Hernquist(200000, 1.0, 1e10, seed=42)
eps = 0.05
t = time.time(); rt, st = O.tree_force(p, m, eps, 0.7, nthreads=0, return_stats=True); print("oracle tree 200k", time.time() - t, st, flush=True)
for prec in ("fp64", "fp32"):
    t = time.time(); a = J.tree_force(p, m, eps, 0.7, precision=prec); dt = time.time() - t
    e = relerr(a, rt)
    out["tree_%s_N200k" % prec] = dict(max=float(e.max()), median=float(np.median(e)), wall_s=dt)
    print("tree 200k", prec, e.max(), np.median(e), dt, flush=True)
sel = np.random.default_rng(0).choice(200000, 2048, replace=False)
rd = O.direct_summation_position(p, m, p[sel], eps, nthreads=0)
for prec in ("fp64", "fp32"):
    a = J.direct_summation_position(p, m, p[sel], eps, precision=prec)
    e = relerr(a, rd)
    out["directpos_%s_N200k" % prec] = dict(max=float(e.max()), median=float(np.median(e)))
    print("direct_pos 200k", prec, e.max(), np.median(e), flush=True)

# ---- timings, device resident ---------------------------------------------------------------
def time_dev(fn, n_iter=3):
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_iter)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return min(a.elapsed_time(b) for a, b in evs)

for N in (65536, 262144, 1048576):
    p, v, m = ic_raw.Plummer(N, 1e-3, 1e6, seed=42)
    tp = torch.from_numpy(np.ascontiguousarray(p)).cuda(); tm = torch.from_numpy(m).cuda()
    for prec in ("fp32", "fp64"):
        if prec == "fp64" and N > 262144:
            continue
        for mode in ([0, 1] if prec == "fp32" else [0]):
          os.environ["GH_F32_MODE"] = str(mode)
          for ki in ([1, 2, 4, 8] if prec == "fp32" else [1, 2, 4]):
            os.environ["GH_F32_KI" if prec == "fp32" else "GH_F64_KI"] = str(ki)
            ms = time_dev(lambda: J.direct_summation(tp, tm, 5e-5, precision=prec), 2)
            rate = N * N / (ms * 1e-3)
            out["direct_%s_N%d_ki%d_mode%d" % (prec, N, ki, mode)] = dict(ms=ms, inter_per_s=rate, tflops20=rate * 20 / 1e12)
            print("direct", prec, N, "mode", mode, "ki", ki, "%.3f ms" % ms, "%.3e int/s" % rate, "%.1f TF(20)" % (rate * 20 / 1e12), flush=True)
        os.environ.pop("GH_F32_KI", None); os.environ.pop("GH_F64_KI", None); os.environ.pop("GH_F32_MODE", None)
    for prec in ("fp32", "fp64"):
        ms = time_dev(lambda: J.tree_force(tp, tm, 5e-5, 0.7, precision=prec), 2)
        out["tree_%s_N%d" % (prec, N)] = dict(ms=ms, part_per_s=N / (ms * 1e-3))
        print("tree", prec, N, "%.3f ms" % ms, "%.3e particles/s" % (N / (ms * 1e-3)), flush=True)

json.dump(out, open("gpurun_out/probe1.json", "w"), indent=1, default=int)
