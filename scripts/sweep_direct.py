"""Sweep launch shapes of the fp32 direct kernel (scratch tool)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gravhopper_b200 import _jbgrav as J, ic_raw

def time_dev(fn, n_iter=3):
    fn(); torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_iter)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return min(a.elapsed_time(b) for a, b in evs)

out = {}
for N in (262144, 1048576):
    x, v, m = ic_raw.Plummer(N, 1e-3, 1e6, seed=42)
    tx = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    for label, tm in (("equal", torch.from_numpy(m).cuda()),
                      ("mixed", torch.from_numpy(m * np.random.default_rng(0).uniform(0.5, 2, N)).cuda())):
        for ki, blk in ((4, 256), (4, 128), (8, 128), (8, 64), (4, 512), (2, 256), (2, 128)):
            for ctas in ((0,) if N > 300000 or label == "mixed" else (0, 148 * 8, 148 * 64)):
                os.environ["GH_F32_KI"] = str(ki); os.environ["GH_F32_BLOCK"] = str(blk)
                if ctas: os.environ["GH_DIRECT_CTAS"] = str(ctas)
                else: os.environ.pop("GH_DIRECT_CTAS", None)
                ms = time_dev(lambda: J.direct_summation(tx, tm, 5e-5, precision="fp32"), 2)
                tf = N * N * 20 / (ms * 1e-3) / 1e12
                out["%d_%s_ki%d_b%d_c%d" % (N, label, ki, blk, ctas)] = tf
                print(N, label, "ki", ki, "block", blk, "ctas", ctas, "%.3f ms %.2f TF(20) %.1f%%" % (ms, tf, 100 * tf / 74.45), flush=True)
json.dump(out, open("gpurun_out/sweep_direct.json", "w"), indent=1)
