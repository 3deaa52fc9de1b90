import os, sys
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

N = 262144
x, v, m = ic_raw.Plummer(N, 1e-3, 1e6, seed=42)
tx = torch.from_numpy(np.ascontiguousarray(x)).cuda(); tm = torch.from_numpy(m).cuda()
for ki, blk, unr in ((8, 128, 4), (8, 128, 2), (8, 128, 1), (8, 128, 8), (4, 256, 4), (4, 256, 2), (4, 256, 8), (6, 128, 4), (6, 192, 4), (12, 64, 2), (16, 64, 2), (8, 128, 4)):
    os.environ["GH_F32_KI"] = str(ki); os.environ["GH_F32_BLOCK"] = str(blk); os.environ["GH_F32_UNROLL"] = str(unr)
    try:
        ms = time_dev(lambda: J.direct_summation(tx, tm, 5e-5, precision="fp32"), 3)
    except Exception as e:
        print("ki", ki, "block", blk, "unr", unr, "FAILED", str(e)[:80]); continue
    tf = N * N * 20 / (ms * 1e-3) / 1e12
    print("ki", ki, "block", blk, "unroll", unr, "%.3f ms %.2f TF(20) %.1f%%" % (ms, tf, 100 * tf / 74.45), flush=True)
