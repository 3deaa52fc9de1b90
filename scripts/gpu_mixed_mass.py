"""fp32 direct summation with MIXED masses at N = 2^20 (scratch tool): the uniform-mass tile fast path
cannot trigger, the launch shape is 128 threads x 8 targets (chosen from the host mass array), and
the rate is compared with the uniform-mass case and with the forced 256 x 4 shape."""
import json, os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    from gravhopper_b200 import _jbgrav as J, ic_raw
    import torch
    n = 1 << 20
    x, v, m = ic_raw.Plummer(n, 1e-3, 1e6, seed=42)
    x = np.ascontiguousarray(x)
    if sys.argv[2] == "mixed":
        m = m * np.random.default_rng(1).uniform(0.5, 2.0, n)
    hx, hm = torch.from_numpy(x).pin_memory().numpy(), torch.from_numpy(m).pin_memory().numpy()
    for _ in range(2):
        J.direct_summation(hx, hm, 5e-5, precision="fp32")
    t0 = time.perf_counter()
    for _ in range(3):
        J.direct_summation(hx, hm, 5e-5, precision="fp32")
    dt = (time.perf_counter() - t0) / 3
    print("RESULT %.6f" % dt)
    sys.exit(0)

out = {}
for masses in ("uniform", "mixed"):
    for shape in ("auto", "256x4", "128x8"):
        env = dict(os.environ)
        if shape != "auto":
            b, k = shape.split("x")
            env["GH_F32_BLOCK"], env["GH_F32_KI"] = b, k
        r = subprocess.run([sys.executable, __file__, "child", masses], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        if not line:
            print(masses, shape, "FAILED", r.stderr[-300:])
            continue
        dt = float(line[-1].split()[1])
        rate = (1 << 20) ** 2 / dt
        out["%s_%s" % (masses, shape)] = {"ms_per_call_e2e": dt * 1e3, "interactions_per_s": rate,
                                          "tflops20": rate * 20 / 1e12}
        print(masses, shape, "%.1f ms  %.4g int/s  %.1f TFLOP/s(20)" % (dt * 1e3, rate, rate * 20 / 1e12), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/mixed_mass_N1M.json", "w"), indent=1)
