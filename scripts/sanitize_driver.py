"""Small run through every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gravhopper_b200 as g
from gravhopper_b200 import _jbgrav as J, ic_raw, potentials as P
from gravhopper_b200.units import u
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
x, v, m = ic_raw.Plummer(n, 1e-3, 1e6, seed=1)
x, v = np.ascontiguousarray(x), np.ascontiguousarray(v)
m2 = m * np.random.default_rng(0).uniform(0.5, 2, n)
t = x[:333] * 1.5
for prec in ("fp64", "fp32"):
    for mm in (m, m2):
        for eps in (5e-5, 0.0):
            J.direct_summation(x, mm, eps, precision=prec)
            J.direct_summation_position(x, mm, t, eps, precision=prec)
            J.tree_force(x, mm, eps, 0.7, precision=prec)
            J.tree_force_position(x, mm, t, eps, 0.7, precision=prec)
    for alg in ("direct", "tree"):
        sim = g.Simulation(dt=0.005 * u.Myr, eps=0.05 * u.pc, algorithm=alg, precision=prec)
        sim.add_IC({"pos": x * u.kpc, "vel": v * u.km / u.s, "mass": m * u.Msun})
        sim.add_external_force(P.Hernquist(1e6 * u.Msun, 2 * u.pc))
        sim.run(3)
        sim.add_external_force(lambda pos, args: pos * 0 / u.Myr ** 2 * 0)
        sim.run(2)
        sim.energy()
from gravhopper_b200.sharded import ShardedSimulation
s = ShardedSimulation(x, v, m, 0.005, 5e-5, algorithm="tree", precision="fp32")
s.run(2); s.gather_state()
print("sanitize driver done")
