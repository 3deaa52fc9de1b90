"""The reference's README example (/root/reference/README.md:30-57) on gravhopper_b200:
an equilibrium Plummer sphere, 2000 particles, 400 steps.  Needs a CUDA device."""
import numpy as np

from gravhopper_b200 import Simulation, IC
from gravhopper_b200.units import u

Plummer_IC = IC.Plummer(b=1 * u.pc, totmass=1e6 * u.Msun, N=2000, seed=42)
sim = Simulation(dt=0.005 * u.Myr, eps=0.05 * u.pc)        # algorithm='tree' by default, as upstream
sim.add_IC(Plummer_IC)
sim.run(400)

r0 = np.linalg.norm(np.asarray(sim.positions[0].to(u.pc).value), axis=1)
r1 = np.linalg.norm(np.asarray(sim.positions[-1].to(u.pc).value), axis=1)
print("half-mass radius: t=0 %.3f pc, t=%.1f Myr %.3f pc" % (np.median(r0), sim.times[-1].value, np.median(r1)))
ke0, pe0 = sim.energy(0)
ke1, pe1 = sim.energy(400)
print("energy drift over 400 steps: %+.3f (the reference's own integrator error at this dt is +0.14)"
      % ((ke1 + pe1 - ke0 - pe0) / abs(ke0 + pe0)))
