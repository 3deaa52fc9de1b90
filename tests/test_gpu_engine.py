"""GPU tests of the device-resident leapfrog (Simulation.run) against the reference-driven golden
trajectories and the oracle's restatement of gravhopper.py:405-416."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import gravhopper_b200 as g
from gravhopper_b200 import ic_raw
from gravhopper_b200.units import u

pytestmark = pytest.mark.gpu


def make_sim(golden, alg, precision="fp64", **kw):
    sim = g.Simulation(dt=float(golden["c1_dt"]) * u.Myr, eps=float(golden["c1_eps"]) * u.kpc,
                       algorithm=alg, precision=precision, **kw)
    sim.add_IC({"pos": golden["c1_pos"] * u.kpc, "vel": golden["c1_vel"] * u.km / u.s,
                "mass": golden["c1_mass"] * u.Msun})
    return sim


@pytest.mark.parametrize("alg", ["direct", "tree"])
def test_first_ten_steps_particlewise(golden, alg):
    sim = make_sim(golden, alg)
    sim.run(10)
    keep = golden["c1_keep"]
    pos = np.asarray(sim.positions.value)
    vel = np.asarray(sim.velocities.value)
    assert pos.shape == (11, 2000, 3) and sim.timestep == 10
    tx, tv = golden["c1_%s_traj_x" % alg], golden["c1_%s_traj_v" % alg]
    scale_x, scale_v = np.abs(tx).max(), np.abs(tv).max()
    for s in range(11):
        assert np.abs(pos[s][keep] - tx[s]).max() <= 1e-12 * scale_x, (alg, s)
        assert np.abs(vel[s][keep] - tv[s]).max() <= 1e-12 * scale_v, (alg, s)
    assert np.abs(pos[10] - golden["c1_%s_x10" % alg]).max() <= 1e-12 * scale_x
    assert np.allclose(np.asarray(sim.times.value), np.arange(11) * 0.005, rtol=0, atol=1e-15)


@pytest.mark.parametrize("alg,precision", [("direct", "fp64"), ("tree", "fp64"), ("direct", "fp32"), ("tree", "fp32")])
def test_400_step_energy_drift_matches_reference(golden, oracle, alg, precision):
    """README config: the reference's own DKD integrator error is +14 % over 400 steps
    (BASELINE.md 2.2); the GPU run must show the same drift.  The system is chaotic: re-running the
    REFERENCE arithmetic (oracle) with the initial positions perturbed by 1e-15 relative gives
    identical drifts to 4 digits at steps 100 and 200, 0.112-0.119 at step 300 and 0.136-0.152 at
    step 400 (5 trials; golden: 0.1152, 0.1426).  With fp32-sized noise (6e-8 relative on the
    force inputs each step) the same reference arithmetic spreads to 0.071-0.079 (direct) /
    0.078-0.096 (tree) at step 200 and 0.136-0.164 at step 400 (4 trials each).  The tolerances
    below are those spreads."""
    sim = make_sim(golden, alg, precision)
    sim.run(400)
    m, eps = golden["c1_mass"], float(golden["c1_eps"])
    gsteps = golden["c1_%s_energy_steps" % alg]
    gE = golden["c1_%s_energy" % alg].sum(axis=1)
    pos, vel = np.asarray(sim.positions.value), np.asarray(sim.velocities.value)
    e0 = gE[0]
    tols = {0: 1e-12, 100: 1e-3, 200: 1e-3, 300: 1e-2, 400: 2e-2} if precision == "fp64" else \
           {0: 1e-6, 100: 5e-3, 200: 2e-2, 300: 3e-2, 400: 3e-2}
    for s, want in zip(gsteps, gE):
        ke, pe = oracle.energy(pos[s], vel[s], m, eps, nthreads=0)
        drift_gpu, drift_ref = (ke + pe - e0) / abs(e0), (want - e0) / abs(e0)
        assert abs(drift_gpu - drift_ref) <= tols[int(s)], (s, drift_gpu, drift_ref)
    # the on-device energy diagnostic agrees with the oracle's
    ke_d, pe_d = sim.energy(400)
    ke, pe = oracle.energy(pos[400], vel[400], m, eps, nthreads=0)
    assert abs(ke_d - ke) <= 1e-10 * abs(ke) and abs(pe_d - pe) <= 1e-10 * abs(pe)


def test_continue_run_and_cadence(golden):
    a = make_sim(golden, "direct")
    a.run(6)
    b = make_sim(golden, "direct")
    b.run(2)
    b.run(4)  # continues from the last snapshot (gravhopper.py:306-314)
    assert b.positions.shape == (7, 2000, 3) and b.timestep == 6
    assert np.array_equal(np.asarray(a.positions.value)[6], np.asarray(b.positions.value)[6])
    c = make_sim(golden, "direct", snapshot_every=4)
    c.run(6)  # snapshots after steps 4 and 6
    assert c.positions.shape == (3, 2000, 3)
    assert np.array_equal(np.asarray(c.positions.value)[1], np.asarray(a.positions.value)[4])
    assert np.array_equal(np.asarray(c.positions.value)[2], np.asarray(a.positions.value)[6])
    assert np.allclose(np.asarray(c.times.value), [0, 0.02, 0.03])


def test_external_force_hooks_match_oracle(golden, oracle):
    """Split step: x_half -> host -> Python callbacks -> device (gravhopper.py:455-457)."""
    from gravhopper_b200.units import const
    x, v, m = golden["c1_pos"][:500], golden["c1_vel"][:500], golden["c1_mass"][:500]
    dt, eps = 0.005, 5e-5
    GM = 4.30091727003628e-06 * 1e7  # kpc (km/s)^2
    centre = np.array([0.01, 0.0, 0.0])

    def point_mass(pos, args):  # docs/source/reference usage: Quantities in, acceleration out
        d = pos - args["pos"]
        r2 = (d ** 2).sum(axis=1)
        return -(d / np.sqrt(r2)[:, None]) * (const.G * args["mass"]) / r2[:, None]

    def drag(pos, vel, args):
        return -vel * args["k"]

    def rotating(pos, time, args):
        amp = np.cos(time.to(u.Myr).value)
        return np.tile(np.array([1e-3, 0, 0]) * amp, (len(pos), 1)) * (u.km / u.s / u.Myr)

    sim = g.Simulation(dt=dt * u.Myr, eps=eps * u.kpc, algorithm="direct")
    sim.add_IC({"pos": x * u.kpc, "vel": v * u.km / u.s, "mass": m * u.Msun})
    sim.add_external_force(point_mass, {"mass": 1e7 * u.Msun, "pos": centre * u.kpc})
    sim.add_external_velocitydependent_force(drag, {"k": 0.1 / u.Myr})
    sim.add_external_timedependent_force(rotating, args=None)
    sim.run(3)
    xo, vo, t = x.copy(), v.copy(), 0.0
    K = oracle.KPC_PER_KMS_MYR
    for _ in range(3):
        xh = oracle.half_drift(xo, vo, dt)
        d = xh - centre
        r2 = (d ** 2).sum(axis=1)
        # G M d / r^3 is in (km/s)^2/kpc; 1 (km/s)^2/kpc = K km/s/Myr
        ext = -(d / np.sqrt(r2)[:, None]) * GM / r2[:, None] * K
        ext = ext - vo * 0.1 + np.array([1e-3, 0, 0]) * np.cos(t + 0.5 * dt)
        xo, vo, _ = oracle.leapfrog_step(xo, vo, m, dt, eps, "direct", ext=ext)
        t += dt
    pos = np.asarray(sim.positions.value)[3]
    assert np.abs(pos - xo).max() <= 1e-11 * np.abs(xo).max()


def test_single_particle_and_two_body():
    sim = g.Simulation(dt=1 * u.Myr, eps=0.1 * u.kpc, algorithm="tree")
    sim.add_IC({"pos": np.array([1., 0, 0]) * u.kpc, "vel": np.array([0, 10., 0]) * u.km / u.s,
                "mass": np.array([1e8]) * u.Msun})
    sim.run(3)  # free motion (gravhopper.py:449-450)
    assert np.allclose(np.asarray(sim.positions.value)[3, 0], [1, 30 * 1.022712165045695e-3, 0])


def test_sharded_world1_equals_simulation(golden):
    from gravhopper_b200.sharded import ShardedSimulation
    for alg, prec in (("direct", "fp64"), ("tree", "fp64"), ("direct", "fp32")):
        a = make_sim(golden, alg, prec)
        a.run(3)
        s = ShardedSimulation(golden["c1_pos"], golden["c1_vel"], golden["c1_mass"], float(golden["c1_dt"]),
                              float(golden["c1_eps"]), algorithm=alg, precision=prec)
        s.run(3)
        pos, vel = s.gather_state()
        assert np.array_equal(pos, np.asarray(a.positions.value)[3]), (alg, prec)


def test_native_potentials_equal_python_callbacks(golden, oracle):
    """analytic fields evaluated on the device give the same trajectories as the
    same formulas registered as reference-style Python callbacks, and as the oracle stepping with
    the host-evaluated field."""
    from gravhopper_b200 import potentials as P
    x, v, m = golden["c1_pos"][:400] * 1e3, golden["c1_vel"][:400], golden["c1_mass"][:400]  # spread to kpc scale
    dt, eps = 0.05, 0.05
    pots = [P.PointMass(1e7 * u.Msun, [2.0, 0, 0] * u.kpc, 0.05 * u.kpc), P.Hernquist(1e9 * u.Msun, 1.5 * u.kpc),
            P.LogHalo(150 * u.km / u.s, 0.5 * u.kpc, 0.8), P.MiyamotoNagai(5e9 * u.Msun, 3 * u.kpc, 0.3 * u.kpc)]

    def run(native, alg):
        sim = g.Simulation(dt=dt * u.Myr, eps=eps * u.kpc, algorithm=alg)
        sim.add_IC({"pos": x * u.kpc, "vel": v * u.km / u.s, "mass": m * u.Msun})
        for p in pots:
            if native:
                sim.add_external_force(p)
            else:
                sim.add_external_force(lambda pos, args, p=p: p(pos, None))
        assert len(sim.native_potentials) == (4 if native else 0)
        sim.run(5)
        return np.asarray(sim.positions.value)[5], np.asarray(sim.velocities.value)[5]

    for alg in ("direct", "tree"):
        xn, vn = run(True, alg)
        xc, vc = run(False, alg)
        assert np.abs(xn - xc).max() <= 1e-13 * np.abs(xc).max()
        assert np.abs(vn - vc).max() <= 1e-12 * np.abs(vc).max()
    xo, vo = x.copy(), v.copy()
    for _ in range(5):
        xh = oracle.half_drift(xo, vo, dt)
        ext = sum(p.acceleration(xh) for p in pots)
        xo, vo, _ = oracle.leapfrog_step(xo, vo, m, dt, eps, "direct", ext=ext)
    xn, vn = run(True, "direct")
    assert np.abs(xn - xo).max() <= 1e-12 * np.abs(xo).max()
    # NFW separately (needs log1p on the device)
    sim = g.Simulation(dt=dt * u.Myr, eps=eps * u.kpc, algorithm="direct")
    sim.add_IC({"pos": x * u.kpc, "vel": v * u.km / u.s, "mass": m * u.Msun})
    nfw = P.NFW(1e11 * u.Msun, 10 * u.kpc)
    sim.add_external_force(nfw)
    sim.run(2)
    xo, vo = x.copy(), v.copy()
    for _ in range(2):
        xh = oracle.half_drift(xo, vo, dt)
        xo, vo, _ = oracle.leapfrog_step(xo, vo, m, dt, eps, "direct", ext=nfw.acceleration(xh))
    assert np.abs(np.asarray(sim.positions.value)[2] - xo).max() <= 1e-12 * np.abs(xo).max()


_SORT_RUN = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
from gravhopper_b200 import Simulation, ic_raw
x, v, m = ic_raw.Hernquist(400000, 1.0, 1e10, seed=17)
sim = Simulation(dt=1.0, eps=0.05, algorithm="tree", precision="fp32")
sim.add_IC({"pos": x, "vel": v, "mass": m})
sim.run(6)
np.save(sys.argv[1], np.asarray(sim.positions.value)[-1])
"""


def test_splitter_sort_steps_equal_classic_sort_steps(tmp_path):
    """A running fp32 tree simulation sorts its Morton keys with the buckets the previous step left
    behind (csrc/bucketsort.cuh); the result is the classic LSD sort's, so six steps of N = 400,000
    (293 buckets; dt = 1 Myr moves every particle out of its bucket every step) give bit-identical
    positions under GH_SORT=bucket, GH_SORT=place (one counting + one placing pass with atomics: the
    order inside a bucket is arbitrary on entry, the in-bucket sort orders by (key, index)) and
    GH_SORT=classic; and under GH_EMIT=warp (the warp-cooperative form of the emit kernel, which
    writes the same entry array) as under GH_EMIT=thread."""
    import os
    import subprocess
    import sys
    out = {}
    modes = {"classic": dict(GH_SORT="classic", GH_EMIT="thread"), "bucket": dict(GH_SORT="bucket", GH_EMIT="thread"),
             "place": dict(GH_SORT="place", GH_EMIT="warp"), "warp": dict(GH_SORT="classic", GH_EMIT="warp"),
             "place2": dict(GH_SORT="place2", GH_EMIT="thread"), "place2+warp": dict(GH_SORT="place2", GH_EMIT="warp")}
    for mode, envs in modes.items():
        f = str(tmp_path / (mode.replace("+", "_") + ".npy"))
        env = dict(os.environ, **envs)
        r = subprocess.run([sys.executable, "-c", _SORT_RUN % ROOT, f], env=env, capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        out[mode] = np.load(f)
    assert np.isfinite(out["classic"]).all()
    for mode in modes:
        assert np.array_equal(out[mode], out["classic"]), mode


def test_fp32_origin_moves_with_the_system():
    """The fp32 engine stores coordinates relative to an origin that moves with the mean velocity
    (gh_engine_set_origin_velocity): a system with a bulk velocity of 10^4 km/s drifts 500 scale
    radii in ten steps, yet its internal evolution equals the same system's at rest to fp32
    rounding -- with a fixed origin the pair separations would have lost three digits."""
    x, v, m = ic_raw.Plummer(4096, 1e-3, 1e6, seed=13)
    dt, eps, steps = 0.005, 5e-5, 10
    vb = np.array([1.0e4, -3.0e3, 2.0e3])
    out = {}
    for name, boost in (("rest", np.zeros(3)), ("moving", vb)):
        sim = g.Simulation(dt=dt, eps=eps, algorithm="direct", precision="fp32")
        sim.add_IC({"pos": x, "vel": v + boost, "mass": m})
        sim.run(steps)
        out[name] = np.asarray(sim.positions.value)[-1] - boost * (steps * dt) * 1.022712165045695e-3
    # scale = the typical radius (1e-3 kpc).  Expected difference: fp32 force rounding (1e-6) x the
    # force's share of ten steps' motion (1e-2) = 1e-8 of it; a fixed origin 0.5 kpc away would cost
    # 3e-8 kpc of coordinate resolution, i.e. 1e-3 force errors and 1e-5 here
    scale = np.median(np.linalg.norm(out["rest"] - out["rest"].mean(axis=0), axis=1))
    diff = np.linalg.norm(out["moving"] - out["rest"], axis=1)
    assert np.median(diff) <= 1e-6 * scale          # a fixed origin would give ~1e-5
    assert diff.max() <= 1e-4 * scale               # measured 3.3e-6: one close pair amplifies the rounding
