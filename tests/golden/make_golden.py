"""Generate the golden fixtures in tests/golden/ from the REFERENCE ITSELF.

Run here (the container with /root/reference): ``python tests/golden/make_golden.py``.
Every array below comes out of the unmodified reference C backend compiled into oracle/_ref by
oracle/Makefile (``_jbgrav.direct_summation`` etc., /root/reference/gravhopper/_jbgrav.c), called
exactly as /root/reference/gravhopper/jbgrav.py:43-48 calls it; the step loop is the unit-free
restatement of gravhopper.py:405-416 in numpy around those calls (astropy is not installed, so the
reference's Python layer itself cannot be imported -- SURVEY F2).  The fixtures pin oracle/ (CPU
tests) and are what the GPU parity tests compare against on the GPU box, where /root/reference
does not exist.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from gravhopper_b200 import ic_raw  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
C_ACC = 4.398600412921223e-09
K = 1.022712165045695e-3
G = 4.30091727003628e-06


def energy(x, v, m, eps):
    ke = 0.5 * np.sum(m * np.sum(v * v, axis=1))
    pe = 0.0
    for i in range(len(m) - 1):
        d = x[i + 1:] - x[i]
        pe -= G * m[i] * np.sum(m[i + 1:] / np.sqrt(np.sum(d * d, axis=1) + eps * eps))
    return ke, pe


def ref_run(ref, x, v, m, dt, eps, nsteps, algorithm, keep, esteps):
    """gravhopper.py:405-416 around the reference C force; returns trajectories of the `keep`
    particles for the first 10 steps and energies at `esteps`."""
    x, v = x.copy(), v.copy()
    traj_x, traj_v, en = [x[keep].copy()], [v[keep].copy()], {0: energy(x, v, m, eps)}
    full10 = None
    for s in range(1, nsteps + 1):
        xh = x + (0.5 * v * dt) * K
        if algorithm == "direct":
            a = ref.direct_summation(xh, m, eps) * C_ACC
        else:
            a = ref.tree_force(xh, m, eps, 0.7) * C_ACC
        v = v + a * dt
        x = xh + (0.5 * v * dt) * K
        if s <= 10:
            traj_x.append(x[keep].copy())
            traj_v.append(v[keep].copy())
        if s == 10:
            full10 = (x.copy(), v.copy())
        if s in esteps:
            en[s] = energy(x, v, m, eps)
    return np.array(traj_x), np.array(traj_v), en, full10


def main():
    ref = O.ref()
    if ref is None:
        O.build()
        ref = O.ref()
    assert ref is not None, "oracle/_ref not built (needs /root/reference)"
    out = {}
    # ---- known-answer tests ----
    p2 = np.array([[0, 0, 0], [1, 0, 0.]])
    m2 = np.array([1., 2.])
    out["kat2_direct"] = ref.direct_summation(p2, m2, 0.0)
    out["kat2_tree"] = ref.tree_force(p2, m2, 0.0, 0.7)
    out["kat1_direct"] = ref.direct_summation(np.array([[1., 2., 3.]]), np.array([5.]), 0.1)
    out["kat1_tree"] = ref.tree_force(np.array([[1., 2., 3.]]), np.array([5.]), 0.1, 0.7)
    # coincident target, eps = 0: position variant guards (-> finite)
    out["kat_coincident_pos"] = ref.direct_summation_position(p2, m2, p2, 0.0)

    # ---- config 1/2: README Plummer N=2000 ----
    N, eps, dt = 2000, 5e-5, 0.005
    x, v, m = ic_raw.Plummer(N, 1e-3, 1e6, seed=42)
    x, v = np.ascontiguousarray(x), np.ascontiguousarray(v)
    out["c1_pos"], out["c1_vel"], out["c1_mass"] = x, v, m
    out["c1_eps"], out["c1_dt"] = eps, dt
    out["c1_acc_direct"] = ref.direct_summation(x, m, eps)
    thetas = np.array([0.0, 0.3, 0.5, 0.7, 1.0])
    out["c1_thetas"] = thetas
    out["c1_acc_tree"] = np.array([ref.tree_force(x, m, eps, th) for th in thetas])
    fp = np.random.default_rng(5).normal(size=(257, 3)) * 3e-3  # arbitrary targets, some outside the core
    out["c1_force_pos"] = fp
    out["c1_acc_direct_pos"] = ref.direct_summation_position(x, m, fp, eps)
    out["c1_acc_tree_pos"] = ref.tree_force_position(x, m, fp, eps, 0.7)
    keep = np.arange(0, N, 31)
    out["c1_keep"] = keep
    esteps = (100, 200, 300, 400)
    for alg in ("direct", "tree"):
        t = time.time()
        tx, tv, en, full10 = ref_run(ref, x, v, m, dt, eps, 400, alg, keep, esteps)
        print(alg, "400 steps: %.1f s" % (time.time() - t), {k: (en[k][0] + en[k][1]) for k in en})
        out["c1_%s_traj_x" % alg], out["c1_%s_traj_v" % alg] = tx, tv
        out["c1_%s_x10" % alg], out["c1_%s_v10" % alg] = full10
        out["c1_%s_energy_steps" % alg] = np.array(sorted(en))
        out["c1_%s_energy" % alg] = np.array([en[k] for k in sorted(en)])

    # ---- a small unequal-mass, off-origin system (exercises fp32 origin handling, mass range) ----
    rng = np.random.default_rng(11)
    xs = rng.normal(size=(777, 3)) * 0.7 + np.array([40.0, -25.0, 10.0])
    ms = 10.0 ** rng.uniform(4, 8, 777)
    out["c0_pos"], out["c0_mass"], out["c0_eps"] = xs, ms, 0.02
    out["c0_acc_direct"] = ref.direct_summation(xs, ms, 0.02)
    out["c0_acc_tree"] = ref.tree_force(xs, ms, 0.02, 0.7)
    out["c0_acc_direct_eps0"] = ref.direct_summation(xs, ms, 0.0)
    out["c0_acc_tree_eps0"] = ref.tree_force(xs, ms, 0.0, 0.5)

    # ---- config 4 analogue, sampled: Hernquist N=20000, 512 targets ----
    xh_, vh_, mh_ = ic_raw.Hernquist(20000, 1.0, 1e10, seed=7)
    xh_ = np.ascontiguousarray(xh_)
    sel = np.random.default_rng(1).choice(20000, 512, replace=False)
    out["c4_seed"], out["c4_N"], out["c4_sel"], out["c4_eps"] = 7, 20000, sel, 0.05
    out["c4_pos_sel"] = xh_[sel]
    out["c4_pos_checksum"] = np.array([xh_.sum(), np.abs(xh_).sum()])
    out["c4_acc_direct_sel"] = ref.direct_summation_position(xh_, mh_, xh_[sel], 0.05)
    out["c4_acc_tree_sel"] = ref.tree_force_position(xh_, mh_, xh_[sel], 0.05, 0.7)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote", os.path.join(HERE, "golden.npz"), os.path.getsize(os.path.join(HERE, "golden.npz")) / 1e3, "kB")


if __name__ == "__main__":
    main()
