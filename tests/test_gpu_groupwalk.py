"""GPU tests of the fp32 group walk (walk_group_kernel, csrc/tree.cu), the default fp32 tree walk.

The group walk applies the reference's opening test (_jbgrav.c:502) in a conservative form: a cell
is accepted for the 32 Morton-consecutive targets of a warp only if every point of their bounding
boxes passes the test, so each target's interaction list is a refinement of the node set the
reference accepts for it.  Checked here, through the C ABI:
  * error against direct summation no worse than the reference tree's (golden fixtures) and no
    worse than the per-target walk's, at every theta;
  * list length per target >= the reference's accepted count (refinement);
  * ragged sizes, separate targets, eps = 0, theta = 0, coincident particles;
  * groups that give up (list limit) reproduce the per-target walk;
  * run-to-run determinism.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from gravhopper_b200 import _jbgrav as J, ic_raw

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relerr(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


@pytest.fixture()
def modes():
    """Runs fn() under both walk modes and restores the default."""
    def both(fn):
        out = {}
        try:
            for mode in ("target", "group"):
                J.tree_walk(mode)
                assert J.tree_walk() == mode
                out[mode] = fn()
        finally:
            J.tree_walk("group")
        return out
    return both


def test_group_is_the_default_fp32_walk():
    if os.environ.get("GH_TREE_WALK"):
        pytest.skip("GH_TREE_WALK set in the environment")
    assert J.tree_walk() == "group"
    with pytest.raises(ValueError):
        J.tree_walk("cluster")


def test_group_error_no_worse_than_reference_and_per_target_walk(golden, modes):
    x, m, eps = golden["c1_pos"], golden["c1_mass"], float(golden["c1_eps"])
    direct = golden["c1_acc_direct"]
    for th, acc in zip(golden["c1_thetas"], golden["c1_acc_tree"]):
        th = float(th)
        if th == 0.0:
            continue
        eref = relerr(acc, direct)
        r = modes(lambda: J.tree_force(x, m, eps, th, precision="fp32"))
        eg, et = relerr(r["group"], direct), relerr(r["target"], direct)
        for e in (eg, et):
            assert e.mean() <= eref.mean() * 1.02 + 1e-6, th
            assert np.percentile(e, 99) <= np.percentile(eref, 99) * 1.05 + 1e-6, th
            assert e.max() <= eref.max() * 1.05 + 1e-6, th
        assert eg.mean() <= et.mean() * 1.02 + 1e-6, th  # refinement: fewer, smaller truncations


def test_group_list_is_a_refinement_of_the_reference_set(golden, modes):
    x, m, eps = golden["c1_pos"], golden["c1_mass"], float(golden["c1_eps"])
    J.tree_stats(True)
    try:
        def run():
            J.tree_force(x, m, eps, 0.7, precision="fp32")
            return J.tree_stats()
        st = modes(run)
    finally:
        J.tree_stats(False)
    assert st["group"]["entries"] == st["target"]["entries"]
    # per-target accepted count of the reference's criterion <= list length of the target's group
    assert st["group"]["accepted"] >= st["target"]["accepted"]
    assert st["group"]["accepted"] <= 8 * st["target"]["accepted"]  # and not absurdly longer
    assert st["group"]["warps"] == (len(m) + 31) // 32


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 255, 257, 1000, 5000])
def test_group_ragged_sizes(oracle, modes, n):
    rng = np.random.default_rng(n)
    x = rng.normal(size=(n, 3))
    m = rng.uniform(0.5, 2, n)
    t = rng.normal(size=(n + 3, 3)) * 2
    r = modes(lambda: (J.tree_force(x, m, 0.05, 0.6, precision="fp32"),
                       J.tree_force_position(x, m, t, 0.05, 0.6, precision="fp32"),
                       J.tree_force(x, m, 0.0, 0.6, precision="fp32"),
                       J.tree_force(x, m, 0.05, 0.0, precision="fp32")))
    for a in r["group"]:
        assert np.isfinite(a).all()
    if n == 1:
        assert not r["group"][0].any()
        return
    d = oracle.direct_summation(x, m, 0.05)
    dp = oracle.direct_summation_position(x, m, t, 0.05)
    d0 = oracle.direct_summation(x, m, 0.0)
    # bound: the reference tree's own worst error at this theta (plus fp32 rounding)
    ref = relerr(oracle.tree_force(x, m, 0.05, 0.6), d).max()
    refp = relerr(oracle.tree_force_position(x, m, t, 0.05, 0.6), dp).max()
    ref0 = relerr(oracle.tree_force(x, m, 0.0, 0.6), d0).max()
    a, ap, a0, at0 = r["group"]
    # (a refinement lowers the error on average, not for every single particle: 25 % slack on max)
    assert relerr(a, d).max() <= ref * 1.25 + 2e-5
    assert relerr(ap, dp).max() <= refp * 1.25 + 2e-5
    assert relerr(a0, d0).max() <= ref0 * 1.25 + 1e-3  # unsoftened close pairs amplify fp32 rounding
    assert relerr(at0, d).max() <= 1e-4               # theta = 0 is direct summation


def test_group_survives_coincident_particles(modes):
    x = np.array([[0., 0, 0], [0, 0, 0], [1, 1, 1], [1, 1, 1], [2, 0, 0]] * 20) + \
        np.repeat(np.arange(20.0)[:, None] * 5.0, 5, axis=0)
    m = np.ones(len(x))
    d = J.direct_summation(x, m, 0.1)
    r = modes(lambda: J.tree_force(x, m, 0.1, 0.0, precision="fp32"))
    for a in r.values():
        assert np.isfinite(a).all() and relerr(a, d).max() < 1e-4


def test_group_hernquist_error_distribution(modes):
    n = 200000
    x, v, m = ic_raw.Hernquist(n, 1.0, 1e10, seed=42)
    x = np.ascontiguousarray(x)
    sel = np.random.default_rng(0).choice(n, 4096, replace=False)
    d = J.direct_summation_position(x, m, x[sel], 0.05)          # fp64 direct, <= 1e-12 of the reference
    ref = relerr(J.tree_force_position(x, m, x[sel], 0.05, 0.7), d)  # fp64 walk = the reference's node set
    r = modes(lambda: J.tree_force(x, m, 0.05, 0.7, precision="fp32"))
    for a in r.values():
        e = relerr(a[sel], d)
        assert e.mean() <= ref.mean() * 1.05 + 1e-6
        assert np.percentile(e, 99) <= np.percentile(ref, 99) * 1.10 + 1e-6
    assert relerr(r["group"][sel], d).mean() <= relerr(r["target"][sel], d).mean() * 1.02
    J.tree_walk("group")
    again = J.tree_force(x, m, 0.05, 0.7, precision="fp32")
    assert np.array_equal(again, r["group"])  # deterministic: no atomics on the force path


def test_group_kernel_equals_the_cpu_model_of_its_criterion(oracle):
    """walk_group_kernel against oracle.tree_force_group (an independent C restatement of the same
    criterion: fp32 decisions, fp64 sums): same list and test counts, forces equal to fp32 rounding
    except for the rare group where a borderline acceptance flips."""
    n = 200000
    x, v, m = ic_raw.Hernquist(n, 1.0, 1e10, seed=42)
    x = np.ascontiguousarray(x)
    J.tree_walk("group")
    J.tree_stats(True)
    default = J.tree_walk_hybrid()
    J.tree_walk_hybrid(0.0)   # the plain criterion; the hybrid rule has its own test below
    try:
        a = J.tree_force(x, m, 0.05, 0.7, precision="fp32")
        st = J.tree_stats()
    finally:
        J.tree_stats(False)
        J.tree_walk_hybrid(default)
    model, info = oracle.tree_force_group(x, m, 0.05, 0.7)
    assert info["fallback_groups"] == st["warp_entries_max"] == 0
    assert abs(st["accepted"] - info["list_sum"]) <= 1e-4 * info["list_sum"]
    assert abs(st["visited"] - info["tested_sum"]) <= 1e-4 * info["tested_sum"]
    assert abs(st["warp_entries"] - info["iterations"]) <= 1e-3 * info["iterations"]
    diff = relerr(a, model)
    assert np.median(diff) <= 1e-5
    assert (diff > 1e-4).mean() <= 0.005
    assert diff.max() <= 2e-2


_FALLBACK = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
from gravhopper_b200 import _jbgrav as J, ic_raw
rng = np.random.default_rng(5)
x = rng.uniform(-1.0, 1.0, size=(4096, 3))   # compact: every group's list is hundreds of entries
m = rng.uniform(0.5, 2.0, 4096)
J.tree_stats(True)
J.tree_walk("group")
g = J.tree_force(x, m, 0.01, 0.7, precision="fp32")
sg = J.tree_stats()
J.tree_walk("target")
t = J.tree_force(x, m, 0.01, 0.7, precision="fp32")
err = float(np.max(np.linalg.norm(g - t, axis=1) / np.linalg.norm(t, axis=1)))
print("RESULT", int(np.array_equal(g, t)), sg["warp_entries_max"], sg["warps"], err)
"""


_PARTIAL = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
from gravhopper_b200 import _jbgrav as J, ic_raw
x, v, m = ic_raw.Hernquist(50000, 1.0, 1e10, seed=9)
x = np.ascontiguousarray(x)
J.tree_stats(True)
J.tree_walk("group")
g = J.tree_force(x, m, 0.05, 0.7, precision="fp32")
sg = J.tree_stats()
np.save(sys.argv[1], g)
print("RESULT", sg["warp_entries_max"], sg["warps"])
"""


def test_partial_fallback_matches_the_model(oracle, tmp_path):
    # a 900-entry limit makes about half of the groups give up; the model applies the same rule
    # (limit checked when a 32-entry chunk is evaluated) and must pick the same groups
    env = dict(os.environ, GH_WALK_LIST_LIMIT="900", GH_WALK_HYBRID="0")
    env.pop("GH_TREE_WALK", None)
    f = str(tmp_path / "g.npy")
    out = subprocess.run([sys.executable, "-c", _PARTIAL % ROOT, f], env=env, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT")][-1].split()
    fallbacks, warps = int(line[1]), int(line[2])
    x, v, m = ic_raw.Hernquist(50000, 1.0, 1e10, seed=9)
    model, info = oracle.tree_force_group(np.ascontiguousarray(x), m, 0.05, 0.7, list_limit=900)
    assert warps == info["groups"]
    assert 0 < info["fallback_groups"] < info["groups"]
    assert abs(fallbacks - info["fallback_groups"]) <= max(2, 0.01 * info["fallback_groups"])
    diff = relerr(np.load(f), model)
    assert np.median(diff) <= 1e-5 and (diff > 1e-4).mean() <= 0.01


def test_groups_over_the_list_limit_reproduce_the_per_target_walk():
    # GH_WALK_LIST_LIMIT is read once per process, hence the subprocess.  With a 32-entry limit
    # every group of a 4096-particle uniform cube gives up after its first evaluated chunk and
    # runs the per-target scan: same entries, same arithmetic, same order as walk_kernel<float>.
    env = dict(os.environ, GH_WALK_LIST_LIMIT="32")
    env.pop("GH_TREE_WALK", None)
    out = subprocess.run([sys.executable, "-c", _FALLBACK % ROOT], env=env, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT")][-1].split()
    fallbacks, warps, err = int(line[2]), int(line[3]), float(line[4])
    assert fallbacks == warps  # stats slot 6 counts the groups that fell back in group mode
    assert err <= 1e-6         # (bit-identical when nvcc contracts both instances alike: line[1])


def test_hybrid_rule_matches_the_model_and_repairs_the_tail(oracle):
    n = 200000
    x, v, m = ic_raw.Hernquist(n, 1.0, 1e10, seed=42)
    x = np.ascontiguousarray(x)
    kappa = 0.1
    default = J.tree_walk_hybrid()
    J.tree_walk("group")
    J.tree_stats(True)
    try:
        J.tree_walk_hybrid(kappa)
        assert J.tree_walk_hybrid() == pytest.approx(kappa)
        a = J.tree_force(x, m, 0.05, 0.7, precision="fp32")
        st = J.tree_stats()
    finally:
        J.tree_walk_hybrid(default)
        J.tree_stats(False)
    model, info = oracle.tree_force_group(x, m, 0.05, 0.7, hybrid=kappa)
    assert info["hybrid_targets"] > 0
    assert abs(st["hybrid_targets"] - info["hybrid_targets"]) <= max(5, 0.05 * info["hybrid_targets"])
    assert st["warp_entries_max"] == 0
    diff = relerr(a, model)
    assert np.median(diff) <= 1e-5 and (diff > 1e-4).mean() <= 0.01
    d = J.direct_summation(x, m, 0.05)
    eref = relerr(J.tree_force(x, m, 0.05, 0.7), d)   # fp64 walk = the reference's node set
    e = relerr(a, d)
    assert e.mean() <= eref.mean() and np.percentile(e, 99) <= np.percentile(eref, 99)
    assert np.percentile(e, 99.99) <= np.percentile(eref, 99.99) * 1.05
    assert e.max() <= eref.max() * 1.05
