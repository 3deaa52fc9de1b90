"""The tree BUILD kernels' real source (csrc/build.cuh, csrc/sortscan.cuh: bbox, keys, radix sort,
levels, scans, emit) run on the CPU (tests/emu/tree_emu.cpp) and, chained with the walk kernels
(tests/emu/walk_emu.cpp), the whole tree path without a GPU:

  * the emitted pre-order entry array is the reference's octree (node count, leaf order, skip links
    and levels exactly; fp64 cell centres and sizes bit for bit; centres of mass to rounding);
  * fp64 build + per-target walk == the oracle's tree force (<= 1e-12, same accepted / visited counts);
  * fp32 build + group walk == the CPU model of the group criterion;
  * coincident particles (the reference segfaults) give the direct-summation answer at theta = 0.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import groupwalk_model as G
from test_walk_emu import build_entries, build_entries64, run_group, relerr, emu  # noqa: F401  (emu: fixture)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "tree_emu.cpp")
LIB = os.path.join(ROOT, "tests", "emu", "libtree_emu.so")


@pytest.fixture(scope="module")
def temu():
    csrc = os.path.join(ROOT, "gravhopper_b200", "csrc")
    deps = [SRC, os.path.join(ROOT, "tests", "emu", "emu_shim.h")] + \
           [os.path.join(csrc, f) for f in ("build.cuh", "sortscan.cuh", "walk.cuh", "common.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        out = subprocess.run(["g++", "-O1", "-std=c++20", "-shared", "-fPIC", "-pthread", "-I" + cuda_inc,
                              "-o", LIB, SRC], capture_output=True, text=True)
        if out.returncode != 0:
            pytest.skip("host build of the tree kernels failed: " + out.stderr[-400:])
    lib = C.CDLL(LIB)
    vp = C.c_void_p
    lib.emu_tree_build.argtypes = [C.c_int, vp, vp, C.c_int64, C.c_double, C.c_double, vp, vp, C.c_int, vp, vp,
                                   vp, vp]
    return lib


def emu_build(lib, prec, x, m, eps, theta):
    n = len(m)
    x, m = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(m, dtype=np.float64)
    cap = 44 * n + 64  # chains of single-child cells included (<= levels per particle)
    nodes = np.zeros((cap, 8), dtype=np.float32 if prec == 32 else np.float64)
    skips = np.zeros(cap, dtype=np.int32)
    sorted4 = np.zeros((n, 4))
    order = np.zeros(n, dtype=np.int32)
    root = np.zeros(10)
    info = np.zeros(2, dtype=np.int32)
    rc = lib.emu_tree_build(prec, x.ctypes.data, m.ctypes.data, n, eps, theta, nodes.ctypes.data, skips.ctypes.data,
                            cap, sorted4.ctypes.data, order.ctypes.data, root.ctypes.data, info.ctypes.data)
    assert rc == 0
    ne = int(info[0])
    return np.ascontiguousarray(nodes[:ne]), np.ascontiguousarray(skips[:ne]), sorted4, order, root, int(info[1])


def test_fp64_build_is_the_reference_octree_and_walks_to_the_oracle(temu, emu, oracle, golden):
    x, m, eps = golden["c1_pos"][:1500], golden["c1_mass"][:1500], float(golden["c1_eps"])
    theta = 0.7
    nodes, skips, sorted4, order, root, maxlevel = emu_build(temu, 64, x, m, eps, theta)
    want, wskips, wroot = build_entries64(x, m, eps, theta)     # the octree built by insertion, in Python
    _, so = oracle.tree_force(x, m, eps, theta, return_stats=True)
    assert len(nodes) == len(want) == so["nodes"]
    assert np.array_equal(skips, wskips)
    assert np.array_equal(root[:4], wroot[:4])                          # bbox midpoint and padded side
    assert np.array_equal(nodes[:, [0, 2, 4, 6]], want[:, [0, 2, 4, 6]])  # centres and side^2/theta^2: bit for bit
    assert np.allclose(nodes[:, 7], want[:, 7], rtol=1e-14, atol=0)     # masses (summed in another order)
    assert np.allclose(nodes[:, [1, 3, 5]], want[:, [1, 3, 5]], rtol=0, atol=1e-15 * abs(x).max() * 16)
    leaves = want[:, 6] < 0
    assert np.array_equal(sorted4[:, :3], want[leaves][:, [1, 3, 5]]) and np.array_equal(x[order], sorted4[:, :3])
    # walk the emitted array with the fp64 kernel: the oracle's forces and node-set counts
    for tpos in (np.ascontiguousarray(x), np.ascontiguousarray(golden["c1_force_pos"])):
        acc = np.zeros_like(tpos)
        st = np.zeros(4, dtype=np.uint64)
        emu.emu_walk_target64(nodes.ctypes.data, skips.ctypes.data, len(nodes), tpos.ctypes.data, None, len(tpos),
                              root.ctypes.data, eps * eps, 1.0 / theta ** 2, acc.ctypes.data, st.ctypes.data, 1)
        ref, s2 = oracle.tree_force_position(x, m, tpos, eps, theta, return_stats=True)
        assert int(st[0]) == s2["accepted"] and int(st[1]) == s2["visited"]
        assert relerr(acc, ref).max() <= 1e-12


def test_fp32_build_and_group_walk_equal_the_model(temu, emu, oracle):
    from gravhopper_b200 import ic_raw
    x, v, m = ic_raw.Hernquist(3000, 1.0, 1e10, seed=11)     # two radix-sort tiles, two scan levels
    x = np.ascontiguousarray(x)
    eps, theta = 0.05, 0.7
    nodes, _, sorted4, order, root, maxlevel = emu_build(temu, 32, x, m, eps, theta)
    want, wsorted, worder, wroot = build_entries(x, m, eps)
    assert len(nodes) == len(want) and np.array_equal(order, worder) and np.array_equal(sorted4, wsorted)
    assert np.array_equal(nodes.view(np.uint32)[:, 6], want.view(np.uint32)[:, 6])   # (level << 27 | skip)
    assert np.allclose(nodes[:, [0, 1, 2, 3, 4, 5, 7]], want[:, [0, 1, 2, 3, 4, 5, 7]], rtol=1e-6,
                       atol=1e-6 * root[3])
    levels = (want.view(np.uint32)[:, 6] >> 27)
    assert maxlevel == levels[levels < 31].max()
    acc, st = run_group(emu, nodes, sorted4, order, root, eps, theta)
    model, info = oracle.tree_force_group(x, m, eps, theta)
    assert st["accepted"] == info["list_sum"] and st["visited"] == info["tested_sum"] and st["fallback"] == 0
    assert relerr(acc, model).max() <= 2e-5


@pytest.mark.parametrize("prec", [32, 64])
def test_coincident_particles_do_not_break_the_build(temu, emu, oracle, prec):
    rng = np.random.default_rng(3)
    base = rng.normal(size=(12, 3))
    x = np.vstack([base, base[:5], base[:2], rng.normal(size=(21, 3))])     # pairs and triples of equal positions
    m = rng.uniform(0.5, 2.0, len(x))
    eps = 0.1
    nodes, skips, sorted4, order, root, maxlevel = emu_build(temu, prec, x, m, eps, 0.0)
    assert sorted(order.tolist()) == list(range(len(x))) and maxlevel <= (21 if prec == 32 else 42)
    d = oracle.direct_summation(x, m, eps)
    if prec == 64:
        tpos = np.ascontiguousarray(x)
        acc = np.zeros_like(tpos)
        st = np.zeros(4, dtype=np.uint64)
        emu.emu_walk_target64(nodes.ctypes.data, skips.ctypes.data, len(nodes), tpos.ctypes.data, None, len(tpos),
                              root.ctypes.data, eps * eps, float("inf"), acc.ctypes.data, st.ctypes.data, 1)
        assert relerr(acc, d).max() <= 1e-12
    else:
        acc, st = run_group(emu, nodes, sorted4, order, root, eps, 0.0)
        assert relerr(acc, d).max() <= 1e-5
