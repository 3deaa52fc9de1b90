"""The REAL walk kernels (gravhopper_b200/csrc/walk.cuh), compiled for the host and run warp by warp
with 32 lockstep threads (tests/emu/walk_emu.cpp), against the CPU model of the group criterion and
the oracle's reference tree.  This is how the CPU suite executes kernel code paths a GPU run has not
covered yet (the hybrid rule), and keeps covering the others without a GPU.

The entry array is built here from the Python octree of tests/groupwalk_model.py in the layout
emit_kernel writes (pre-order, (cx,mx,cy,my),(cz,mz,level<<27|skip,m) relative to the root centre).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import groupwalk_model as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "walk_emu.cpp")
LIB = os.path.join(ROOT, "tests", "emu", "libwalk_emu.so")
SKIP_BITS, LEAF_LEVEL = 27, 31


def relerr(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


@pytest.fixture(scope="module")
def emu():
    deps = [SRC, os.path.join(ROOT, "gravhopper_b200", "csrc", "walk.cuh"),
            os.path.join(ROOT, "gravhopper_b200", "csrc", "common.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        out = subprocess.run(["g++", "-O1", "-std=c++20", "-shared", "-fPIC", "-pthread", "-I" + cuda_inc,
                              "-o", LIB, SRC], capture_output=True, text=True)
        if out.returncode != 0:
            pytest.skip("host build of the walk kernels failed: " + out.stderr[-400:])
    lib = C.CDLL(LIB)
    vp, i64 = C.c_void_p, C.c_int64
    lib.emu_walk_group.argtypes = [vp, C.c_int, vp, vp, i64, vp, C.c_float, C.c_double, C.c_int, C.c_float, vp,
                                   vp, C.c_int, vp]
    lib.emu_walk_target.argtypes = [vp, C.c_int, vp, vp, i64, vp, C.c_float, C.c_double, vp, vp, C.c_int]
    lib.emu_walk_target64.argtypes = [vp, vp, C.c_int, vp, vp, i64, vp, C.c_double, C.c_double, vp, vp, C.c_int, vp]
    return lib


def build_entries(x, m, eps):
    """Pre-order fp32 entry array, sorted double4 sources, leaf order and root block, as the GPU
    build produces them."""
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))
    root = G.build(x, m, eps)
    o = root.c.copy()
    rows, order = [], []

    def rec(nd, level):
        e = len(rows)
        rows.append(None)
        if nd.count == 1:
            p = nd.p
            order.append(p)
            packed = (LEAF_LEVEL << SKIP_BITS) | (e + 1)
            rows[e] = (0.0, x[p, 0] - o[0], 0.0, x[p, 1] - o[1], 0.0, x[p, 2] - o[2], packed, m[p])
            return
        for ch in nd.child:
            if ch is not None:
                rec(ch, level + 1)
        com = nd.mx / nd.m
        packed = (level << SKIP_BITS) | len(rows)
        rows[e] = (nd.c[0] - o[0], com[0] - o[0], nd.c[1] - o[1], com[1] - o[1], nd.c[2] - o[2], com[2] - o[2],
                   packed, nd.m)
    rec(root, 0)
    n = len(rows)
    nodes = np.zeros((n, 8), dtype=np.float32)
    bits = nodes.view(np.uint32)
    for e, r in enumerate(rows):
        nodes[e, :6] = r[:6]
        nodes[e, 7] = r[7]
        bits[e, 6] = r[6]
    order = np.array(order, dtype=np.int32)
    sorted4 = np.ascontiguousarray(np.hstack([x[order], m[order, None]]))
    rootblk = np.zeros(10)
    rootblk[:3] = o
    rootblk[3] = root.size
    return nodes, sorted4, order, rootblk


def run_group(lib, nodes, sorted4, order, rootblk, eps, theta, list_limit=3000, kappa=0.0, stats=True, walkctl=None):
    ni = len(order)
    acc = np.zeros((ni, 3))
    st = np.zeros(4, dtype=np.uint64)
    flags = (1 if stats else 0) | (2 if eps == 0.0 else 0) | (4 if kappa > 0 else 0)
    inv_theta2 = float("inf") if theta == 0 else 1.0 / theta ** 2
    lib.emu_walk_group(nodes.ctypes.data, len(nodes), sorted4.ctypes.data, order.ctypes.data, ni,
                       rootblk.ctypes.data, np.float32(eps * eps), inv_theta2, list_limit, np.float32(kappa),
                       acc.ctypes.data, st.ctypes.data, flags,
                       None if walkctl is None else np.ascontiguousarray(walkctl, dtype=np.int32).ctypes.data)
    return acc, dict(accepted=int(st[0]), visited=int(st[1]), iterations=int(st[2]),
                     fallback=int(st[3]) & 0xffffffff, hybrid_targets=int(st[3]) >> 32)


@pytest.fixture(scope="module")
def system():
    from gravhopper_b200 import ic_raw
    x, v, m = ic_raw.Hernquist(3000, 1.0, 1e10, seed=11)
    x = np.ascontiguousarray(x)
    return (x, m, 0.05) + build_entries(x, m, 0.05)


def test_group_kernel_source_equals_the_model(emu, oracle, system):
    x, m, eps, nodes, sorted4, order, rootblk = system
    acc, st = run_group(emu, nodes, sorted4, order, rootblk, eps, 0.7)
    model, info = oracle.tree_force_group(x, m, eps, 0.7)
    assert np.array_equal(info["order"], order)
    assert st["fallback"] == 0 and info["fallback_groups"] == 0
    assert st["accepted"] == info["list_sum"] and st["visited"] == info["tested_sum"]
    assert st["iterations"] == info["iterations"]
    assert relerr(acc, model).max() <= 2e-5          # fp32 sums vs the model's fp64 sums
    # and the per-target kernel against the reference tree itself
    ni = len(order)
    acc_t = np.zeros((ni, 3))
    stt = np.zeros(4, dtype=np.uint64)
    emu.emu_walk_target(nodes.ctypes.data, len(nodes), sorted4.ctypes.data, order.ctypes.data, ni,
                        rootblk.ctypes.data, np.float32(eps * eps), 1.0 / 0.49, acc_t.ctypes.data,
                        stt.ctypes.data, 1)
    ref, so = oracle.tree_force(x, m, eps, 0.7, return_stats=True)
    assert relerr(acc_t, ref).max() <= 2e-5
    assert abs(int(stt[0]) - so["accepted"]) <= 1e-3 * so["accepted"]   # fp32 opening tests: rare flips


def test_group_kernel_give_up_paths(emu, oracle, system):
    x, m, eps, nodes, sorted4, order, rootblk = system
    for limit in (32, 400):
        acc, st = run_group(emu, nodes, sorted4, order, rootblk, eps, 0.7, list_limit=limit)
        model, info = oracle.tree_force_group(x, m, eps, 0.7, list_limit=limit)
        assert st["fallback"] == info["fallback_groups"] > 0
        assert relerr(acc, model).max() <= 2e-5
    # theta = 0 and eps = 0 (guarded kernel): every cell is opened -> direct summation
    acc, st = run_group(emu, nodes, sorted4, order, rootblk, 0.0, 0.0, list_limit=1 << 20)
    assert relerr(acc, oracle.direct_summation(x, m, 0.0)).max() <= 1e-4


def test_hybrid_kernel_path_equals_the_model(emu, oracle, system):
    """walk_group_kernel<..., HYBRID = true>: not yet run on a GPU; here its source runs
    on the CPU and must flag the same targets and produce the same forces as the model's rule."""
    x, m, eps, nodes, sorted4, order, rootblk = system
    kappa = 0.1
    acc, st = run_group(emu, nodes, sorted4, order, rootblk, eps, 0.7, kappa=kappa)
    model, info = oracle.tree_force_group(x, m, eps, 0.7, hybrid=kappa)
    plain, _ = run_group(emu, nodes, sorted4, order, rootblk, eps, 0.7)
    assert info["hybrid_targets"] > 0
    assert abs(st["hybrid_targets"] - info["hybrid_targets"]) <= max(2, 0.05 * info["hybrid_targets"])
    d = relerr(acc, model)
    assert np.median(d) <= 1e-5 and (d > 1e-4).mean() <= 0.005   # a borderline flag may differ
    changed = np.any(acc != plain, axis=1)
    assert 0 < changed.sum() <= 1.2 * info["hybrid_targets"] + 2
    ref = oracle.tree_force(x, m, eps, 0.7)
    assert relerr(acc[changed], ref[changed]).max() <= 2e-5       # re-evaluated = the reference's value


def build_entries64(x, m, eps, theta):
    """fp64 entry array + skip links as emit_kernel<., double> writes them (absolute coordinates)."""
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))
    root = G.build(x, m, eps)
    inv_theta2 = float("inf") if theta == 0 else 1.0 / (theta * theta)
    rows, skips = [], []

    def rec(nd):
        e = len(rows)
        rows.append(None)
        skips.append(0)
        if nd.count == 1:
            p = nd.p
            rows[e] = (0.0, x[p, 0], 0.0, x[p, 1], 0.0, x[p, 2], -1.0, m[p])
            skips[e] = e + 1
            return
        for ch in nd.child:
            if ch is not None:
                rec(ch)
        com = nd.mx / nd.m
        rows[e] = (nd.c[0], com[0], nd.c[1], com[1], nd.c[2], com[2], (nd.size * nd.size) * inv_theta2, nd.m)
        skips[e] = len(rows)
    rec(root)
    rootblk = np.zeros(10)
    rootblk[:3] = root.c
    rootblk[3] = root.size
    return np.array(rows, dtype=np.float64), np.array(skips, dtype=np.int32), rootblk


@pytest.mark.parametrize("theta,eps", [(0.7, 5e-5), (0.3, 5e-5), (0.5, 0.0)])
def test_fp64_walk_kernel_source_equals_the_reference_tree(emu, oracle, golden, theta, eps):
    """walk_kernel<double> (the default-precision tree walk) on the CPU: the reference's node set
    (accepted / visited counts equal the oracle's exactly) and its forces to rounding, for the
    particles themselves and for separate target positions -- the GPU parity test of
    tests/test_gpu_parity.py, executed here on the kernel source."""
    x, m = golden["c1_pos"][:1500], golden["c1_mass"][:1500]
    nodes, skips, rootblk = build_entries64(x, m, eps, theta)
    for tpos in (x, golden["c1_force_pos"]):
        tpos = np.ascontiguousarray(tpos)
        acc = np.zeros_like(tpos)
        st = np.zeros(4, dtype=np.uint64)
        emu.emu_walk_target64(nodes.ctypes.data, skips.ctypes.data, len(nodes), tpos.ctypes.data, None, len(tpos),
                              rootblk.ctypes.data, eps * eps, 1.0 / (theta * theta), acc.ctypes.data,
                              st.ctypes.data, 1 | (2 if eps == 0.0 else 0), None)
        ref, so = oracle.tree_force_position(x, m, tpos, eps, theta, return_stats=True)
        assert len(nodes) == so["nodes"]
        assert int(st[0]) == so["accepted"] and int(st[1]) == so["visited"]
        assert relerr(acc, ref).max() <= 1e-12
