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
           [os.path.join(csrc, f) for f in ("build.cuh", "sortscan.cuh", "bucketsort.cuh", "walk.cuh", "common.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        out = subprocess.run(["g++", "-O1", "-std=c++20", "-shared", "-fPIC", "-pthread", "-I" + cuda_inc,
                              "-o", LIB, SRC], capture_output=True, text=True)
        if out.returncode != 0:
            pytest.skip("host build of the tree kernels failed: " + out.stderr[-400:])
    lib = C.CDLL(LIB)
    vp = C.c_void_p
    lib.emu_tree_build.argtypes = [C.c_int, vp, vp, C.c_int64, C.c_double, C.c_double, vp, vp, C.c_int, vp, vp,
                                   vp, vp, vp, C.c_int, vp]
    lib.emu_splitter_sort_test.argtypes = [vp, C.c_int64, C.c_int64, vp, C.c_int64, C.c_int, vp, vp, C.c_int, vp]
    lib.emu_tree_build_dist.argtypes = [C.c_int, vp, vp, vp, C.c_int64, C.c_double, C.c_double, vp, C.c_int, vp, vp,
                                        vp, vp, vp, vp, vp, C.c_int]
    lib.emu_cell_end_test.argtypes = [vp, C.c_int64, vp, vp, C.c_int64, vp, vp]
    return lib


def emu_build(lib, prec, x, m, eps, theta, want_keys=False, seg_cap=0, want_quad=False):
    n = len(m)
    x, m = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(m, dtype=np.float64)
    cap = 44 * n + 64  # chains of single-child cells included (<= levels per particle)
    nodes = np.zeros((cap, 8), dtype=np.float32 if prec == 32 else np.float64)
    skips = np.zeros(cap, dtype=np.int32)
    sorted4 = np.zeros((n, 4))
    order = np.zeros(n, dtype=np.int32)
    root = np.zeros(10)
    info = np.zeros(3, dtype=np.int32)
    keys = np.zeros(n, dtype=np.uint64)
    quad = np.zeros((cap, 6), dtype=nodes.dtype) if want_quad else None
    rc = lib.emu_tree_build(prec, x.ctypes.data, m.ctypes.data, n, eps, theta, nodes.ctypes.data, skips.ctypes.data,
                            cap, sorted4.ctypes.data, order.ctypes.data, root.ctypes.data, info.ctypes.data,
                            keys.ctypes.data, seg_cap, None if quad is None else quad.ctypes.data)
    assert rc == 0
    ne = int(info[0])
    out = (np.ascontiguousarray(nodes[:ne]), np.ascontiguousarray(skips[:ne]), sorted4, order, root, int(info[1]))
    if want_quad:
        return out + (np.ascontiguousarray(quad[:ne]),)
    if seg_cap:
        return out + (int(info[2]),)
    return out + (keys,) if want_keys else out


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
                              root.ctypes.data, eps * eps, 1.0 / theta ** 2, acc.ctypes.data, st.ctypes.data, 1, None)
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
                              root.ctypes.data, eps * eps, float("inf"), acc.ctypes.data, st.ctypes.data, 1, None)
        assert relerr(acc, d).max() <= 1e-12
    else:
        acc, st = run_group(emu, nodes, sorted4, order, root, eps, 0.0)
        assert relerr(acc, d).max() <= 1e-5


# ---- distributed build: P ranks, each sorting / scanning / emitting its key range --------
def emu_build_dist(lib, world, split, x, m, eps, theta, stride, walk_blk=0):
    n = len(m)
    x, m = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(m, dtype=np.float64)
    split = np.ascontiguousarray(split, dtype=np.uint64)
    assert len(split) == world + 1
    nodes = np.zeros((world * stride, 8), dtype=np.float32)
    sorted4 = np.zeros((n, 4))
    order = np.zeros(n, dtype=np.int32)
    root = np.zeros(10)
    counts = np.zeros(2 * world, dtype=np.int32)
    split_next = np.zeros(world + 1, dtype=np.uint64)
    maxlevel = np.zeros(2, dtype=np.int32)
    acc = np.zeros((n, 3)) if walk_blk else None
    rc = lib.emu_tree_build_dist(world, split.ctypes.data, x.ctypes.data, m.ctypes.data, n, eps, theta,
                                 nodes.ctypes.data, stride, sorted4.ctypes.data, order.ctypes.data, root.ctypes.data,
                                 counts.ctypes.data, split_next.ctypes.data, maxlevel.ctypes.data,
                                 None if acc is None else acc.ctypes.data, walk_blk)
    out = (rc, nodes, sorted4, order, root, counts.reshape(world, 2), split_next, (int(maxlevel[0]), int(maxlevel[1])))
    return out + (acc,) if walk_blk else out


def compact_segments(nodes, counts, stride):
    """Global virtual-index array (rank r's entries at [r*stride, r*stride + count_r)) -> the contiguous
    pre-order array a single rank would have written, skip links remapped."""
    world = len(counts)
    ne = counts[:, 1].astype(np.int64)
    start = np.concatenate([[0], np.cumsum(ne)])
    parts = [nodes[r * stride:r * stride + ne[r]] for r in range(world)]
    out = np.ascontiguousarray(np.vstack(parts))
    bits = out.view(np.uint32)
    packed = bits[:, 6].astype(np.int64)
    skip = packed & ((1 << 27) - 1)
    rank = np.minimum(skip // stride, world)          # skip == world*stride: the end
    local = skip - rank * stride
    assert np.all((local == 0) | (rank < world)) and np.all(local <= np.append(ne, 0)[rank])
    bits[:, 6] = ((packed >> 27) << 27 | (start[rank] + local)).astype(np.uint32)
    return out


@pytest.mark.parametrize("world,how", [(2, "equal"), (8, "equal"), (4, "random"), (4, "empty"), (5, "tiny")])
def test_distributed_build_is_the_single_rank_tree(temu, emu, world, how):
    from gravhopper_b200 import ic_raw
    n = 1800
    x, v, m = ic_raw.Hernquist(n, 1.0, 1e10, seed=5)
    x = np.ascontiguousarray(x)
    eps, theta = 0.05, 0.7
    nodes1, _, sorted1, order1, root1, maxlevel1, keys = emu_build(temu, 32, x, m, eps, theta, want_keys=True)
    rng = np.random.default_rng(world)
    split = np.zeros(world + 1, dtype=np.uint64)
    split[world] = np.uint64(2 ** 64 - 1)
    if how == "equal":
        split[1:world] = keys[(n * np.arange(1, world)) // world]
    elif how == "random":       # uneven ranges, cut inside deep cells
        split[1:world] = np.sort(keys[rng.choice(n, world - 1, replace=False)] + np.uint64(1))
    elif how == "empty":        # two ranks own nothing (equal splitters; a range below every key)
        split[1] = keys[0]
        split[2] = keys[n // 2]
        split[3] = keys[n // 2]
    else:                       # ranks of one and two particles
        split[1] = keys[1]
        split[2] = keys[3]
        split[3] = keys[n - 2]
        split[4] = keys[n - 1]
    stride = int(len(nodes1) * 1.2) if how in ("random", "empty", "tiny") else int(len(nodes1) / world * 1.3) + 64
    rc, nodes, sorted4, order, root, counts, split_next, (maxlevel, first) = emu_build_dist(temu, world, split, x, m,
                                                                                          eps, theta, stride)
    assert rc == 0
    assert first == stride * int(np.argmax(counts[:, 0] > 0))   # the root entry: first non-empty rank's segment
    assert counts[:, 0].sum() == n and counts[:, 1].sum() == len(nodes1)
    if how == "equal":
        assert abs(counts[:, 0] - n / world).max() <= 1
    if how == "empty":
        assert (counts[:, 0] == 0).sum() == 2
    assert np.array_equal(order, order1) and np.array_equal(sorted4, sorted1) and np.array_equal(root, root1)
    assert maxlevel == maxlevel1
    flat = compact_segments(nodes, counts, stride)
    assert np.array_equal(flat.view(np.uint32)[:, 6], nodes1.view(np.uint32)[:, 6])   # levels and skip links
    assert np.array_equal(flat[:, [0, 2, 4]], nodes1[:, [0, 2, 4]])                   # cell centres
    assert np.allclose(flat[:, [1, 3, 5, 7]], nodes1[:, [1, 3, 5, 7]], rtol=2e-6, atol=2e-6 * root[3])
    # the walk runs on the virtual-index array as it is (chains end at world * stride)
    acc1, st1 = run_group(emu, nodes1, sorted1, order1, root1, eps, theta)
    acc, st = run_group(emu, nodes, sorted4, order, root, eps, theta, walkctl=[0, first])
    assert st["accepted"] == st1["accepted"] and st["visited"] == st1["visited"]
    assert relerr(acc, acc1).max() <= 2e-6
    # next step's key ranges: equal counts to the sampling granularity (1/64 of a rank's particles)
    if how in ("equal", "random"):
        cnt = np.diff(np.searchsorted(keys, split_next[1:world], side="left"), prepend=0, append=n)
        assert abs(cnt - n / world).max() <= max(counts[:, 0].max() / 64 + 2, 3)


def test_segment_overflow_raises_the_flag_and_the_walk_stands_still(temu, emu):
    from gravhopper_b200 import ic_raw
    x, v, m = ic_raw.Hernquist(2000, 1.0, 1e10, seed=6)
    x = np.ascontiguousarray(x)
    nodes, _, sorted4, order, root, maxlevel, overflow = emu_build(temu, 32, x, m, 0.05, 0.7, seg_cap=2500)
    assert overflow == 1 and len(nodes) > 2500
    nodes, _, sorted4, order, root, maxlevel, overflow = emu_build(temu, 32, x, m, 0.05, 0.7, seg_cap=len(nodes) + 7)
    assert overflow == 0
    # a generous capacity: every chain ends at the capacity, entries beyond the fill are never read
    cap = len(nodes) + 7
    padded = np.full((cap, 8), np.nan, dtype=np.float32)
    padded[:len(nodes)] = nodes
    tight = emu_build(temu, 32, x, m, 0.05, 0.7)
    acc1, st1 = run_group(emu, tight[0], tight[2], tight[3], tight[4], 0.05, 0.7)
    acc, st = run_group(emu, padded, sorted4, order, root, 0.05, 0.7)
    assert st["accepted"] == st1["accepted"] and np.array_equal(acc, acc1)


@pytest.mark.parametrize("world,blk,how", [(3, 64, "equal"), (4, 32, "random")])
def test_distributed_walk_deals_the_global_morton_order(temu, emu, world, blk, how):
    """Every rank walks its share of the GLOBAL Morton order (blocks dealt round-robin), the
    accelerations travel through the gathered buffer and the owners pick theirs up: the result per
    particle is the single-rank group walk's (same groups of 32 when blk is a multiple of 32 and
    the ranks' ranges start on group boundaries; otherwise equal to the tree's own accuracy)."""
    from gravhopper_b200 import ic_raw
    n = 1536
    x, v, m = ic_raw.Hernquist(n, 1.0, 1e10, seed=12)
    x = np.ascontiguousarray(x.astype(np.float32).astype(np.float64))   # exactly representable in fp32
    m = m.astype(np.float32).astype(np.float64)
    eps, theta = 0.05, 0.7
    nodes1, _, sorted1, order1, root1, maxlevel1, keys = emu_build(temu, 32, x, m, eps, theta, want_keys=True)
    acc1, st1 = run_group(emu, nodes1, sorted1, order1, root1, eps, theta)
    single = np.zeros_like(acc1)
    single[...] = acc1            # run_group returns accelerations in source order (order1 maps them)
    split = np.zeros(world + 1, dtype=np.uint64)
    split[world] = np.uint64(2 ** 64 - 1)
    rng = np.random.default_rng(7)
    if how == "equal":            # ranges of 512 particles: group boundaries coincide with the single rank's
        split[1:world] = keys[(n * np.arange(1, world)) // world]
    elif how == "random":
        split[1:world] = np.sort(keys[rng.choice(n, world - 1, replace=False)] + np.uint64(1))
    else:
        split[1], split[2], split[3] = keys[0], keys[n // 2], keys[n // 2]
    stride = int(len(nodes1) * 1.2)
    rc, nodes, sorted4, order, root, counts, split_next, _, acc = emu_build_dist(temu, world, split, x, m, eps, theta,
                                                                              stride, walk_blk=blk)
    assert rc == 0 and np.isfinite(acc).all() and (np.abs(acc).sum(axis=1) > 0).all()
    d = relerr(acc, single)
    if how == "equal":
        assert d.max() <= 2e-6     # identical groups, identical lists: fp32 rounding of the entries only
    else:
        assert np.median(d) <= 2e-3 and d.max() <= 0.2   # other groups of 32: the tree's own accuracy


# ---- splitter sort (csrc/bucketsort.cuh): the running simulation's sort ---------------------------------
@pytest.mark.parametrize("place", [0, 1, 5, 37], ids=["partition", "place", "place2-5ctas", "place2-37ctas"])
@pytest.mark.parametrize("case", ["fresh", "stale", "oversize", "duplicates", "partial", "clump", "lowbits", "equal", "adjacent", "manybuckets"])
def test_splitter_sort_equals_the_stable_sort(temu, case, place):
    """Both forms of the splitter sort -- two partition passes by bucket id + one in-shared-memory
    sort per bucket; one counting + one placing pass with atomics + a compact ranking per bucket --
    give the stable sort's result, bit for bit: with this step's own splitters, with another data
    set's (stale) splitters, with buckets far beyond a tile (the CTA-local global-memory path), with
    many equal keys (also of two adjacent values), with thousands of small buckets, with a device-side count below the capacity, with a tight clump in a corner of
    a bucket's key range (long runs of ties in the ranked bits: the long way), with keys that differ
    only below the ranked bits (short runs: insertion), and with buckets of identical keys."""
    rng = np.random.default_rng(11)
    n = 9000
    keys = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64)
    keys[rng.integers(0, n, 600)] >>= np.uint64(30)          # a dense corner: skewed, like Morton keys
    nb, n_real = 300, n
    spl_from = np.sort(keys)
    if case == "stale":
        other = rng.integers(0, 2 ** 63, size=5000, dtype=np.uint64)
        spl_from = np.sort(other)
    elif case == "oversize":                                  # all splitters in the upper half: bucket 0 holds ~4500
        spl_from = np.sort(keys[keys > np.uint64(2 ** 62)])
        nb = 260
    elif case == "duplicates":
        keys[:4000] = keys[rng.integers(0, 40, 4000)]         # 40 distinct values, 100 copies each
        spl_from = np.sort(keys)
    elif case == "partial":
        n_real = 6500
    elif case == "clump":
        # 700 keys within 2^20 of each other inside buckets that span ~2^55 (stale, coarse splitters)
        keys[:700] = np.uint64(2 ** 61) + rng.integers(0, 2 ** 20, 700, dtype=np.uint64)
        keys[700:1400] = np.uint64(3 * 2 ** 60) + rng.integers(0, 4, 700, dtype=np.uint64)   # and heavy ties
        spl_from = np.sort(rng.integers(0, 2 ** 63, size=5000, dtype=np.uint64))
        nb = 257
    elif case == "lowbits":
        # pairs / triples that agree in all but their lowest bits, spread over the whole key space
        base = rng.integers(0, 2 ** 62, size=3000, dtype=np.uint64) << np.uint64(1)
        keys = np.concatenate([base, base + np.uint64(1), base[:1500] + np.uint64(1), base[:1500]])
        rng.shuffle(keys)
        n = n_real = len(keys)
        spl_from = np.sort(keys)
    elif case == "equal":
        keys[:] = keys[rng.integers(0, 3, n)]                 # three distinct values: buckets of one value
        spl_from = np.sort(keys)
    elif case == "manybuckets":
        # 9000 buckets of ~7 pairs: the shared-memory histograms of the place2 form near their limit
        n = n_real = 60000
        keys = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64)
        keys[rng.integers(0, n, 6000)] >>= np.uint64(25)
        spl_from = np.sort(keys)
        nb = 9000
    elif case == "adjacent":
        # 1000 copies each of two ADJACENT key values: the bucket of the first spans no bits at all
        keys[:1000] = np.uint64(5 * 2 ** 60)
        keys[1000:2000] = np.uint64(5 * 2 ** 60 + 1)
        spl_from = np.sort(keys)
    keys = np.ascontiguousarray(keys)
    out_k = np.zeros(n_real, dtype=np.uint64)
    out_v = np.zeros(n_real, dtype=np.int32)
    spl_from = np.ascontiguousarray(spl_from)
    stats = np.zeros(4, dtype=np.int64)
    rc = temu.emu_splitter_sort_test(keys.ctypes.data, n, n_real, spl_from.ctypes.data, len(spl_from), nb,
                                     out_k.ctypes.data, out_v.ctypes.data, place, stats.ctypes.data)
    assert rc == 0
    order = np.argsort(keys[:n_real], kind="stable")
    assert np.array_equal(out_k, keys[:n_real][order])
    assert np.array_equal(out_v, order.astype(np.int32))
    if place:  # the cases reach the paths they were made for
        compact, long_way, oversize, runs = stats
        print(case, stats)
        assert compact > 0 or case == "equal"
        if case in ("clump", "duplicates", "adjacent"):
            assert long_way > 0
        if case in ("oversize", "equal"):
            assert oversize > 0
        if case == "lowbits":
            assert runs > 100 and long_way == 0


def test_quadrupole_extension_equals_its_cpu_model(temu, emu, oracle, golden):
    """Opt-in quadrupoles (beyond the reference): emit_kernel's per-cell traceless
    tensors from the second-moment prefixes + walk_kernel<double, QUAD> reproduce
    oracle.tree_force_quad (the reference's octree and node set, quadrupoles built bottom-up with
    the parallel-axis rule), and the error against direct summation drops well below the monopole
    tree's at the same theta."""
    x, m, eps = golden["c1_pos"][:1500], golden["c1_mass"][:1500], float(golden["c1_eps"])
    theta = 0.7
    nodes, skips, sorted4, order, root, maxlevel, quad = emu_build(temu, 64, x, m, eps, theta, want_quad=True)
    leaves = nodes[:, 6] < 0
    assert not quad[leaves].any() and np.abs(quad[~leaves, 0] + quad[~leaves, 1] + quad[~leaves, 2]).max() <= \
        1e-9 * np.abs(quad).max()                                   # traceless
    tpos = np.ascontiguousarray(x)
    acc = np.zeros_like(tpos)
    st = np.zeros(4, dtype=np.uint64)
    emu.emu_walk_target64(nodes.ctypes.data, skips.ctypes.data, len(nodes), tpos.ctypes.data, None, len(tpos),
                          root.ctypes.data, eps * eps, 1.0 / theta ** 2, acc.ctypes.data, st.ctypes.data, 1,
                          quad.ctypes.data)
    model = oracle.tree_force_quad(x, m, x, eps, theta)
    assert relerr(acc, model).max() <= 1e-9
    d = oracle.direct_summation(x, m, eps)
    mono = oracle.tree_force(x, m, eps, theta)
    assert relerr(acc, d).mean() <= 0.5 * relerr(mono, d).mean()


@pytest.mark.parametrize("n", [31, 32, 33, 1024, 40000, 1100000])
def test_cell_end_from_level_min_tables(temu, n):
    """cell_end (build.cuh): the end of the level-L cell that starts at p is the first q > p whose
    common level drops below L (or the last particle), found in the level-min tables by climbing and
    descending -- against a brute-force scan, for arrays that need 1 to 5 table levels, with long
    plateaus (big cells) and with queries near the end."""
    rng = np.random.default_rng(n)
    clev = rng.integers(3, 22, size=n).astype(np.int8)
    # long stretches above a level (big cells) and a few deep drops
    for _ in range(12):
        a = int(rng.integers(0, n)); b = min(n, a + int(rng.integers(1, max(2, n // 3))))
        clev[a:b] = np.maximum(clev[a:b], rng.integers(6, 15))
    clev[rng.integers(0, n, size=max(1, n // 5000))] = rng.integers(0, 3)
    clev[n - 1] = -1
    nq = 4000 if n < 100000 else 500
    p = rng.integers(0, n, size=nq).astype(np.int64)
    p[:50] = np.maximum(0, n - 1 - np.arange(50) % n)
    level = np.array([rng.integers(0, clev[q] + 1) if clev[q] >= 0 else 0 for q in p], dtype=np.int32)
    out = np.zeros(nq, dtype=np.int64)
    ntab = C.c_int(0)
    temu.emu_cell_end_test(clev.ctypes.data, n, p.ctypes.data, level.ctypes.data, nq, out.ctypes.data, C.byref(ntab))
    expect_tabs = 1
    m = n
    while m > 32:
        m = (m + 31) // 32
        expect_tabs += 1
    assert ntab.value == expect_tabs
    # brute force: next position to the right whose level is below L
    for i in range(nq):
        tail = clev[int(p[i]) + 1:] < int(level[i])
        q = int(p[i]) + 1 + int(np.argmax(tail)) if tail.any() else n
        assert out[i] == min(q, n - 1), (i, p[i], level[i])
