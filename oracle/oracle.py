"""ctypes front-end of the parity checker.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  Nothing under ``gravhopper_b200/`` does.

Two checkers live here:

* ``oracle.*``  -- the C restatement ``oracle/gh_oracle.c`` (``libgh_oracle.so``), whose
  functions cite the reference lines they follow;
* ``ref()``     -- the reference's own, unmodified CPython extension ``_jbgrav`` compiled into
  ``oracle/_ref/`` by ``oracle/Makefile`` (None when it has not been built).

All arrays are float64, C-contiguous, raw units (kpc, Msun, G=1) exactly as at the reference's
``_jbgrav`` level (``/root/reference/gravhopper/_jbgrav.c:68-135``).
"""
import ctypes as C
import importlib.util
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# jbgrav.py:40,48 / gravhopper.py:409 unit factors (astropy CODATA-2018 definitions)
C_ACC = 4.398600412921223e-09          # km/s/Myr per (Msun/kpc^2)
KPC_PER_KMS_MYR = 1.022712165045695e-3  # kpc per (km/s * Myr)
G_KPC_KMS2_MSUN = 4.30091727003628e-06


def build(quiet=True):
    """(Re)build libgh_oracle.so and, when /root/reference exists, oracle/_ref."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libgh_oracle.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        i64 = C.c_int64
        lib.gho_direct.argtypes = [dp, dp, i64, C.c_double, dp, C.c_int]
        lib.gho_direct_position.argtypes = [dp, dp, i64, dp, i64, C.c_double, dp, C.c_int]
        lib.gho_tree_force.argtypes = [dp, dp, i64, dp, i64, C.c_double, C.c_double, dp,
                                       C.POINTER(i64), C.c_int]
        lib.gho_tree_force_quad.argtypes = [dp, dp, i64, dp, i64, C.c_double, C.c_double, dp, C.c_int]
        lib.gho_tree_force_group.argtypes = [dp, dp, i64, C.c_double, C.c_double, C.c_int, C.c_int, dp,
                                             C.POINTER(i64), C.POINTER(C.c_int32), C.POINTER(i64), C.c_int]
        lib.gho_tree_force_group2.argtypes = lib.gho_tree_force_group.argtypes + [C.POINTER(i64), C.c_int, C.c_int, dp,
                                                                                  C.c_double]
        lib.gho_leapfrog_step.argtypes = [dp, dp, dp, i64, C.c_double, C.c_double, C.c_double,
                                          C.c_int, dp, dp, C.c_int]
        lib.gho_half_drift.argtypes = [dp, dp, i64, C.c_double, dp]
        lib.gho_half_drift.restype = None
        lib.gho_energy.argtypes = [dp, dp, dp, i64, C.c_double, dp, C.c_int]
        lib.gho_energy.restype = None
        lib.gho_max_threads.restype = C.c_int
        _LIB = lib
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads():
    return int(_lib().gho_max_threads())


def _check(rc, what):
    if rc == 1:
        raise MemoryError(what)
    if rc == 2:
        raise RecursionError(what + ": coincident particles (the reference would segfault)")
    if rc:
        raise RuntimeError(what)


def direct_summation(pos, mass, eps, nthreads=1):
    pos, mass = _f64(pos), _f64(mass)
    acc = np.empty_like(pos)
    _check(_lib().gho_direct(_p(pos), _p(mass), pos.shape[0], float(eps), _p(acc), nthreads),
           "gho_direct")
    return acc


def direct_summation_position(pos, mass, force_pos, eps, nthreads=1):
    pos, mass, fpos = _f64(pos), _f64(mass), _f64(force_pos)
    acc = np.empty_like(fpos)
    _check(_lib().gho_direct_position(_p(pos), _p(mass), pos.shape[0], _p(fpos), fpos.shape[0],
                                      float(eps), _p(acc), nthreads), "gho_direct_position")
    return acc


def tree_force_position(pos, mass, force_pos, eps, theta, nthreads=1, return_stats=False):
    pos, mass, fpos = _f64(pos), _f64(mass), _f64(force_pos)
    acc = np.empty_like(fpos)
    stats = (C.c_int64 * 4)()
    _check(_lib().gho_tree_force(_p(pos), _p(mass), pos.shape[0], _p(fpos), fpos.shape[0],
                                 float(eps), float(theta), _p(acc), stats, nthreads),
           "gho_tree_force")
    if return_stats:
        return acc, dict(nodes=stats[0], maxdepth=stats[1], accepted=stats[2], visited=stats[3])
    return acc


def tree_force(pos, mass, eps, theta, nthreads=1, return_stats=False):
    return tree_force_position(pos, mass, pos, eps, theta, nthreads, return_stats)


def tree_force_quad(pos, mass, force_pos, eps, theta, nthreads=0):
    """CPU model of the PRODUCT's opt-in quadrupole extension (not a reference function): the
    reference's octree and accepted node set, monopole + traceless quadrupole per accepted cell."""
    pos, mass, fpos = _f64(pos), _f64(mass), _f64(force_pos)
    acc = np.empty_like(fpos)
    _check(_lib().gho_tree_force_quad(_p(pos), _p(mass), pos.shape[0], _p(fpos), fpos.shape[0], float(eps),
                                      float(theta), _p(acc), nthreads), "gho_tree_force_quad")
    return acc


def tree_force_group(pos, mass, eps, theta, list_limit=3000, stack_limit=320, nthreads=0, order=None,
                     group_size=32, two_boxes=True, hybrid=0.0):
    """CPU model of the PRODUCT's fp32 group walk (walk_group_kernel), not a reference function:
    one traversal per 32 depth-first-consecutive targets, cell accepted only if _jbgrav.c:502 holds
    on the targets' two bounding boxes; groups that exceed ``list_limit`` / ``stack_limit`` use the
    per-target walk.  ``hybrid`` = kappa of the kernel's hybrid rule (0: off): targets with
    |acc| < kappa * abs_sum are re-evaluated with the per-target walk.  ``order`` / ``group_size`` /
    ``two_boxes`` are design-study knobs (the kernel: depth-first order, 32, two boxes).
    Returns (acc, info) with info = dict(order, list_len (-1: group gave up), abs_sum, nodes,
    list_sum, tested_sum, iterations, fallback_groups, groups, hybrid_targets)."""
    pos, mass = _f64(pos), _f64(mass)
    n = pos.shape[0]
    acc = np.zeros_like(pos)
    llen = np.zeros(n, dtype=np.int32)
    cabs = np.zeros(n)
    stats = (C.c_int64 * 7)()
    oin = None
    if order is not None:  # design studies: another target order (default: depth-first = Morton)
        order_in = np.ascontiguousarray(order, dtype=np.int64)
        assert order_in.shape == (n,)
        oin = order_in.ctypes.data_as(C.POINTER(C.c_int64))
    order = np.zeros(n, dtype=np.int64)
    _check(_lib().gho_tree_force_group2(_p(pos), _p(mass), n, float(eps), float(theta), int(list_limit),
                                        int(stack_limit), _p(acc), order.ctypes.data_as(C.POINTER(C.c_int64)),
                                        llen.ctypes.data_as(C.POINTER(C.c_int32)), stats, nthreads, oin,
                                        int(group_size), 1 if two_boxes else 0, _p(cabs), float(hybrid)),
           "gho_tree_force_group")
    return acc, dict(order=order, list_len=llen, abs_sum=cabs, nodes=stats[0], list_sum=stats[1], tested_sum=stats[2],
                     iterations=stats[3], fallback_groups=stats[4], groups=stats[5], hybrid_targets=stats[6])


def leapfrog_step(x, v, mass, dt, eps, algorithm="direct", theta=0.7, ext=None, nthreads=1):
    """One DKD step in internal units (kpc, km/s, Msun, Myr); returns (x_new, v_new, x_half)."""
    x, v, mass = _f64(x).copy(), _f64(v).copy(), _f64(mass)
    xh = np.empty_like(x)
    e = None if ext is None else _f64(ext)
    rc = _lib().gho_leapfrog_step(_p(x), _p(v), _p(mass), x.shape[0], float(dt), float(eps),
                                  float(theta), 0 if algorithm == "direct" else 1,
                                  None if e is None else _p(e), _p(xh), nthreads)
    _check(rc, "gho_leapfrog_step")
    return x, v, xh


def half_drift(x, v, dt):
    x, v = _f64(x), _f64(v)
    xh = np.empty_like(x)
    _lib().gho_half_drift(_p(x), _p(v), x.shape[0], float(dt), _p(xh))
    return xh


def energy(x, v, mass, eps, nthreads=0):
    """(KE, PE) in Msun (km/s)^2 consistent with the softened force law."""
    x, v, mass = _f64(x), _f64(v), _f64(mass)
    out = np.zeros(2)
    _lib().gho_energy(_p(x), _p(v), _p(mass), x.shape[0], float(eps), _p(out), nthreads)
    return float(out[0]), float(out[1])


_REF = False


def ref():
    """The reference's own compiled ``_jbgrav`` module from oracle/_ref, or None."""
    global _REF
    if _REF is False:
        cands = glob.glob(os.path.join(_HERE, "_ref", "_jbgrav*.so"))
        if not cands:
            _REF = None
        else:
            spec = importlib.util.spec_from_file_location("_jbgrav", cands[0])
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _REF = mod
    return _REF
