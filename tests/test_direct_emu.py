"""The direct-summation kernels' real source (csrc/direct.cuh: direct_f64_kernel, direct_f32_kernel,
finalize_kernel, the split heuristic) run on the CPU (tests/emu/direct_emu.cpp) against the golden
outputs of the reference's own C backend and the oracle: the GPU parity tests of
tests/test_gpu_parity.py / test_gpu_engine.py, executed on the kernel source without a GPU.
Tolerances are the contract's: fp64 <= 1e-12, fp32 <= 1e-5 relative per-particle acceleration error.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "direct_emu.cpp")
LIB = os.path.join(ROOT, "tests", "emu", "libdirect_emu.so")


def relerr(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


@pytest.fixture(scope="module")
def demu():
    csrc = os.path.join(ROOT, "gravhopper_b200", "csrc")
    deps = [SRC, os.path.join(ROOT, "tests", "emu", "emu_shim.h"), os.path.join(csrc, "direct.cuh"),
            os.path.join(csrc, "common.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        out = subprocess.run(["g++", "-O1", "-std=c++20", "-shared", "-fPIC", "-pthread", "-I" + cuda_inc,
                              "-o", LIB, SRC], capture_output=True, text=True)
        if out.returncode != 0:
            pytest.skip("host build of the direct kernels failed: " + out.stderr[-400:])
    lib = C.CDLL(LIB)
    vp = C.c_void_p
    lib.emu_direct.argtypes = [C.c_int, vp, vp, C.c_int64, vp, C.c_int64, C.c_double, C.c_int, C.c_int, vp, vp, vp,
                               C.c_double, vp, vp, vp, vp]
    lib.emu_potentials.argtypes = [C.c_int, vp, vp, vp, C.c_int64, vp]
    return lib


def direct64(lib, x, m, t, eps, shape=1128):
    x, m, t = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, m, t))
    acc = np.zeros_like(t)
    info = np.zeros(2, dtype=np.int32)
    lib.emu_direct(64, x.ctypes.data, m.ctypes.data, len(m), t.ctypes.data, len(t), eps, shape, 0, acc.ctypes.data,
                   None, None, 0.0, None, None, None, info.ctypes.data)
    return acc, info


def pack32(x, m, origin):
    out = np.zeros((len(x), 4), dtype=np.float32)
    out[:, :3] = (x - origin).astype(np.float32)
    if m is not None:
        out[:, 3] = m.astype(np.float32)
    return out


def direct32(lib, x, m, t, eps, shape=1128):
    origin = x.mean(axis=0)      # the engine packs relative to the sources' mean position
    s32, t32 = pack32(x, m, origin), pack32(t, None, origin)
    acc = np.zeros((len(t), 3))
    info = np.zeros(2, dtype=np.int32)
    lib.emu_direct(32, s32.ctypes.data, None, len(m), t32.ctypes.data, len(t), eps, shape, 0, acc.ctypes.data,
                   None, None, 0.0, None, None, None, info.ctypes.data)
    return acc, info


@pytest.mark.parametrize("ctas", ["1", ""])
def test_fp64_kernel_source_against_the_reference(demu, golden, monkeypatch, ctas):
    # GH_DIRECT_CTAS=1: one source chunk, epilogue inside the force kernel; default: the sources are
    # split into chunks, partial sums + finalize kernel
    if ctas:
        monkeypatch.setenv("GH_DIRECT_CTAS", ctas)
    x, m, eps = golden["c1_pos"], golden["c1_mass"], float(golden["c1_eps"])
    acc, info = direct64(demu, x, m, x, eps)
    assert (info[0] == 1) == bool(ctas)
    assert relerr(acc, golden["c1_acc_direct"]).max() <= 1e-12
    ap, _ = direct64(demu, x, m, golden["c1_force_pos"], eps, shape=2256)
    assert relerr(ap, golden["c1_acc_direct_pos"]).max() <= 1e-12
    # unequal masses far from the origin, eps = 0 (guarded kernel)
    xs, ms = golden["c0_pos"], golden["c0_mass"]
    a0, _ = direct64(demu, xs, ms, xs, 0.0, shape=4128)
    assert relerr(a0, golden["c0_acc_direct_eps0"]).max() <= 1e-12
    a1, _ = direct64(demu, xs, ms, xs, float(golden["c0_eps"]))
    assert relerr(a1, golden["c0_acc_direct"]).max() <= 1e-12


@pytest.mark.parametrize("shape", [1128, 2128, 4256])
def test_fp32_kernel_source_within_1e5(demu, golden, shape):
    x, m, eps = golden["c1_pos"], golden["c1_mass"], float(golden["c1_eps"])     # equal masses: uniform tiles
    acc, _ = direct32(demu, x, m, x, eps, shape)
    assert relerr(acc, golden["c1_acc_direct"]).max() <= 1e-5
    xs, ms = golden["c0_pos"], golden["c0_mass"]                                 # unequal masses: general tiles
    a1, _ = direct32(demu, xs, ms, xs, float(golden["c0_eps"]), shape)
    assert relerr(a1, golden["c0_acc_direct"]).max() <= 1e-5
    t = golden["c1_force_pos"]
    ap, _ = direct32(demu, x, m, t, eps, shape)
    assert relerr(ap, golden["c1_acc_direct_pos"]).max() <= 1e-5


def test_fused_leapfrog_epilogue_equals_the_oracle_step(demu, golden, oracle, monkeypatch):
    """EP_STEP inside the fp64 force kernel (and inside finalize when the sources are split): kick and
    both drifts of gravhopper.py:409-416, against the oracle's restatement of perform_timestep."""
    x, v, m = golden["c1_pos"][:1000], golden["c1_vel"][:1000], golden["c1_mass"][:1000]
    eps, dt = float(golden["c1_eps"]), 0.005
    xh = oracle.half_drift(x, v, dt)
    xo, vo, _ = oracle.leapfrog_step(x, v, m, dt, eps, "direct")
    for ctas in ("1", "4736"):
        monkeypatch.setenv("GH_DIRECT_CTAS", ctas)
        xn, vn, xhn = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
        info = np.zeros(2, dtype=np.int32)
        mm, vv = np.ascontiguousarray(m), np.ascontiguousarray(v)
        demu.emu_direct(64, xh.ctypes.data, mm.ctypes.data, len(m), xh.ctypes.data, len(m), eps, 1128, 1, None,
                        xh.ctypes.data, vv.ctypes.data, dt, xn.ctypes.data, vn.ctypes.data, xhn.ctypes.data,
                        info.ctypes.data)
        assert np.abs(xn - xo).max() <= 1e-13 * np.abs(xo).max()
        assert np.abs(vn - vo).max() <= 1e-12 * np.abs(vo).max()
        assert np.abs(xhn - oracle.half_drift(xn, vn, dt)).max() <= 1e-15 * np.abs(xo).max()


def test_ten_fused_steps_follow_the_reference_trajectory(demu, golden, oracle, monkeypatch):
    """README Plummer N = 2000: ten DKD steps through the fused fp64 kernel (the engine's loop: each
    step's epilogue also writes the next step's x_half) against the golden trajectory that
    tests/golden/make_golden.py produced with the reference's own compiled C force."""
    monkeypatch.setenv("GH_DIRECT_CTAS", "1")
    x, v, m = (np.ascontiguousarray(golden[k]) for k in ("c1_pos", "c1_vel", "c1_mass"))
    eps, dt, keep = float(golden["c1_eps"]), float(golden["c1_dt"]), golden["c1_keep"]
    xh = oracle.half_drift(x, v, dt)
    info = np.zeros(2, dtype=np.int32)
    for step in range(1, 11):
        xn, vn, xhn = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
        demu.emu_direct(64, xh.ctypes.data, m.ctypes.data, len(m), xh.ctypes.data, len(m), eps, 1128, 1, None,
                        xh.ctypes.data, v.ctypes.data, dt, xn.ctypes.data, vn.ctypes.data, xhn.ctypes.data,
                        info.ctypes.data)
        x, v, xh = xn, vn, xhn
        sx, sv = np.abs(golden["c1_direct_traj_x"][step]).max(), np.abs(golden["c1_direct_traj_v"][step]).max()
        assert np.abs(x[keep] - golden["c1_direct_traj_x"][step]).max() <= 1e-12 * sx
        assert np.abs(v[keep] - golden["c1_direct_traj_v"][step]).max() <= 1e-12 * sv
    assert np.abs(x - golden["c1_direct_x10"]).max() <= 1e-12 * np.abs(golden["c1_direct_x10"]).max()


def test_native_potentials_source_equals_the_host_formulas(demu):
    """gh::eval_potentials (what the engine's potentials kernel evaluates at x_half on the device,
    SURVEY 8f rank 1) against the host-side formulas of gravhopper_b200/potentials.py, which double
    as reference-style callbacks: all five kinds, off-centre, summed."""
    from gravhopper_b200 import potentials as P
    rng = np.random.default_rng(8)
    x = rng.normal(size=(500, 3)) * 5.0
    x[0] = [0.3, -0.2, 0.1]
    pots = [P.PointMass(1e9, center=[0.3, -0.2, 0.1], softening=0.05), P.Hernquist(1e11, 8.0, center=[1.0, 0.0, -0.5]),
            P.NFW(3e11, 15.0), P.LogHalo(180.0, rc=2.0, q=0.8), P.MiyamotoNagai(5e10, 3.0, 0.3, center=[0.0, 0.5, 0.0])]
    for group in ([p] for p in pots):
        _check_pots(demu, group, x)
    _check_pots(demu, pots[:4], x)          # GH_MAX_POTENTIALS = 4 at once


def _check_pots(demu, group, x):
    kinds = np.array([p.kind for p in group], dtype=np.int32)
    prm = np.ascontiguousarray(np.array([p.params() for p in group], dtype=np.float64))
    out = np.zeros_like(x)
    xx = np.ascontiguousarray(x)
    assert demu.emu_potentials(len(group), kinds.ctypes.data, prm.ctypes.data, xx.ctypes.data, len(x),
                               out.ctypes.data) == 0
    want = sum(p.acceleration(x) for p in group)
    assert np.isfinite(out).all()
    scale = np.linalg.norm(want, axis=1).max()
    assert np.abs(out - want).max() <= 1e-12 * scale
