"""CPU tests of the host-side mirror of the reference interface (no compute: there is no GPU)."""
import numpy as np
import pytest

import gravhopper_b200 as g
from gravhopper_b200 import ic_raw, _lib
from gravhopper_b200.units import u
from gravhopper_b200.sharded import partition


def test_public_surface_matches_reference():
    # /root/reference/gravhopper/__init__.py:3-4 and jbgrav.py:12
    assert hasattr(g, "Simulation") and hasattr(g, "IC") and hasattr(g, "grav")
    assert g.grav.__all__ == ['direct_summation', 'direct_summation_position', 'tree_force',
                              'tree_force_position']
    import inspect
    assert inspect.signature(g.grav.tree_force).parameters["theta"].default == 0.7
    assert inspect.signature(g.grav.tree_force_position).parameters["theta"].default == 0.7
    s = inspect.signature(g.Simulation.__init__).parameters
    assert s["algorithm"].default == "tree"  # gravhopper.py:116
    for name in ("TSIS", "Plummer", "Hernquist", "expdisk"):
        assert hasattr(g.IC, name)


def test_simulation_parameters_and_errors():
    sim = g.Simulation()
    assert sim.get_algorithm() == "tree"
    assert sim.get_dt().to(u.Myr).value == 1.0 and sim.get_eps().to(u.pc).value == 100.0
    with pytest.raises(ValueError, match="dt must have dimensions of time."):
        sim.set_dt(1 * u.kpc)
    with pytest.raises(ValueError, match="eps must have dimensions of length."):
        sim.set_eps(1 * u.Myr)
    with pytest.raises(ValueError, match="algorithm must be 'tree' or 'direct'."):
        sim.set_algorithm("fmm")
    with pytest.raises(g.UninitializedSimulationException):
        sim.run(1)
    assert sim.lenunit == u.kpc and sim.massunit == u.Msun and sim.timeunit == u.Myr
    for k in ("dt", "eps", "algorithm"):
        assert k in sim.params


def test_add_ic_accumulates_and_validates(capsys):
    sim = g.Simulation(dt=0.005 * u.Myr, eps=0.05 * u.pc, algorithm="direct")
    ic = g.IC.Plummer(N=50, b=1 * u.pc, totmass=1e6 * u.Msun, seed=1)
    assert ic["pos"].unit == u.pc and ic["pos"].shape == (50, 3)
    sim.add_IC(ic)
    sim.add_IC({'pos': np.array([10, 0, 0]) * u.kpc, 'vel': np.array([0, 200, 0]) * u.km / u.s,
                'mass': np.array([1e8]) * u.Msun})
    assert len(sim.ICarrays['pos']) == 51
    assert np.allclose(sim.ICarrays['pos'][-1].to(u.kpc).value, [10, 0, 0])
    with pytest.raises(g.ICException):
        sim.add_IC({'pos': np.zeros((2, 3)) * u.kpc, 'vel': np.zeros((2, 3)) * u.km / u.s})
    with pytest.raises(g.ICException):
        sim.add_IC({'pos': np.zeros((2, 3)) * u.kpc, 'vel': np.zeros((3, 3)) * u.km / u.s,
                    'mass': np.ones(2) * u.Msun})
    sim.init_run(4)
    assert sim.positions.shape == (5, 51, 3) and sim.Nsnap == 5 and sim.Np == 51
    assert np.allclose(sim.snap(0)['pos'].value[-1], [10, 0, 0])
    sim.reset()
    assert sim.positions is None and sim.running is False


def test_hooks_are_registered_like_the_reference():
    sim = g.Simulation()
    f = lambda pos, args: pos * 0  # noqa: E731
    sim.add_external_force([f, f], {"a": 1})
    sim.add_external_timedependent_force(lambda p, t, a: p * 0)
    sim.add_external_velocitydependent_force(lambda p, v, a: p * 0)
    assert len(sim.extra_force_functions) == 2
    assert len(sim.extra_timedependent_force_functions) == 1
    assert len(sim.extra_velocitydependent_force_functions) == 1


def test_run_fails_loudly_without_gpu():
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    sim = g.Simulation(dt=0.005 * u.Myr, eps=0.05 * u.pc, algorithm="direct")
    sim.add_IC(g.IC.Plummer(N=16, b=1 * u.pc, totmass=1e6 * u.Msun, seed=1))
    with pytest.raises(_lib.GravHopperB200Error):
        sim.run(2)


def test_ic_generators_follow_reference_statistics():
    x, v, m = ic_raw.Plummer(20000, 1e-3, 1e6, seed=42)
    r = np.linalg.norm(x, axis=1)
    # Plummer half-mass radius = 1.3048 b
    assert abs(np.median(r) / 1e-3 - 1.3048) < 0.03
    assert np.allclose(x.mean(axis=0), 0, atol=1e-12) and np.allclose(v.mean(axis=0), 0, atol=1e-9)
    assert np.allclose(m, 50.0)
    x, v, m = ic_raw.Hernquist(20000, 1.0, 1e10, seed=42)
    r = np.linalg.norm(x, axis=1)
    assert r.max() < 10.5 and abs(np.median(r) - 1.80) < 0.1  # truncated at 10 a: median xi = 50/121
    x, v, m = ic_raw.galaxy_model(5000)
    assert x.shape == (5000, 3) and np.isfinite(x).all() and np.isfinite(v).all()
    assert len(np.unique(m)) == 2


def test_partition():
    assert partition(10, 1) == [(0, 10)]
    assert partition(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert partition(1 << 20, 8) == [(r * (1 << 17), 1 << 17) for r in range(8)]
    with pytest.raises(ValueError):
        partition(3, 4)


def test_morton_layout_is_a_balanced_permutation():
    from gravhopper_b200.sharded import interleaved_layout, morton_order
    x, v, m = ic_raw.Hernquist(50000, 1.0, 1e10, seed=3)
    order = morton_order(x)
    assert sorted(order.tolist()) == list(range(50000))
    # Z-order locality: consecutive particles are close compared with the system size
    d = np.linalg.norm(np.diff(x[order], axis=0), axis=1)
    d_random = np.linalg.norm(np.diff(x, axis=0), axis=1)  # the generator's order is random
    assert np.median(d) < 0.2 * np.median(d_random)
    for world in (1, 2, 3, 8):
        perm = interleaved_layout(x, world)
        assert sorted(perm.tolist()) == list(range(50000))
        parts = partition(50000, world)
        # every rank gets a mix of inner and outer particles (load balance of the walk)
        r = np.linalg.norm(x[perm], axis=1)
        med = [np.median(r[b:b + c]) for b, c in parts]
        assert max(med) < 1.5 * min(med)


def test_pinned_output_pool_recycles_blocks(monkeypatch):
    """gravhopper_b200/_pinned.py with the page-locked allocator swapped for libc malloc: result
    arrays are new, writable, C-contiguous; a block returns to the pool when the last view dies and
    is handed out again; the byte caps fall back to np.empty."""
    import ctypes
    import gc
    from gravhopper_b200 import _pinned
    libc = ctypes.CDLL(None)
    libc.malloc.restype = ctypes.c_void_p
    libc.malloc.argtypes = [ctypes.c_size_t]
    libc.free.argtypes = [ctypes.c_void_p]
    live = set()

    def alloc(n):
        p = libc.malloc(n)
        live.add(p)
        return p

    def free(p):
        live.remove(p)
        libc.free(p)
    monkeypatch.setattr(_pinned, "_raw_alloc", alloc)
    monkeypatch.setattr(_pinned, "_raw_free", free)
    monkeypatch.setattr(_pinned, "ENABLED", True)
    monkeypatch.setattr(_pinned, "stats", {"allocated": 0, "reused": 0, "fallback": 0})
    shape = (100000, 3)
    a = _pinned.empty_f64(shape)
    assert a.shape == shape and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.flags["WRITEABLE"]
    a[:] = 1.5
    p0 = a.ctypes.data
    view = a[10:20]
    b = _pinned.empty_f64(shape)
    assert b.ctypes.data != p0 and _pinned.stats["allocated"] == 2
    del a
    gc.collect()
    c = _pinned.empty_f64(shape)              # `view` still holds the first block
    assert c.ctypes.data != p0 and _pinned.stats["allocated"] == 3
    assert view[0, 0] == 1.5
    del view
    gc.collect()
    d = _pinned.empty_f64(shape)              # now it is recycled
    assert d.ctypes.data == p0 and _pinned.stats["reused"] == 1
    assert _pinned.empty_f64((10, 3)).base is None        # small: plain numpy
    monkeypatch.setattr(_pinned, "MAX_BYTES", _pinned._total)
    e = _pinned.empty_f64((200000, 3))                    # over the cap: plain numpy
    assert e.base is None and _pinned.stats["fallback"] == 1
    del b, c, d, e
    gc.collect()
    _pinned.trim()
    assert not live and _pinned._total == 0 and _pinned._cached == 0
