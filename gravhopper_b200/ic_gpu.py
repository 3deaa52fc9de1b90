"""Device-side IC sampling: ``gh_ic_sample`` behind the same call shapes as
``ic_raw.Plummer / Hernquist / TSIS``.  The inverse-CDF tables are built here exactly as the
reference builds them (/root/reference/gravhopper/gravhopper.py:1469-1477 for Plummer's q table,
:1563-1581 for Hernquist's f(E)); the per-particle sampling runs on the GPU with Philox streams.
The random numbers are NOT numpy's: use ``ic_raw`` when a run must be reproducible against the
host generators, and this module when N is large and only the distribution matters.
"""
import ctypes as C

import numpy as np

from . import _lib, ic_raw

PLUMMER, HERNQUIST, TSIS_KIND = 1, 2, 3
_TABLES = {}


def _plummer_table():
    if "plummer" not in _TABLES:
        qax = np.arange(0, 1.01, 0.01)
        q_prob = qax ** 2 * (1. - qax ** 2) ** (3.5)
        cum = np.cumsum(q_prob)
        cum /= cum[-1]
        _TABLES["plummer"] = (np.ascontiguousarray(cum), np.ascontiguousarray(qax))
    return _TABLES["plummer"]


def _hernquist_table(cutoff):
    key = ("hernquist", float(cutoff))
    if key not in _TABLES:
        from scipy import integrate
        # the reference's grid (:1566) plus one point near the divergence of f(E) at E -> 1 (its
        # discarded np.append, :1573-1575): E_top = 1 - 1e-5 covers r/a >= 1e-5, i.e. every
        # particle of any sample up to ~1e10 particles; deeper potentials clamp to the table end.
        Eax = np.append(np.arange(0.0, 1.0, 0.002), 1.0 - 1e-5)
        cum = np.zeros(len(Eax))
        for k in range(1, len(Eax)):  # piecewise: accurate next to the singularity
            cum[k] = cum[k - 1] + integrate.quad(ic_raw._hernquist_fE, Eax[k - 1], Eax[k], limit=200)[0]
        # a common factor cancels in Einterp(inverse_Einterp(-Phi) * u): normalise to the top
        cum /= cum[-1]
        _TABLES[key] = (np.ascontiguousarray(Eax), np.ascontiguousarray(cum))
    return _TABLES[key]


def _sample(kind, N, params, table, seed, device_out):
    L = _lib.lib()
    _lib.require_gpu()
    prm = (C.c_double * len(params))(*[float(p) for p in params])
    tx, ty = (None, None) if table is None else table
    seed = int(np.random.SeedSequence(seed).generate_state(1, dtype=np.uint64)[0]) if seed is None else int(seed)
    if device_out:
        import torch
        pos = torch.empty((N, 3), dtype=torch.float64, device="cuda")
        vel = torch.empty((N, 3), dtype=torch.float64, device="cuda")
        mass = torch.empty((N,), dtype=torch.float64, device="cuda")
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        args = (C.c_void_p(pos.data_ptr()), C.c_void_p(vel.data_ptr()), C.c_void_p(mass.data_ptr()),
                _lib.GH_MEM_DEVICE, stream)
    else:
        pos, vel, mass = np.empty((N, 3)), np.empty((N, 3)), np.empty(N)
        args = (C.c_void_p(pos.ctypes.data), C.c_void_p(vel.ctypes.data), C.c_void_p(mass.ctypes.data),
                _lib.GH_MEM_HOST, None)
    _lib.check(L.gh_ic_sample(kind, N, prm, len(params),
                              None if tx is None else C.c_void_p(tx.ctypes.data),
                              None if ty is None else C.c_void_p(ty.ctypes.data),
                              0 if tx is None else len(tx), C.c_uint64(seed & (2 ** 64 - 1)), *args),
               "gh_ic_sample")
    return pos, vel, mass


def Plummer(N, b, totmass, seed=None, device_out=False):
    """Isotropic Plummer sphere sampled on the GPU; b in kpc, totmass in Msun."""
    return _sample(PLUMMER, int(N), (b, totmass), _plummer_table(), seed, device_out)


def Hernquist(N, a, totmass, cutoff=10., seed=None, device_out=False):
    """Isotropic Hernquist sphere truncated at cutoff*a, sampled on the GPU."""
    return _sample(HERNQUIST, int(N), (a, totmass, cutoff), _hernquist_table(cutoff), seed, device_out)


def TSIS(N, maxrad, totmass, seed=None, device_out=False):
    """Truncated singular isothermal sphere sampled on the GPU."""
    return _sample(TSIS_KIND, int(N), (maxrad, totmass), None, seed, device_out)
