"""Device-side IC sampling: ``gh_ic_sample`` / ``gh_ic_sample_expdisk`` behind the
same call shapes as ``ic_raw.Plummer / Hernquist / TSIS / expdisk``.  The inverse-CDF tables are built here exactly as the
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


def expdisk_tables(sigma0, Rd, external_rotcurve=None):
    """The four radial tables of gh_ic_sample_expdisk, from the same expressions as the host
    generator ic_raw.expdisk (reference: gravhopper.py:1671-1714): grid R = 0 and
    arange(0.001 Rd, 10 Rd, 0.01 Rd); cumulative mass fraction; R*Omega(R); 4 Omega^2 / kappa^2 with
    kappa^2 = 4 Omega^2 + R dOmega^2/dR (central difference, h = 1e-3 kpc)."""
    from scipy import special
    G = ic_raw.G
    Rax = np.arange(0.001 * Rd, 10 * Rd, 0.01 * Rd)
    cum = Rd ** 2 - Rd * np.exp(-Rax / Rd) * (Rax + Rd)
    cum /= cum[-1]

    def om2(rad):
        y = rad / (2. * Rd)
        o2 = np.pi * G * sigma0 / Rd * (special.iv(0, y) * special.kv(0, y) - special.iv(1, y) * special.kv(1, y))
        if external_rotcurve is not None:
            o2 = o2 + (external_rotcurve(rad) / rad) ** 2
        return o2
    h = 1e-3
    lo = np.maximum(Rax - h, 1e-9)
    Omega2 = om2(Rax)
    dom2 = (om2(Rax + h) - om2(lo)) / (Rax + h - lo)
    kappa2 = 4. * Omega2 + Rax * dom2
    vphi = Rax * np.sqrt(Omega2)
    ratio = 4. * Omega2 / kappa2
    c = np.ascontiguousarray
    return (c(np.concatenate(([0.0], Rax))), c(np.concatenate(([0.0], cum))), c(np.concatenate(([0.0], vphi))),
            c(np.concatenate(([ratio[0]], ratio))))


def expdisk(N, sigma0, Rd, z0, sigmaR_Rd, external_rotcurve=None, seed=None, device_out=False):
    """Exponential disk sampled on the GPU; sigma0 in Msun/kpc^2, Rd and z0 in kpc, sigmaR_Rd in
    km/s, external_rotcurve(R kpc) -> km/s (same arguments as ic_raw.expdisk)."""
    L = _lib.lib()
    _lib.require_gpu()
    N = int(N)
    tR, tcum, tvphi, tratio = expdisk_tables(sigma0, Rd, external_rotcurve)
    prm = (C.c_double * 4)(float(sigma0), float(Rd), float(z0), float(sigmaR_Rd))
    seed = int(np.random.SeedSequence(seed).generate_state(1, dtype=np.uint64)[0]) if seed is None else int(seed)
    if device_out:
        import torch
        pos = torch.empty((N, 3), dtype=torch.float64, device="cuda")
        vel = torch.empty((N, 3), dtype=torch.float64, device="cuda")
        mass = torch.empty((N,), dtype=torch.float64, device="cuda")
        out = (C.c_void_p(pos.data_ptr()), C.c_void_p(vel.data_ptr()), C.c_void_p(mass.data_ptr()),
               _lib.GH_MEM_DEVICE, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    else:
        pos, vel, mass = np.empty((N, 3)), np.empty((N, 3)), np.empty(N)
        out = (C.c_void_p(pos.ctypes.data), C.c_void_p(vel.ctypes.data), C.c_void_p(mass.ctypes.data),
               _lib.GH_MEM_HOST, None)
    _lib.check(L.gh_ic_sample_expdisk(N, prm, C.c_void_p(tR.ctypes.data), C.c_void_p(tcum.ctypes.data),
                                      C.c_void_p(tvphi.ctypes.data), C.c_void_p(tratio.ctypes.data), len(tR),
                                      C.c_uint64(seed & (2 ** 64 - 1)), *out), "gh_ic_sample_expdisk")
    return pos, vel, mass
