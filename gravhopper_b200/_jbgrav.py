"""Drop-in for the reference's CPython extension ``gravhopper._jbgrav``
(/root/reference/gravhopper/_jbgrav.c:25-63 module table): the same four callables, same
positional signatures, same raw units (kpc, Msun, G = 1), same array coercion (anything
convertible to a C-contiguous float64 array, _jbgrav.c:79-80), same RuntimeError messages for
shape errors (_jbgrav.c:92,98,106,228,...), a NEW float64 C-contiguous output shaped like the
positions it was evaluated at (_jbgrav.c:111,267,607,700).

The arithmetic runs in libgravhopper_b200.so on a B200; there is no CPU path.

Extensions (keyword-only, defaults reproduce the reference contract):
    precision : 'fp64' (default; <=1e-12 of the reference) or 'fp32' (fp32 pair maths, <=1e-5)
Torch CUDA tensors are also accepted (float64, contiguous): they are used in place, the result is
a torch tensor on the same device, and the call is asynchronous on torch's current stream.
"""
import ctypes as C

import numpy as np

from . import _lib, _pinned

__all__ = ["direct_summation", "direct_summation_position", "tree_force", "tree_force_position"]

_PREC = {"fp64": _lib.GH_PREC_F64, "f64": _lib.GH_PREC_F64, 64: _lib.GH_PREC_F64,
         "fp32": _lib.GH_PREC_F32, "f32": _lib.GH_PREC_F32, 32: _lib.GH_PREC_F32}


def _prec(precision):
    try:
        return _PREC[precision]
    except KeyError:
        raise ValueError("precision must be 'fp64' or 'fp32'")


def _is_torch_cuda(a):
    return type(a).__module__.startswith("torch") and hasattr(a, "is_cuda") and a.is_cuda


def _as_f64(a):
    # PyArray_FROM_OTF(obj, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY): C-contiguous aligned float64
    if hasattr(a, "value") and hasattr(a, "unit"):
        a = a.value
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _check_pos(pos, who):
    if pos.ndim != 2:
        raise RuntimeError("%s array does not have 2 dimensions." % who)
    if pos.shape[1] != 3:
        raise RuntimeError("%s array is not Nx3." % who)


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _torch_call(fn_name, prec, pos, mass, fpos, eps, theta):
    import torch
    if pos.dtype != torch.float64 or mass.dtype != torch.float64 or not pos.is_contiguous() \
            or not mass.is_contiguous():
        raise ValueError("device tensors must be contiguous float64")
    tgt = pos if fpos is None else fpos
    if fpos is not None and (fpos.dtype != torch.float64 or not fpos.is_contiguous()):
        raise ValueError("device tensors must be contiguous float64")
    out = torch.empty_like(tgt)
    stream = C.c_void_p(torch.cuda.current_stream(pos.device).cuda_stream)
    L = _lib.lib()
    with torch.cuda.device(pos.device):
        args = [prec, C.c_void_p(pos.data_ptr()), C.c_void_p(mass.data_ptr()), pos.shape[0]]
        if fpos is not None:
            args += [C.c_void_p(fpos.data_ptr()), fpos.shape[0]]
        args += [float(eps)]
        if theta is not None:
            args += [float(theta)]
        args += [C.c_void_p(out.data_ptr()), _lib.GH_MEM_DEVICE, stream]
        _lib.check(getattr(L, fn_name)(*args), fn_name)
    return out


def _call(fn_name, pos, mass, fpos, eps, theta, precision, self_names):
    prec = _prec(precision)
    if _is_torch_cuda(pos):
        _check_pos(pos, self_names[0])
        if mass.shape[0] != pos.shape[0]:
            raise RuntimeError("Mass array and %s array contain different numbers of particles."
                               % self_names[0].lower())
        if fpos is not None:
            _check_pos(fpos, "Force position")
        return _torch_call(fn_name, prec, pos, mass, fpos, eps, theta)
    eps = float(eps)
    pos = _as_f64(pos)
    mass = _as_f64(mass)
    _check_pos(pos, self_names[0])
    if mass.ndim < 1 or mass.shape[0] != pos.shape[0]:
        raise RuntimeError("Mass array and %s array contain different numbers of particles."
                           % self_names[0].lower())
    if fpos is not None:
        fpos = _as_f64(fpos)
        _check_pos(fpos, "Force position")
    tgt = pos if fpos is None else fpos
    if tgt.shape[0] == 0:
        return np.empty_like(tgt)
    out = _pinned.empty_f64(tgt.shape)  # PyArray_NewLikeArray: a NEW array (page-locked when large)
    L = _lib.lib()
    args = [prec, _ptr(pos), _ptr(mass), pos.shape[0]]
    if fpos is not None:
        args += [_ptr(fpos), fpos.shape[0]]
    args += [eps]
    if theta is not None:
        args += [float(theta)]
    args += [_ptr(out), _lib.GH_MEM_HOST, None]
    _lib.check(getattr(L, fn_name)(*args), fn_name)
    return out


def direct_summation(pos, mass, eps, *, precision="fp64"):
    """Acceleration on every particle from every other one.  Replaces
    ``_jbgrav.direct_summation`` (_jbgrav.c:68-135 + :140-193)."""
    return _call("gh_direct_summation", pos, mass, None, eps, None, precision, ("Position",))


def direct_summation_position(pos, mass, force_pos, eps, *, precision="fp64"):
    """Acceleration at ``force_pos`` from all particles.  Replaces
    ``_jbgrav.direct_summation_position`` (_jbgrav.c:200-294 + :299-353)."""
    return _call("gh_direct_summation_position", pos, mass, force_pos, eps, None, precision,
                 ("Particle position",))


def tree_force(pos, mass, eps, theta, *, precision="fp64"):
    """Barnes-Hut acceleration on every particle.  Replaces ``_jbgrav.tree_force``
    (_jbgrav.c:564-630 + :737-806); ``theta`` is positional and required at this level, as in
    the reference (the 0.7 default lives in jbgrav.py:52)."""
    return _call("gh_tree_force", pos, mass, None, eps, theta, precision, ("Position",))


def tree_force_position(pos, mass, force_pos, eps, theta, *, precision="fp64"):
    """Barnes-Hut acceleration at ``force_pos``.  Replaces ``_jbgrav.tree_force_position``
    (_jbgrav.c:634-727 + :737-806)."""
    return _call("gh_tree_force_position", pos, mass, force_pos, eps, theta, precision,
                 ("Particle position",))


def tree_stats(enable=None):
    """Enable/disable accepted/visited counting, or return the last tree evaluation's stats."""
    L = _lib.lib()
    if enable is not None:
        _lib.check(L.gh_set_tree_stats(1 if enable else 0))
        return None
    out = (C.c_int64 * 8)()
    _lib.check(L.gh_tree_last_stats(out))
    # group walk: out[6] = groups that gave up (low 32 bits) | targets re-evaluated by the hybrid
    # rule (high 32 bits); per-target walk: the largest per-warp entry count
    return dict(entries=out[0], cells=out[1], maxlevel=out[2], accepted=out[3], visited=out[4],
                warp_entries=out[5], warp_entries_max=out[6] & 0xffffffff, hybrid_targets=out[6] >> 32,
                warps=out[7])


def tree_walk(mode=None):
    """Select / query how fp32 tree evaluations walk the tree: ``"group"`` (default: one traversal
    per 32 Morton-consecutive targets, conservative bounding-box form of the reference's opening
    test) or ``"target"`` (every target applies _jbgrav.c:502 itself).  fp64 always uses "target"."""
    L = _lib.lib()
    if mode is not None:
        if mode not in ("group", "target"):
            raise ValueError("tree walk mode must be 'group' or 'target'")
        _lib.check(L.gh_set_tree_walk(1 if mode == "group" else 0))
        return None
    return "group" if L.gh_get_tree_walk() == 1 else "target"


def tree_walk_hybrid(kappa=None):
    """Set / query the hybrid rule of the fp32 group walk (0 = off, the default): targets whose net
    acceleration is below ``kappa`` times the summed magnitude of their list's contributions are
    re-evaluated with the per-target criterion (see include/gravhopper_b200.h)."""
    L = _lib.lib()
    if kappa is not None:
        _lib.check(L.gh_set_tree_walk_hybrid(float(kappa)))
        return None
    return float(L.gh_get_tree_walk_hybrid())


def tree_quadrupoles(enable=None):
    """Set / query the opt-in quadrupole extension of the tree (an accuracy upgrade BEYOND the
    reference, which is monopole only): accepted cells also contribute their traceless quadrupole.
    Off by default; with it on, tree evaluations use the per-target walk in either precision."""
    L = _lib.lib()
    if enable is not None:
        _lib.check(L.gh_set_tree_quadrupoles(1 if enable else 0))
        return None
    return bool(L.gh_get_tree_quadrupoles())
