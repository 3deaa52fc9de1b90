"""Recycling pool of page-locked host buffers for the result arrays of the host entry points.

The reference returns a NEW ndarray per call (PyArray_NewLikeArray, _jbgrav.c:111,267,607,700) and
so does this binding.  A fresh pageable array makes the device-to-host copy a staged transfer into
memory whose pages are touched for the first time (measured on B200, tree N = 4M: 28.0 ms per call
around 4.4 ms of kernels; 8.5 ms with this pool in steady state).  Large results therefore live in page-locked blocks (gh_host_alloc):
the copy is one DMA transfer, and a block returns to the pool when the last array viewing it is
garbage collected, so steady-state calls neither allocate nor fault.

Small results (< 1 MiB), a library without a GPU, a pool that already holds MAX_BYTES, or
GH_PINNED_OUTPUT=0 fall back to ``np.empty``.
"""
import ctypes as C
import os
import threading
import weakref

import numpy as np

from . import _lib

MIN_BYTES = 1 << 20
MAX_BYTES = int(os.environ.get("GH_PINNED_MAX_BYTES", 8 << 30))    # outstanding + cached
CACHE_BYTES = int(os.environ.get("GH_PINNED_CACHE_BYTES", 2 << 30))  # idle blocks kept for reuse
ENABLED = os.environ.get("GH_PINNED_OUTPUT", "1") != "0"

_lock = threading.Lock()
_free = {}         # nbytes -> [ptr, ...]
_total = 0         # bytes in live + idle blocks
_cached = 0        # bytes in idle blocks
stats = {"allocated": 0, "reused": 0, "fallback": 0}


class _Block(object):
    """Owner of one page-locked block; numpy keeps it alive as the base of every view."""
    __slots__ = ("__array_interface__", "__weakref__")

    def __init__(self, ptr, shape):
        self.__array_interface__ = {"data": (ptr, False), "shape": tuple(shape), "typestr": "<f8", "version": 3}


def _raw_alloc(nbytes):
    p = C.c_void_p()
    if _lib.lib().gh_host_alloc(C.byref(p), nbytes) != _lib.GH_OK or not p.value:
        return None
    return p.value


def _raw_free(ptr):
    _lib.lib().gh_host_free(C.c_void_p(ptr))


def _release(ptr, nbytes):
    global _total, _cached
    with _lock:
        if _cached + nbytes <= CACHE_BYTES:
            _free.setdefault(nbytes, []).append(ptr)
            _cached += nbytes
            return
        _total -= nbytes
    try:
        _raw_free(ptr)
    except Exception:  # interpreter shutdown
        pass


def _acquire(nbytes):
    global _total, _cached
    with _lock:
        lst = _free.get(nbytes)
        if lst:
            _cached -= nbytes
            stats["reused"] += 1
            return lst.pop()
        if _total + nbytes > MAX_BYTES:
            return None
        _total += nbytes
    ptr = _raw_alloc(nbytes)
    if ptr is None:
        with _lock:
            _total -= nbytes
        return None
    stats["allocated"] += 1
    return ptr


def empty_f64(shape):
    """A new C-contiguous float64 array of ``shape``; page-locked when large and possible."""
    nbytes = 8 * int(np.prod(shape, dtype=np.int64))
    if not ENABLED or nbytes < MIN_BYTES:
        return np.empty(shape, dtype=np.float64)
    ptr = _acquire(nbytes)
    if ptr is None:
        stats["fallback"] += 1
        return np.empty(shape, dtype=np.float64)
    blk = _Block(ptr, shape)
    weakref.finalize(blk, _release, ptr, nbytes)
    return np.asarray(blk)


def trim():
    """Free the idle blocks (live arrays are untouched)."""
    global _total, _cached
    with _lock:
        items = [(n, p) for n, lst in _free.items() for p in lst]
        _free.clear()
        _total -= _cached
        _cached = 0
    for _, p in items:
        _raw_free(p)
