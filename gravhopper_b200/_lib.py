"""ctypes binding of libgravhopper_b200.so (include/gravhopper_b200.h).

The library is the product's only compute path.  If it is missing it is built with nvcc; if it
cannot be loaded, or no CUDA device is present when a compute call is made, the call raises --
there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GH_B200_LIB: load another build of the same ABI (kernel-variant experiments, scripts/gpu_variants.sh)
LIB_PATH = os.environ.get("GH_B200_LIB") or os.path.join(_HERE, "libgravhopper_b200.so")

GH_OK, GH_EINVAL, GH_ECUDA, GH_ENOMEM, GH_ESTATE = 0, 1, 2, 3, 4
GH_PREC_F32, GH_PREC_F64 = 32, 64
GH_MEM_HOST, GH_MEM_DEVICE = 0, 1
GH_ALG_DIRECT, GH_ALG_TREE = 0, 1

# every symbol include/gravhopper_b200.h declares: name -> (restype, argtypes)
_dp = C.POINTER(C.c_double)
_i64 = C.c_int64
_vp = C.c_void_p
_eng = C.c_void_p
PROTOTYPES = {
    "gh_last_error": (C.c_char_p, []),
    "gh_version": (C.c_int, []),
    "gh_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "gh_fp32_fma_probe": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "gh_direct_summation": (C.c_int, [C.c_int, _vp, _vp, _i64, C.c_double, _vp, C.c_int, _vp]),
    "gh_direct_summation_position": (C.c_int, [C.c_int, _vp, _vp, _i64, _vp, _i64, C.c_double, _vp,
                                               C.c_int, _vp]),
    "gh_tree_force": (C.c_int, [C.c_int, _vp, _vp, _i64, C.c_double, C.c_double, _vp, C.c_int, _vp]),
    "gh_tree_force_position": (C.c_int, [C.c_int, _vp, _vp, _i64, _vp, _i64, C.c_double, C.c_double,
                                         _vp, C.c_int, _vp]),
    "gh_release_thread_scratch": (C.c_int, []),
    "gh_host_alloc": (C.c_int, [C.POINTER(_vp), _i64]),
    "gh_host_free": (C.c_int, [_vp]),
    "gh_tree_last_stats": (C.c_int, [C.POINTER(_i64)]),
    "gh_set_tree_stats": (C.c_int, [C.c_int]),
    "gh_set_tree_walk": (C.c_int, [C.c_int]),
    "gh_get_tree_walk": (C.c_int, []),
    "gh_set_tree_quadrupoles": (C.c_int, [C.c_int]),
    "gh_get_tree_quadrupoles": (C.c_int, []),
    "gh_set_tree_walk_hybrid": (C.c_int, [C.c_double]),
    "gh_get_tree_walk_hybrid": (C.c_double, []),
    "gh_ic_sample": (C.c_int, [C.c_int, _i64, _dp, C.c_int, _vp, _vp, C.c_int, C.c_uint64, _vp, _vp, _vp,
                               C.c_int, _vp]),
    "gh_ic_sample_expdisk": (C.c_int, [_i64, _dp, _vp, _vp, _vp, _vp, C.c_int, C.c_uint64, _vp, _vp, _vp, C.c_int,
                                       _vp]),
    "gh_engine_upload_device": (C.c_int, [_eng, _vp, _vp, _vp]),
    "gh_engine_create": (C.c_int, [C.POINTER(_eng), C.c_int, _i64, _i64, _i64, C.c_int]),
    "gh_engine_destroy": (C.c_int, [_eng]),
    "gh_engine_upload": (C.c_int, [_eng, _vp, _vp, _vp]),
    "gh_engine_bind_sources": (C.c_int, [_eng, _vp, _vp]),
    "gh_engine_source_index": (C.c_int, [_eng, C.POINTER(C.c_int)]),
    "gh_engine_source_stride_bytes": (C.c_int, [_eng, C.POINTER(_i64)]),
    "gh_engine_set_origin": (C.c_int, [_eng, _dp]),
    "gh_engine_set_origin_velocity": (C.c_int, [_eng, _dp]),
    "gh_engine_prepare": (C.c_int, [_eng, C.c_double]),
    "gh_engine_add_potential": (C.c_int, [_eng, C.c_int, _dp, C.c_int]),
    "gh_engine_clear_potentials": (C.c_int, [_eng]),
    "gh_engine_step": (C.c_int, [_eng, C.c_double, C.c_double, C.c_double, C.c_int, _vp, C.c_int]),
    "gh_engine_run": (C.c_int, [_eng, _i64, C.c_double, C.c_double, C.c_double, C.c_int, _i64, _vp,
                                _vp]),
    "gh_engine_download": (C.c_int, [_eng, _vp, _vp]),
    "gh_engine_download_xhalf": (C.c_int, [_eng, _vp]),
    "gh_engine_set_dt": (C.c_int, [_eng, C.c_double]),
    "gh_engine_energy": (C.c_int, [_eng, C.c_double, _dp]),
    "gh_engine_synchronize": (C.c_int, [_eng]),
    "gh_engine_state_ptrs": (C.c_int, [_eng, C.POINTER(_vp), C.POINTER(_vp)]),
    "gh_engine_stream": (C.c_int, [_eng, C.POINTER(_vp)]),
    "gh_engine_tree_stats": (C.c_int, [_eng, C.POINTER(_i64)]),
    "gh_engine_launch_count": (C.c_int, [_eng, C.POINTER(_i64)]),
    "gh_engine_history_pinned": (C.c_int, [_eng, C.POINTER(C.c_int)]),
    "gh_engine_last_force_ms": (C.c_int, [_eng, C.POINTER(C.c_float)]),
    "gh_engine_force_ms_mean": (C.c_int, [_eng, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
}

_grp = C.c_void_p
PROTOTYPES.update({
    "gh_nccl_version": (C.c_int, [C.POINTER(C.c_int)]),
    "gh_group_create_local": (C.c_int, [C.POINTER(_grp), C.c_int, C.POINTER(C.c_int), _i64, C.c_int]),
    "gh_group_unique_id": (C.c_int, [_vp]),
    "gh_group_create_rank": (C.c_int, [C.POINTER(_grp), _vp, C.c_int, C.c_int, C.c_int, _i64, C.c_int]),
    "gh_group_destroy": (C.c_int, [_grp]),
    "gh_group_size": (C.c_int, [_grp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gh_group_engine": (C.c_int, [_grp, C.c_int, C.POINTER(_eng), C.POINTER(_i64), C.POINTER(_i64)]),
    "gh_group_prepare": (C.c_int, [_grp, C.c_double]),
    "gh_group_step": (C.c_int, [_grp, _i64, C.c_double, C.c_double, C.c_double, C.c_int]),
    "gh_group_synchronize": (C.c_int, [_grp]),
    "gh_group_set_tree_distributed": (C.c_int, [_grp, C.c_int]),
    "gh_group_phase_ms": (C.c_int, [_grp, C.POINTER(C.c_float)]),
})

_lib = None


class GravHopperB200Error(RuntimeError):
    """A libgravhopper_b200 call failed (message from gh_last_error())."""


def lib():
    """Load (building first if necessary) the C-ABI library.  Raises if that is impossible."""
    global _lib
    if _lib is None:
        if "GH_B200_LIB" not in os.environ:
            # (re)build when the library is missing or older than any of its sources; a box without
            # nvcc (or a read-only tree) keeps a library that exists -- and says so if it is stale
            from . import build as _build
            if not os.path.exists(LIB_PATH):
                _build.build()
            elif _build.stale():
                try:
                    _build.build()
                except (RuntimeError, OSError) as exc:
                    import warnings
                    warnings.warn("libgravhopper_b200.so is older than its sources and could not be rebuilt: %s"
                                  % (str(exc).splitlines() or [""])[0])
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what="libgravhopper_b200"):
    if rc == GH_OK:
        return
    msg = lib().gh_last_error().decode("utf-8", "replace")
    if rc == GH_ENOMEM:
        raise MemoryError("%s: %s" % (what, msg))
    if rc == GH_EINVAL:
        raise ValueError("%s: %s" % (what, msg))
    raise GravHopperB200Error("%s failed (code %d): %s" % (what, rc, msg))


def device_count():
    n = C.c_int(0)
    lib().gh_device_count(C.byref(n))
    return n.value


def require_gpu():
    if device_count() <= 0:
        raise GravHopperB200Error(
            "gravhopper_b200 needs a CUDA device (B200, sm_100a); none is visible and there is "
            "no CPU fallback")
