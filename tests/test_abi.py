"""CPU tests: the C-ABI library loads here (no GPU) and exports every symbol the header declares."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from gravhopper_b200 import _lib, _jbgrav

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "gravhopper_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gh_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = header_functions()
    assert len(names) >= 25
    lib = _lib.lib()
    for n in names:
        assert hasattr(lib, n), "libgravhopper_b200.so does not export %s" % n
        assert n in _lib.PROTOTYPES, "ctypes binding lacks %s" % n
    for n in _lib.PROTOTYPES:
        assert n in names, "binding %s is not declared in include/gravhopper_b200.h" % n


def test_no_torch_or_python_in_abi():
    out = subprocess.run(["nm", "-D", "--undefined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "Py_" not in out and "torch" not in out and "c10" not in out


def test_built_for_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_version_and_error_string():
    lib = _lib.lib()
    assert lib.gh_version() >= 100
    assert isinstance(lib.gh_last_error(), bytes)


def test_argument_validation_without_gpu():
    lib = _lib.lib()
    # invalid arguments are rejected before any CUDA call
    assert lib.gh_direct_summation(16, None, None, 4, 0.1, None, 0, None) == _lib.GH_EINVAL
    assert lib.gh_direct_summation(64, None, None, 4, 0.1, None, 0, None) == _lib.GH_EINVAL
    assert lib.gh_tree_force(64, None, None, 4, 0.1, -1.0, None, 0, None) == _lib.GH_EINVAL
    h = ctypes.c_void_p()
    assert lib.gh_engine_create(ctypes.byref(h), 0, 10, 5, 6, 64) == _lib.GH_EINVAL
    assert lib.gh_set_tree_walk_hybrid(-0.5) == _lib.GH_EINVAL and lib.gh_set_tree_walk_hybrid(2.0) == _lib.GH_EINVAL
    if not os.environ.get("GH_WALK_HYBRID"):
        assert abs(lib.gh_get_tree_walk_hybrid() - 0.10) < 1e-6   # the hybrid rule is on by default
    prm = (ctypes.c_double * 4)(1.0, 2.0, 0.2, 20.0)
    assert lib.gh_ic_sample_expdisk(10, prm, None, None, None, None, 0, 1, None, None, None, 0, None) == _lib.GH_EINVAL
    p = ctypes.c_void_p()
    assert lib.gh_host_alloc(ctypes.byref(p), 0) == _lib.GH_EINVAL and lib.gh_host_free(None) == _lib.GH_OK


def test_group_argument_validation_without_gpu():
    """Engine groups (multi-GPU): bad arguments are rejected before any CUDA or NCCL call; NCCL itself
    is found at run time (dlopen), never linked."""
    lib = _lib.lib()
    g = ctypes.c_void_p()
    assert lib.gh_group_create_local(ctypes.byref(g), 0, None, 100, 32) == _lib.GH_EINVAL      # no device
    assert lib.gh_group_create_local(ctypes.byref(g), 4, None, 3, 32) == _lib.GH_EINVAL        # fewer particles than ranks
    assert lib.gh_group_create_rank(ctypes.byref(g), None, 5, 4, 0, 100, 32) == _lib.GH_EINVAL  # rank >= world
    assert lib.gh_group_create_rank(ctypes.byref(g), None, 0, 2, 0, 100, 32) == _lib.GH_EINVAL  # world > 1 needs an id
    assert lib.gh_group_unique_id(None) == _lib.GH_EINVAL
    assert lib.gh_group_step(None, 1, 0.1, 0.1, 0.7, 0) == _lib.GH_EINVAL
    assert lib.gh_group_destroy(None) == _lib.GH_OK
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "nccl" not in out
    v = ctypes.c_int()
    if lib.gh_nccl_version(ctypes.byref(v)) == _lib.GH_OK:   # a libnccl.so.2 is on this machine
        assert v.value >= 20000
    # the opt-in quadrupole switch: off unless asked for
    if not os.environ.get("GH_TREE_QUADRUPOLES"):
        assert lib.gh_get_tree_quadrupoles() == 0
    assert lib.gh_set_tree_quadrupoles(1) == _lib.GH_OK and lib.gh_get_tree_quadrupoles() == 1
    assert lib.gh_set_tree_quadrupoles(0) == _lib.GH_OK and lib.gh_get_tree_quadrupoles() == 0


def test_reference_error_messages():
    """Shape errors: same type and text as _jbgrav.c:92,98,106,228,235,244,253,260."""
    with pytest.raises(RuntimeError, match="Position array is not Nx3."):
        _jbgrav.direct_summation(np.zeros((4, 2)), np.ones(4), 0.1)
    with pytest.raises(RuntimeError, match="Position array does not have 2 dimensions."):
        _jbgrav.tree_force(np.zeros(4), np.ones(4), 0.1, 0.7)
    with pytest.raises(RuntimeError, match="Mass array and position array contain different numbers of particles."):
        _jbgrav.direct_summation(np.zeros((4, 3)), np.ones(5), 0.1)
    with pytest.raises(RuntimeError, match="Particle position array is not Nx3."):
        _jbgrav.direct_summation_position(np.zeros((4, 2)), np.ones(4), np.zeros((2, 3)), 0.1)
    with pytest.raises(RuntimeError, match="Force position array is not Nx3."):
        _jbgrav.tree_force_position(np.zeros((4, 3)), np.ones(4), np.zeros((2, 2)), 0.1, 0.7)
    with pytest.raises(RuntimeError, match="Mass array and particle position array contain different"):
        _jbgrav.tree_force_position(np.zeros((4, 3)), np.ones(3), np.zeros((2, 3)), 0.1, 0.7)
    with pytest.raises(TypeError):
        _jbgrav.tree_force(np.zeros((4, 3)), np.ones(4), 0.1)  # theta is required at this level


def test_compute_fails_loudly_without_gpu():
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.GravHopperB200Error, match="no CPU fallback"):
        _jbgrav.direct_summation(np.zeros((4, 3)), np.ones(4), 0.1)


def test_product_never_imports_oracle():
    """The product path must not route through the checker."""
    pkg = os.path.join(ROOT, "gravhopper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), os.path.join(dirpath, f)
