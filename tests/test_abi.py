"""CPU-side checks of the drop-in boundary: the library builds, loads without a GPU, exports every
symbol include/rlgs.h declares, its plain structs have the documented sizes, and it fails loudly
(no CPU fallback) when no device is present."""
import ctypes as C
import os
import re

import pytest

import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import _ffi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    so = build.build()
    lib = C.CDLL(so)
    hdr = open(os.path.join(ROOT, 'include', 'rlgs.h')).read()
    declared = set(re.findall(r'\b(rlgs_[a-z_0-9]+)\s*\(', hdr))
    assert declared, 'no declarations found'
    assert declared == set(_ffi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert _ffi.lib().rlgs_version() == 200


def test_struct_sizes_match_header():
    assert _ffi.JOB_DTYPE.itemsize == 32
    assert _ffi.ROW_DTYPE.itemsize == 64
    assert C.sizeof(_ffi.ClusterSpec) == 24
    assert C.sizeof(_ffi.Summary) == 6 * 8 + 8 * 4
    assert C.sizeof(_ffi.Opts) == 12 * 4 + 8 * 4 + 2 * 8 + 2 * 8 + 2 * 4 + 2 * 4
    assert _ffi.ROW16_DTYPE.itemsize == 16 and _ffi.ROW12_DTYPE.itemsize == 12
    assert C.sizeof(_ffi.PackInputs) == 4 * 8 + 2 * 4 + 3 * 8


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_ffi.RlgsError) as e:
        rl.Simulator(rl.Cluster())
    assert e.value.code == _ffi.ERR_CUDA


def test_bad_arguments_are_rejected_before_touching_the_device():
    L = _ffi.lib()
    spec = rl.Cluster(num_gpu_p_node=64).to_ffi()
    o = _ffi.Opts(); o.n_replicas = 1
    h = C.c_void_p()
    assert L.rlgs_create(C.byref(spec), C.byref(o), C.byref(h)) == _ffi.ERR_BAD_ARG
    assert b'num_gpu_p_node' in L.rlgs_last_error()
    with pytest.raises(NotImplementedError):
        rl.Simulator(rl.Cluster(), schedule='lpjf')
