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


def _event_rows(o):
    """rlgs_row4e words of an oracle run: idle_nodes:12 | started:1 | queued[18:0]:19 (include/rlgs.h)"""
    import numpy as np
    started = np.zeros(o['n_ticks'], np.uint32)
    started[o['start'][o['start'] >= 0]] = 1
    return (o['rows']['idle_nodes'].astype(np.uint32) | (started << 12) | ((o['rows']['queued'].astype(np.uint32) & 0x7ffff) << 13)).astype('<u4')


@pytest.mark.parametrize('name', ['probe100', 'ties', 'multi_node', 'big_mem_leak', 'dense', 'kat6', 'nondivisible'])
def test_host_only_replay_of_event_rows(name):
    """rlgs_replay_rows4e (no device): the 4-byte event rows of a run + its trace give back the job tables and the pending-time
    columns.  Rows are built from the oracle's run of a reference fixture; outputs are compared with the oracle's tables, with
    its per-tick statistics and (through the product's formatter) with the reference's job.csv."""
    import numpy as np
    import cpu_sim
    import goldutil
    from rlgpuschedule_b200 import log_manager as lm
    g = goldutil.load(name)
    ti = goldutil.trace_input(g)
    cluster = rl.cluster_from_flags(g['flags'])
    tr = rl.prepare_trace(ti, cluster)
    o = cpu_sim.run_fifo_yarn(cpu_sim.make_cluster(**g['flags']), cpu_sim.prepare_trace(ti))
    w = _event_rows(o)
    J, n = len(tr.records), len(w)
    st, en, fo = (np.empty(J, np.int32) for _ in range(3))
    pend = np.empty(3 * n, np.int32)
    nf = C.c_int32(-1)
    rec = np.ascontiguousarray(tr.records)
    L = _ffi.lib()
    _ffi.check(L.rlgs_replay_rows4e(rec.ctypes.data, J, w.ctypes.data, n, st.ctypes.data, en.ctypes.data, fo.ctypes.data, C.byref(nf), pend.ctypes.data))
    assert nf.value == len(o['finish_order'])
    assert np.array_equal(st, o['start']) and np.array_equal(en, o['end']) and np.array_equal(fo[:nf.value], o['finish_order']) and (fo[nf.value:] == -1).all()
    pend = pend.reshape(n, 3)
    q = o['rows']['queued'] > 0
    assert np.array_equal(pend[q, 0].astype(np.float64), o['rows']['max_pending'][q]) and (pend[~q] == 0).all()
    assert np.array_equal((pend[q, 1].astype(np.float64) + pend[q, 2]) / 2.0, o['rows']['median_pending'][q])
    assert lm.format_job_csv(tr, fo[:nf.value], st, en) == g['job']
    r = rl.replay_event_rows(tr, w)                               # the Python wrapper of the same call
    assert np.array_equal(r['start'], st) and np.array_equal(r['finish_order'], fo[:nf.value]) and np.array_equal(r['median_hi'], pend[:, 2])
    # outputs are optional; rows that do not belong to the trace are refused
    _ffi.check(L.rlgs_replay_rows4e(rec.ctypes.data, J, w.ctypes.data, n, None, None, None, None, None))
    bad = w.copy()
    first_start = int(np.nonzero(bad & 0x1000)[0][0])
    bad[first_start] &= ~np.uint32(0x1000)                      # drop one start event: the queue length check must trip
    assert L.rlgs_replay_rows4e(rec.ctypes.data, J, bad.ctypes.data, n, None, None, None, None, None) == _ffi.ERR_STATE
