"""ctypes binding of librlgs.so (include/rlgs.h).  Fails loudly: there is no CPU fallback.

The library is built in-tree by rlgpuschedule_b200/build.py (nvcc, sm_100a) and loaded from the
package directory so the driver can see which .so a test / bench process used.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get('RLGS_LIB') or os.path.join(HERE, 'librlgs.so')   # RLGS_LIB: development override

OK, ERR_BAD_ARG, ERR_CUDA, ERR_OOM, ERR_UNSUPPORTED, ERR_CAPACITY, ERR_STATE, ERR_SLOTS, ERR_WIRE = 0, -1, -2, -3, -4, -5, -6, -7, -8
ROWFMT_WIDE, ROWFMT_WIRE16, ROWFMT_WIRE12, ROWFMT_EVENT16, ROWFMT_EVENT4 = 0, 1, 2, 3, 4
ROWFMT = {'wide': 0, 'wire16': 1, 'wire12': 2, 'event16': 3, 'event4': 4}
SCHED = {'fifo': 0, 'sjf': 1, 'dlas-gpu': 2, 'dlas': 3, 'shortest': 4, 'shortest-gpu': 5, 'horus': 6, 'gandiva': 7, 'horus+': 8}
PLACE = {'yarn': 0, 'count': 1, 'horus': 2, 'horus+': 2, 'gandiva': 2}   # the three pack names share horus_placement (algorithm.py:182-187)
ROWS_NONE, ROWS_FULL, ROWS_DEVICE = 0, 1, 2
MAX_QUEUES = 8
ROWS_PER_CHUNK = 4096
PLANE_START, PLANE_END, PLANE_FINISH_ORDER, PLANE_AUX, PLANE_PREEMPT, PLANE_RESUME = range(6)


class ClusterSpec(C.Structure):
    _fields_ = [('num_switch', C.c_int32), ('num_node_p_switch', C.c_int32), ('num_gpu_p_node', C.c_int32),
                ('num_cpu_p_node', C.c_int32), ('mem_p_node', C.c_int32), ('reserved', C.c_int32)]


class Opts(C.Structure):
    _fields_ = [('device', C.c_int32), ('n_replicas', C.c_int32), ('schedule', C.c_int32), ('placement', C.c_int32),
                ('rows_mode', C.c_int32), ('slot_cap', C.c_int32), ('n_streams', C.c_int32),
                ('ticks_per_launch', C.c_int32), ('num_queue', C.c_int32), ('enable_network_costs', C.c_int32),
                ('fetch_jobs', C.c_int32), ('num_buffer', C.c_int32), ('queue_limit', C.c_int32 * MAX_QUEUES),
                ('bandwidth', C.c_double), ('internode_latency', C.c_double), ('max_ticks', C.c_int64),
                ('rows_cap', C.c_int64), ('pack_rng', C.c_int32), ('pack_seed', C.c_uint32),
                ('rows_format', C.c_int32), ('lanes_per_replica', C.c_int32)]


class PackInputs(C.Structure):
    _fields_ = [('util_avg', C.POINTER(C.c_double)), ('util_sd', C.POINTER(C.c_double)), ('task_mem', C.POINTER(C.c_int64)),
                ('heap_cap', C.POINTER(C.c_int32)), ('mem_shift', C.c_int32), ('gpu_mem_cap_mib', C.c_int32),
                ('util_max', C.POINTER(C.c_double)), ('mem_avg_mib', C.POINTER(C.c_double)), ('used_gpus', C.POINTER(C.c_double))]


class NetcostInputs(C.Structure):
    _fields_ = [('duration', C.POINTER(C.c_double)), ('model_mb', C.POINTER(C.c_double)),
                ('iterations', C.POINTER(C.c_double))]


class Summary(C.Structure):
    _fields_ = [('n_ticks', C.c_int64), ('makespan', C.c_int64), ('sum_jct', C.c_int64), ('sum_queued', C.c_int64),
                ('sum_running', C.c_int64), ('events', C.c_int64), ('n_jobs', C.c_int32), ('n_arrived', C.c_int32),
                ('n_started', C.c_int32), ('n_finished', C.c_int32), ('max_queued', C.c_int32), ('max_running', C.c_int32),
                ('status', C.c_int32), ('done', C.c_int32)]


# numpy views of the plain structs of rlgs.h
JOB_DTYPE = np.dtype([('arrival_tick', '<i4'), ('dur_ticks', '<i4'), ('gpus', '<u2'), ('tasks', '<u2'),
                      ('gpus_per_task', '<u2'), ('least_nodes_fits', '<u2'), ('mem_term', '<i8'),
                      ('util_mu_q', '<u2'), ('util_sd_q', '<u2'), ('index', '<i4')])
ROW_DTYPE = np.dtype([('idle_nodes', '<i4'), ('busy_gpus', '<i4'), ('running', '<i4'), ('queued', '<i4'),
                      ('finished', '<i4'), ('median_lo', '<i4'), ('median_hi', '<i4'), ('max_pending', '<i4'),
                      ('sum_pending', '<i8'), ('mem_sum', '<i8'), ('util_mu_sum', '<i8'), ('util_var_sum', '<i8')])
ROW16_DTYPE = np.dtype([('w', '<u4', (4,))])
ROW12_DTYPE = np.dtype([('w', '<u4', (3,))])
ROW4_DTYPE = np.dtype([('w', '<u4')])
assert JOB_DTYPE.itemsize == 32 and ROW_DTYPE.itemsize == 64 and ROW16_DTYPE.itemsize == 16 and ROW12_DTYPE.itemsize == 12 and ROW4_DTYPE.itemsize == 4

EXPORTS = ['rlgs_version', 'rlgs_last_error', 'rlgs_create', 'rlgs_destroy', 'rlgs_load_trace', 'rlgs_load_pack_inputs', 'rlgs_run',
           'rlgs_last_run_ms', 'rlgs_set_stream', 'rlgs_get_summary', 'rlgs_read_jobs', 'rlgs_read_rows',
           'rlgs_rows_view', 'rlgs_read_rows16', 'rlgs_rows16_view', 'rlgs_read_rows12', 'rlgs_rows12_view', 'rlgs_read_rows16e', 'rlgs_rows16e_view', 'rlgs_read_rows4e', 'rlgs_rows4e_view', 'rlgs_replay_rows4e', 'rlgs_read_job_plane', 'rlgs_returns', 'rlgs_returns_device_ptr',
           'rlgs_read_durations', 'rlgs_env_obs_dim', 'rlgs_env_reset', 'rlgs_env_step', 'rlgs_env_observe', 'rlgs_env_sync']

_lib = None


class RlgsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('librlgs error %d: %s' % (code, msg))
        self.code = code


def lib():
    """Loads librlgs.so; raises if it has not been built (python -m rlgpuschedule_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError('%s is missing: build it with `python -m rlgpuschedule_b200.build` '
                           '(nvcc, sm_100a). There is no CPU fallback.' % SO_PATH)
    L = C.CDLL(SO_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.rlgs_version.restype = i32
    L.rlgs_last_error.restype = C.c_char_p
    L.rlgs_create.argtypes = [C.POINTER(ClusterSpec), C.POINTER(Opts), C.POINTER(vp)]
    L.rlgs_destroy.argtypes = [vp]
    L.rlgs_destroy.restype = None
    L.rlgs_load_trace.argtypes = [vp, i32, i32, vp, i32, C.POINTER(NetcostInputs)]
    L.rlgs_load_pack_inputs.argtypes = [vp, i32, i32, C.POINTER(PackInputs), i32]
    L.rlgs_run.argtypes = [vp]
    L.rlgs_last_run_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
    L.rlgs_set_stream.argtypes = [vp, vp, i32]
    L.rlgs_get_summary.argtypes = [vp, i32, C.POINTER(Summary)]
    L.rlgs_read_jobs.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.rlgs_read_rows.argtypes = [vp, i32, i64, i64, vp]
    L.rlgs_rows_view.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(i64)]
    L.rlgs_read_rows16.argtypes = [vp, i32, i64, i64, vp]
    L.rlgs_rows16_view.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(i64)]
    L.rlgs_read_rows12.argtypes = [vp, i32, i64, i64, vp]
    L.rlgs_rows12_view.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(i64)]
    L.rlgs_read_rows16e.argtypes = [vp, i32, i64, i64, vp]
    L.rlgs_rows16e_view.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(i64)]
    L.rlgs_read_rows4e.argtypes = [vp, i32, i64, i64, vp]
    L.rlgs_rows4e_view.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(i64)]
    L.rlgs_replay_rows4e.argtypes = [vp, i32, vp, i64, vp, vp, vp, C.POINTER(i32), vp]
    L.rlgs_read_job_plane.argtypes = [vp, i32, i32, vp]
    L.rlgs_returns.argtypes = [vp, vp]
    L.rlgs_returns_device_ptr.argtypes = [vp, C.POINTER(vp)]
    L.rlgs_read_durations.argtypes = [vp, i32, vp]
    L.rlgs_env_obs_dim.argtypes = [vp, i32, C.POINTER(i32)]
    L.rlgs_env_reset.argtypes = [vp]
    L.rlgs_env_step.argtypes = [vp, vp, vp, vp, vp, i32, i32, C.c_uint32, i32]
    L.rlgs_env_observe.argtypes = [vp, vp, vp, vp, i32]
    L.rlgs_env_sync.argtypes = [vp]
    for name in EXPORTS:
        if name not in ('rlgs_last_error', 'rlgs_destroy'):
            getattr(L, name).restype = i32
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise RlgsError(rc, lib().rlgs_last_error().decode(errors='replace'))
