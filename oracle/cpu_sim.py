"""ctypes front-end of the CPU oracle (oracle/cpu_sim.c, oracle/cpu_legacy.c).  TEST INFRASTRUCTURE:
imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.

`prepare_trace` restates JobTraceReader.prepare_jobs (/root/reference/core/jobs/job_generator.py:181-196)
and the Job constructor's derived fields (/root/reference/core/jobs/jobs_manager.py:233-238);
`format_job_csv` / `format_cluster_csv` restate LogManager.jcts / step_cluster
(/root/reference/log_manager.py:118-155) with the csv module, as the reference does.
"""
import csv
import ctypes as C
import io
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Cluster(C.Structure):
    _fields_ = [('num_switch', C.c_int32), ('num_node_p_switch', C.c_int32), ('num_gpu_p_node', C.c_int32),
                ('num_cpu_p_node', C.c_int32), ('mem_p_node', C.c_int32), ('gpu_memory_capacity_mib', C.c_int32)]


ROW_DTYPE = np.dtype([('idle_nodes', 'i4'), ('busy_nodes', 'i4'), ('busy_gpus', 'i4'), ('idle_gpus', 'i4'),
                      ('running', 'i4'), ('queued', 'i4'), ('finished', 'i4'), ('max_is_int_zero', 'i4'),
                      ('avg_gpu_memory_allocated', 'f8'), ('avg_pending', 'f8'), ('median_pending', 'f8'),
                      ('max_pending', 'f8'), ('util_mu_sum', 'f8')])


def build(force=False):
    so = os.path.join(HERE, 'liboracle.so')
    srcs = [os.path.join(HERE, f) for f in ('cpu_sim.c', 'cpu_legacy.c') if os.path.exists(os.path.join(HERE, f))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['make', '-C', HERE, '-s', '-B', 'liboracle.so'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def make_cluster(num_switch=1, num_node_p_switch=32, num_gpu_p_node=8, num_cpu_p_node=128, mem_p_node=512,
                 gpu_memory_capacity=32, cluster_spec=None, **_):
    """Defaults = /root/reference/run_sim.py:50-82; cluster_spec = infra/infrastructure.py:78-105."""
    if cluster_spec and os.path.exists(cluster_spec):
        with open(cluster_spec) as f:
            rd = csv.DictReader(f)
            need = ['num_switch', 'num_node_p_switch', 'num_gpu_p_node', 'num_cpu_p_node', 'mem_p_node']
            if all(k in rd.fieldnames for k in need):
                for row in rd:
                    num_switch, num_node_p_switch, num_gpu_p_node, num_cpu_p_node, mem_p_node = (int(row[k]) for k in need)
    return Cluster(num_switch, num_node_p_switch, num_gpu_p_node, num_cpu_p_node, mem_p_node, gpu_memory_capacity * 1024)


def prepare_trace(trace, scale_factor=0.5):
    """trace: CSV path or DataFrame -> dict of per-job arrays in queue-arrival order."""
    import pandas as pd
    df = pd.read_csv(trace) if isinstance(trace, (str, os.PathLike)) else trace.copy()
    df = df[df['type'] == 'noninteractive']
    df = df.sort_values(by='normalized_time')
    df = df.dropna()
    nt = df['normalized_time'] - df['normalized_time'].min()
    nt = nt / 10000
    return dict(label=df.index.to_numpy().astype(np.int64),
                nt=np.ascontiguousarray(nt.to_numpy(dtype=np.float64)),
                duration=np.ascontiguousarray(df['minutes'].to_numpy(dtype=np.float64) * scale_factor),
                used_gpus=np.ascontiguousarray(df['used_gpus'].to_numpy(dtype=np.float64)),
                gpc=np.ascontiguousarray(df['gpu_per_container'].to_numpy(dtype=np.int32)),
                mem_mib=np.ascontiguousarray(df['memory_max'].to_numpy(dtype=np.float64) / 1024 / 1024),
                util_avg=np.ascontiguousarray(df['gpu_utilization_avg'].to_numpy(dtype=np.float64)),
                util_max=np.ascontiguousarray(df['gpu_utilization_max'].to_numpy(dtype=np.float64)),
                mem_avg_mib=np.ascontiguousarray(df['memory_avg'].to_numpy(dtype=np.float64) / 1024 / 1024))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def run_fifo_yarn(cluster, tr, rows_cap=None, netcost=None):
    """Returns dict(finish_order, start, end, rows (ROW_DTYPE), n_ticks, counters).
    netcost = dict(model_mb, iterations, bandwidth, latency) enables the build-defined network-cost model."""
    L = lib()
    n = len(tr['nt'])
    dur_out = None
    L.oracle_set_netcost.restype = None
    if netcost:
        dur_out = np.zeros(max(n, 1), np.float64)
        nm = np.ascontiguousarray(netcost['model_mb'], np.float64); ni = np.ascontiguousarray(netcost['iterations'], np.float64)
        L.oracle_set_netcost(_p(nm, C.c_double), _p(ni, C.c_double), C.c_double(netcost['bandwidth']), C.c_double(netcost['latency']), _p(dur_out, C.c_double))
    else:
        L.oracle_set_netcost(None, None, C.c_double(0), C.c_double(0), None)
    fin = np.empty(max(n, 1), np.int32); st = np.empty(max(n, 1), np.int32); en = np.empty(max(n, 1), np.int32)
    nfin = C.c_int32(0); nticks = C.c_int64(0); counters = np.zeros(4, np.int64)
    cap = rows_cap or max(4096, 4 * n)
    while True:
        rows = np.zeros(cap, ROW_DTYPE)
        rc = L.oracle_fifo_yarn(C.byref(cluster), C.c_int32(n), _p(tr['nt'], C.c_double), _p(tr['duration'], C.c_double),
                                _p(tr['used_gpus'], C.c_double), _p(tr['gpc'], C.c_int32), _p(tr['mem_mib'], C.c_double),
                                _p(tr['util_avg'], C.c_double), _p(fin, C.c_int32), _p(st, C.c_int32), _p(en, C.c_int32),
                                C.byref(nfin), rows.ctypes.data_as(C.c_void_p), C.c_int64(cap), C.byref(nticks),
                                _p(counters, C.c_int64))
        if rc == -1:
            cap *= 4
            continue
        if rc != 0:
            raise RuntimeError('oracle_fifo_yarn rc=%d (the reference would raise on this input)' % rc)
        break
    k = nfin.value
    L.oracle_set_netcost(None, None, C.c_double(0), C.c_double(0), None)
    out = dict(finish_order=fin[:k].copy(), start=st[:n], end=en[:n], rows=rows[:nticks.value], n_ticks=nticks.value,
               counters=dict(sum_queued=int(counters[0]), sum_running=int(counters[1]), ticks=int(counters[2]),
                             starts=int(counters[3])))
    if dur_out is not None:
        out['actual_duration'] = dur_out[:n]
    return out


def run_pack(cluster, tr, schedule='horus', num_buffer=5, seed=None, replica=0, rows_cap=None, scheme=None, num_queue=0, inject_seed=0):
    """Restated `--schedule horus|gandiva` with horus_placement (oracle_pack).  seed=None pins every utilisation draw
    to its mean (the reference's behaviour on traces with gpu_utilization_max == gpu_utilization_avg: PINNED); a seed
    enables the build-defined counter-based draw (UNPINNED)."""
    L = lib()
    n = len(tr['nt'])
    fin = np.empty(max(n, 1), np.int32); st = np.empty(max(n, 1), np.int32); en = np.empty(max(n, 1), np.int32)
    dur = np.zeros(max(n, 1), np.float64)
    jct = np.zeros(max(n, 1), np.int32); starts = np.zeros(max(n, 1), np.int32)
    nfin = C.c_int32(0); nticks = C.c_int64(0); counters = np.zeros(4, np.int64)
    cap = rows_cap or max(4096, 4 * n)
    while True:
        rows = np.zeros(cap, ROW_DTYPE)
        rc = L.oracle_pack(C.byref(cluster), C.c_int32(n), _p(tr['nt'], C.c_double), _p(tr['duration'], C.c_double),
                           _p(tr['used_gpus'], C.c_double), _p(tr['gpc'], C.c_int32), _p(tr['mem_mib'], C.c_double),
                           _p(tr['util_avg'], C.c_double), _p(tr['util_max'], C.c_double),
                           C.c_int32(1 if schedule == 'gandiva' else 0), C.c_int32(1 if scheme == 'yarn' else 0), C.c_int32(num_buffer),
                           C.c_int32(0 if seed is None else 1), C.c_uint32(seed or 0), C.c_uint32(replica),
                           _p(fin, C.c_int32), _p(st, C.c_int32), _p(en, C.c_int32), C.byref(nfin), _p(dur, C.c_double),
                           _p(jct, C.c_int32), _p(starts, C.c_int32),
                           C.c_int32(num_queue if schedule == 'horus+' else 0), _p(tr['mem_avg_mib'], C.c_double), C.c_uint32(inject_seed), rows.ctypes.data_as(C.c_void_p), C.c_int64(cap), C.byref(nticks), _p(counters, C.c_int64))
        if rc == -1:
            cap *= 4
            continue
        if rc != 0:
            raise RuntimeError('oracle_pack rc=%d (the reference would raise on this input)' % rc)
        break
    k = nfin.value
    return dict(finish_order=fin[:k].copy(), start=st[:n], end=en[:n], rows=rows[:nticks.value], n_ticks=nticks.value,
                actual_duration=dur[:n], orig_duration=tr['duration'], jct=jct[:n], preempt=starts[:n],
                counters=dict(sum_queued=int(counters[0]), sum_running=int(counters[1]), ticks=int(counters[2]),
                              starts=int(counters[3])))


def format_job_csv(tr, res):
    """job.csv as LogManager.jcts writes it (/root/reference/log_manager.py:45-54,137-155)."""
    buf = io.StringIO(newline='')
    w = csv.writer(buf)
    w.writerow(['job_id', 'num_gpu', 'submit_time', 'start_time', 'end_time', 'original_duration',
                'actual_duration', 'jct', 'preempt'])
    dur = res.get('actual_duration', tr['duration'])
    pre = res.get('preempt')
    for k, i in enumerate(res['finish_order']):
        i = int(i)
        w.writerow([str(int(tr['label'][i])), float(tr['used_gpus'][i]), int(tr['nt'][i]), int(res['start'][i]),
                    int(res['end'][i]), float(res['orig_duration'][i] if 'orig_duration' in res else dur[i]),   # Job.duration itself carries the network cost (job.py:196-197)
                    float(dur[i]) if dur[i] > 0 else 0,   # Job.get_duration: max(0, d) keeps the int 0 (job.py:206-210)
                    int(res['jct'][i]) if 'jct' in res else int(res['end'][i] - res['start'][i]),
                    int(pre[i]) if pre is not None else 1])
    return buf.getvalue()


def format_cluster_csv(res, with_util=False):
    """cluster.csv as LogManager.step_cluster writes it (log_manager.py:37-44,118-135); by default
    WITHOUT the avg_gpu_utilization column (unseeded RNG in the reference)."""
    hdr = ['delta', 'num_idle_nodes', 'num_busy_nodes', 'num_busy_gpus', 'num_idle_gpus', 'avg_gpu_utilization',
           'avg_gpu_memory_allocated', 'avg_pending_time', 'median_pending_time', 'max_pending_time',
           'num_running_jobs', 'num_queuing_jobs', 'num_finish_jobs']
    if not with_util:
        hdr.remove('avg_gpu_utilization')
    buf = io.StringIO(newline='')
    w = csv.writer(buf)
    w.writerow(hdr)
    rows = res['rows']
    for d in range(len(rows)):
        r = rows[d]
        row = [d + 1, int(r['idle_nodes']), int(r['busy_nodes']), int(r['busy_gpus']), int(r['idle_gpus'])]
        if with_util:
            row.append(0.0)
        row += [float(r['avg_gpu_memory_allocated']), float(r['avg_pending']), float(r['median_pending']),
                0 if r['max_is_int_zero'] else float(r['max_pending']),
                int(r['running']), int(r['queued']), int(r['finished'])]
        w.writerow(row)
    return buf.getvalue()


def run_fifo_yarn_batch(cluster, tr, n_runs, n_threads, rows_cap=None):
    """n_runs replica-runs of one trace on n_threads pthreads (CPU baseline). Returns total events."""
    L = lib()
    L.oracle_fifo_yarn_batch.restype = C.c_int64
    n = len(tr['nt'])
    cap = rows_cap or max(4096, 4 * n)
    ev = L.oracle_fifo_yarn_batch(C.byref(cluster), C.c_int32(n), _p(tr['nt'], C.c_double), _p(tr['duration'], C.c_double),
                                  _p(tr['used_gpus'], C.c_double), _p(tr['gpc'], C.c_int32), _p(tr['mem_mib'], C.c_double),
                                  _p(tr['util_avg'], C.c_double), C.c_int64(cap), C.c_int(n_runs), C.c_int(n_threads))
    if ev < 0:
        raise RuntimeError('oracle_fifo_yarn_batch rc=%d' % ev)
    return int(ev)


LEGACY_ROW_DTYPE = np.dtype([('time', 'i4'), ('idle_nodes', 'i4'), ('full_nodes', 'i4'), ('busy_gpus', 'i4'),
                             ('pending', 'i4'), ('running', 'i4'), ('completed', 'i4'), ('pad', 'i4')])


def _run_legacy(fn_name, cluster, tr, extra_args, rows_cap=None):
    L = lib()
    n = len(tr['nt'])
    outs = [np.full(max(n, 1), -1, np.int32) for _ in range(6)]   # finish_order, start, end, pending, preempt, resume
    nfin = C.c_int32(0); nev = C.c_int64(0); counters = np.zeros(4, np.int64)
    cap = rows_cap or max(4096, 8 * n)
    while True:
        rows = np.zeros(cap, LEGACY_ROW_DTYPE)
        args = [C.byref(cluster), C.c_int32(n), _p(tr['nt'], C.c_double), _p(tr['duration'], C.c_double),
                _p(tr['used_gpus'], C.c_double)] + extra_args + [_p(o, C.c_int32) for o in outs] + [
                C.byref(nfin), rows.ctypes.data_as(C.c_void_p), C.c_int64(cap), C.byref(nev), _p(counters, C.c_int64)]
        rc = getattr(L, fn_name)(*args)
        if rc == -1:
            cap *= 4
            continue
        if rc != 0:
            raise RuntimeError('%s rc=%d' % (fn_name, rc))
        break
    k = nfin.value
    return dict(finish_order=outs[0][:k].copy(), start=outs[1][:n], end=outs[2][:n], pending=outs[3][:n], preempt=outs[4][:n],
                resume=outs[5][:n], rows=rows[:nev.value], n_events=nev.value,
                counters=dict(sweep_jobs=int(counters[0]), events=int(counters[1]), demotions=int(counters[2])))


def run_sjf_yarn(cluster, tr, rows_cap=None, sort_mode=0):
    """Restated smallest_first_sim_jobs (run_sim.py:162-287; sort_mode 0) / shortest_first_sim_jobs
    (run_sim.py:299-431; 1 = shortest, 2 = shortest-gpu) with the live yarn fit.  Pinned against the reference's dead
    code run unmodified under shim globals (oracle/ref_legacy_runner.py, tests/golden/sjf_*, shortest*_*)."""
    return _run_legacy('oracle_sjf_yarn', cluster, tr, [_p(tr['gpc'], C.c_int32), _p(tr['mem_mib'], C.c_double), C.c_int32(sort_mode)], rows_cap)


def run_dlas_gpu(cluster, tr, queue_limit=(30, 60, 150), rows_cap=None, gputime=True):
    """Restated dlas_sim_jobs (run_sim.py:664-947), count-based admission; gputime=False is `--schedule dlas`.
    Pinned against the reference's dead code run unmodified under shim globals (tests/golden/dlasgpu_*, dlas_*)."""
    ql = np.asarray(queue_limit, np.int32)
    return _run_legacy('oracle_dlas', cluster, tr, [C.c_int32(int(gputime)), C.c_int32(len(ql) + 1), _p(ql, C.c_int32)], rows_cap)


def format_legacy_job_csv(tr, res, count_scheme):
    """job.csv in the legacy layout (/root/reference/log.py:86-88,316-330)."""
    buf = io.StringIO(newline='')
    w = csv.writer(buf)
    hdr = ['time', 'job_id', 'num_gpu', 'submit_time', 'start_time', 'end_time', 'executed_time', 'JCT', 'duration',
           'pending_time', 'preempt'] + (['resume'] if count_scheme else []) + ['promote']
    w.writerow(hdr)
    sub = np.ceil(tr['nt']).astype(np.int64)
    dur = np.maximum(1, np.ceil(tr['duration'])).astype(np.int64)
    for i in res['finish_order']:
        i = int(i)
        row = [int(res['end'][i]), str(int(tr['label'][i])), int(np.ceil(tr['used_gpus'][i])), int(sub[i]), int(res['start'][i]),
               int(res['end'][i]), int(res['end'][i] - res['start'][i]), int(res['end'][i] - sub[i]), int(dur[i]),
               int(res['pending'][i]), int(res['preempt'][i])] + ([int(res['resume'][i])] if count_scheme else []) + [0]
        w.writerow(row)
    return buf.getvalue()


def format_legacy_cluster_csv(res, cluster, count_scheme):
    """cluster.csv in the legacy layout (/root/reference/log.py:43-45,137-258).  Under --scheme count the node / gpu columns
    come from CLUSTER.free_gpu (log.py:233-238); under a placement scheme the reference's code for them is commented out
    (log.py:171-189) and the columns stay 0."""
    buf = io.StringIO(newline='')
    w = csv.writer(buf)
    w.writerow(['time', 'idle_node', 'busy_node', 'full_node', 'idle_gpu', 'busy_gpu', 'pending_job', 'running_job', 'completed_job'])
    N = cluster.num_switch * cluster.num_node_p_switch
    D = N * cluster.num_gpu_p_node
    for r in res['rows']:
        if count_scheme:
            w.writerow([int(r['time']), int(r['idle_nodes']), N - int(r['idle_nodes']), int(r['full_nodes']), D - int(r['busy_gpus']),
                        int(r['busy_gpus']), int(r['pending']), int(r['running']), int(r['completed'])])
        else:
            w.writerow([int(r['time']), 0, 0, 0, 0, 0, int(r['pending']), int(r['running']), int(r['completed'])])
    return buf.getvalue()


def run_legacy(cluster, tr, schedule, queue_limit=(30, 60, 150)):
    """Dispatch by --schedule name; returns (result, count_scheme)."""
    if schedule in ('dlas-gpu', 'dlas'):
        return run_dlas_gpu(cluster, tr, queue_limit, gputime=schedule == 'dlas-gpu'), True
    return run_sjf_yarn(cluster, tr, sort_mode={'sjf': 0, 'shortest': 1, 'shortest-gpu': 2}[schedule]), False


def run_env_yarn(cluster, tr, policy, window_k=5, seed=0, replica=0, actions=None, rows_cap=None):
    """Build-defined RL environment semantics on the CPU (oracle_env_yarn). policy: 0 head, 1 random window,
    2 = `actions` tape (one action per tick).  PARITY UNPINNED (model/env.py is a stub in the reference)."""
    L = lib()
    n = len(tr['nt'])
    fin = np.empty(max(n, 1), np.int32); st = np.empty(max(n, 1), np.int32); en = np.empty(max(n, 1), np.int32)
    nfin = C.c_int32(0); nticks = C.c_int64(0); counters = np.zeros(4, np.int64)
    acts = np.ascontiguousarray(actions, np.int32) if actions is not None else np.zeros(1, np.int32)
    cap = rows_cap or max(4096, 4 * n)
    while True:
        rows = np.zeros(cap, ROW_DTYPE)
        rc = L.oracle_env_yarn(C.byref(cluster), C.c_int32(n), _p(tr['nt'], C.c_double), _p(tr['duration'], C.c_double),
                               _p(tr['used_gpus'], C.c_double), _p(tr['gpc'], C.c_int32), _p(tr['mem_mib'], C.c_double),
                               _p(tr['util_avg'], C.c_double), C.c_int32(policy), C.c_int32(window_k), C.c_uint32(seed),
                               C.c_uint32(replica), _p(acts, C.c_int32), C.c_int64(len(acts) if actions is not None else 0),
                               _p(fin, C.c_int32), _p(st, C.c_int32), _p(en, C.c_int32), C.byref(nfin),
                               rows.ctypes.data_as(C.c_void_p), C.c_int64(cap), C.byref(nticks), _p(counters, C.c_int64))
        if rc == -1:
            cap *= 4
            continue
        if rc != 0:
            raise RuntimeError('oracle_env_yarn rc=%d' % rc)
        break
    k = nfin.value
    return dict(finish_order=fin[:k].copy(), start=st[:n], end=en[:n], rows=rows[:nticks.value], n_ticks=nticks.value,
                counters=dict(sum_queued=int(counters[0]), sum_running=int(counters[1]), ticks=int(counters[2]), starts=int(counters[3])))
