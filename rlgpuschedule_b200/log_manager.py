"""CSV outputs in the reference's exact format, produced from the device's integer statistics.

Host mirror of /log_manager.py of the reference (LogInfo :5-30, LogManager.init :56-116,
step_cluster :118-135, jcts :137-155): same six files, same headers, same column order, same
`str(value)` rendering through the csv module ('\\r\\n' line ends).  The float columns are
finished here in float64 with the reference's own expressions:
  avg_gpu_memory_allocated = sum(min(cap, mem_max)) / sum(cap)          schedule.py:109-121
  avg_pending_time         = sum(pending) / (count + 1e-9)              jobs_manager.py:87
  median_pending_time      = float(np.median(pending))  (nan when empty) jobs_manager.py:87
  max_pending_time         = max(pending, 0)  (the int 0 when empty)     jobs_manager.py:80-82
avg_gpu_utilization is an unseeded normal draw per busy device in the reference
(infra/device.py:48-54); here it is one draw per row from the summed distribution (util_mode
'sample', seedable) or its mean ('mean').
"""
import csv
import io
import os

import numpy as np

CLUSTER_HEADER = ['delta', 'num_idle_nodes', 'num_busy_nodes', 'num_busy_gpus', 'num_idle_gpus',
                  'avg_gpu_utilization', 'avg_gpu_memory_allocated', 'avg_pending_time', 'median_pending_time',
                  'max_pending_time', 'num_running_jobs', 'num_queuing_jobs', 'num_finish_jobs']
JOB_HEADER = ['job_id', 'num_gpu', 'submit_time', 'start_time', 'end_time', 'original_duration',
              'actual_duration', 'jct', 'preempt']


LEGACY_CLUSTER_HEADER = ['time', 'idle_node', 'busy_node', 'full_node', 'idle_gpu', 'busy_gpu', 'pending_job',
                         'running_job', 'completed_job']                                  # log.py:43-45
LEGACY_JOB_HEADER_COUNT = ['time', 'job_id', 'num_gpu', 'submit_time', 'start_time', 'end_time', 'executed_time', 'JCT',
                           'duration', 'pending_time', 'preempt', 'resume', 'promote']        # log.py:86
LEGACY_JOB_HEADER = [h for h in LEGACY_JOB_HEADER_COUNT if h != 'resume']                  # log.py:88


class LogInfo(object):
    """Same fields as the reference's LogInfo (log_manager.py:5-30)."""

    def __init__(self, num_idle_nodes, num_busy_nodes, num_busy_gpus, num_idle_gpus, avg_gpu_utilization,
                 avg_gpu_memory_allocated, avg_pending_time, median_pending_time, max_pending_time,
                 num_running_jobs, num_queuing_jobs, num_finish_jobs):
        self.idle_ns = num_idle_nodes
        self.busy_ns = num_busy_nodes
        self.busy_gs = num_busy_gpus
        self.idle_gs = num_idle_gpus
        self.avg_g_utils = avg_gpu_utilization
        self.avg_g_mem = avg_gpu_memory_allocated
        self.avg_pending = avg_pending_time
        self.median_pending = median_pending_time
        self.max_pending = max_pending_time
        self.num_running_jobs = num_running_jobs
        self.num_queuing_jobs = num_queuing_jobs
        self.num_finish_jobs = num_finish_jobs


def finish_rows(rows, cluster, mem_shift, util_mode='sample', seed=None):
    """rows: _ffi.ROW_DTYPE array -> dict of python-ready columns (lists)."""
    n = len(rows)
    N, Dv = cluster.num_nodes, cluster.num_gpus
    q = rows['queued'].astype(np.float64)
    mem = (rows['mem_sum'].astype(np.float64) / float(2 ** mem_shift)) / float(Dv * cluster.cap_mib)
    with np.errstate(invalid='ignore'):
        avg_p = rows['sum_pending'].astype(np.float64) / (q + 1e-9)
    med = (rows['median_lo'].astype(np.float64) + rows['median_hi'].astype(np.float64)) / 2.0
    med = np.where(rows['queued'] > 0, med, np.nan)
    busy = rows['busy_gpus']
    mu = rows['util_mu_sum'].astype(np.float64) / 512.0
    if util_mode == 'sample':
        sd = np.sqrt(rows['util_var_sum'].astype(np.float64)) / 512.0
        draw = np.random.default_rng(seed).normal(mu, sd)
        util = np.clip(draw, 0.0, 100.0 * busy) / Dv
    else:
        util = mu / Dv
    maxp = rows['max_pending']
    qi = rows['queued']
    return dict(n=n, idle_nodes=rows['idle_nodes'].tolist(), busy_nodes=(N - rows['idle_nodes']).tolist(),
                busy_gpus=busy.tolist(), idle_gpus=(Dv - busy).tolist(),
                util=['[%s]' % np.format_float_positional(u, precision=8, unique=True, fractional=True, trim='-')
                      if b > 0 else '0.0' for u, b in zip(util.tolist(), busy.tolist())],
                mem=mem.tolist(), avg_pending=avg_p.tolist(), median=med.tolist(),
                max_pending=[float(m) if k > 0 else 0 for m, k in zip(maxp.tolist(), qi.tolist())],
                running=rows['running'].tolist(), queued=qi.tolist(), finished=rows['finished'].tolist())


def format_cluster_csv(rows, cluster, mem_shift, with_header=True, with_util=True, util_mode='sample', seed=None,
                       first_delta=1):
    """cluster.csv text for rows (one per tick, delta = first_delta + i)."""
    c = finish_rows(rows, cluster, mem_shift, util_mode, seed)
    buf = io.StringIO(newline='')
    w = csv.writer(buf)
    hdr = list(CLUSTER_HEADER)
    if not with_util:
        hdr.remove('avg_gpu_utilization')
    if with_header:
        w.writerow(hdr)
    cols = [range(first_delta, first_delta + c['n']), c['idle_nodes'], c['busy_nodes'], c['busy_gpus'], c['idle_gpus']]
    if with_util:
        cols.append(c['util'])
    cols += [c['mem'], c['avg_pending'], c['median'], c['max_pending'], c['running'], c['queued'], c['finished']]
    w.writerows(zip(*cols))
    return buf.getvalue()


def format_job_csv(trace, finish_order, start, end, preempt=None, actual_duration=None, jct=None, with_header=True,
                   get_duration=None):
    """job.csv text: one row per finished job in finish order (finished_jobs dict order, log_manager.py:141-153)."""
    buf = io.StringIO(newline='')
    w = csv.writer(buf)
    if with_header:
        w.writerow(JOB_HEADER)
    fo = np.asarray(finish_order, dtype=np.int64)
    # with network costs Job.duration itself is increased (job.py:196-197), so both columns show the new value
    dur = trace.duration[fo] if actual_duration is None else np.asarray(actual_duration)[fo]
    # horus: Job.duration stays, Job.get_duration() grows by 5 when a task was de-interfered (jobs_manager.py:184-185)
    act = dur if get_duration is None else np.asarray(get_duration)[fo]
    st, en = np.asarray(start)[fo], np.asarray(end)[fo]
    jc = (en - st) if jct is None else np.asarray(jct)[fo]
    pre = np.ones(len(fo), dtype=np.int64) if preempt is None else np.asarray(preempt)[fo]
    w.writerows(zip([str(x) for x in trace.label[fo].tolist()], trace.used_gpus[fo].tolist(),
                    trace.nt[fo].astype(np.int64).tolist(), st.tolist(), en.tolist(), dur.tolist(),
                    [a if a > 0 else 0 for a in act.tolist()],  # Job.get_duration(): max(0, d) keeps the int 0
                    jc.tolist(), pre.tolist()))
    return buf.getvalue()


def format_legacy_cluster_csv(rows, cluster, count_scheme, node_stats=False, with_header=True):
    """cluster.csv of the event-driven schedules as the reference's log._Log.checkpoint writes it (log.py:137-258).
    Under --scheme count the node / gpu columns come from CLUSTER.free_gpu (log.py:225-238).  Under a placement scheme the
    reference's code for those columns is commented out (log.py:171-189) and it prints 0; node_stats=True prints what
    that commented code would (the device keeps the numbers in the row)."""
    N, Dv = cluster.num_nodes, cluster.num_gpus
    idle = rows['idle_nodes']; full = rows['median_hi']; busy_g = rows['busy_gpus']
    n = len(rows)
    zero = [0] * n
    if count_scheme:
        cols = [idle.tolist(), full.tolist(), full.tolist(), (Dv - busy_g).tolist(), busy_g.tolist()]
    elif node_stats:
        cols = [idle.tolist(), (N - idle - full).tolist(), full.tolist(), (Dv - busy_g).tolist(), busy_g.tolist()]
    else:
        cols = [zero, zero, zero, zero, zero]
    buf = io.StringIO(newline='')
    w = csv.writer(buf)
    if with_header:
        w.writerow(LEGACY_CLUSTER_HEADER)
    w.writerows(zip(rows['median_lo'].tolist(), *cols, rows['queued'].tolist(), rows['running'].tolist(), rows['finished'].tolist()))
    return buf.getvalue()


def format_legacy_job_csv(trace, jobs, pending, resume, count_scheme, with_header=True):
    """job.csv of the event-driven schedules as log._Log.job_complete writes it (log.py:86-88,316-330)."""
    fo = np.asarray(jobs['finish_order'], dtype=np.int64)
    sub = trace.records['arrival_tick'][fo].astype(np.int64)
    st, en = jobs['start'][fo].astype(np.int64), jobs['end'][fo].astype(np.int64)
    cols = [en.tolist(), [str(x) for x in trace.label[fo].tolist()], trace.records['gpus'][fo].tolist(), sub.tolist(),
            st.tolist(), en.tolist(), (en - st).tolist(), (en - sub).tolist(), trace.records['dur_ticks'][fo].tolist(),
            np.asarray(pending)[fo].tolist(), np.asarray(jobs['preempt'])[fo].tolist()]
    if count_scheme:
        cols.append(np.asarray(resume)[fo].tolist())
    cols.append([0] * len(fo))
    buf = io.StringIO(newline='')
    w = csv.writer(buf)
    if with_header:
        w.writerow(LEGACY_JOB_HEADER_COUNT if count_scheme else LEGACY_JOB_HEADER)
    w.writerows(zip(*cols))
    return buf.getvalue()


class LogManager(object):
    """Drop-in for the reference's LogManager: same constructor, init(), step_cluster(), jcts();
    plus bulk writers used by the device backend."""

    def __init__(self, log_path, flags):
        self.log_path = log_path
        self.flags = flags
        self.is_count = getattr(flags, 'scheme', 'yarn') == 'count'
        self.cluster_stats_header = list(CLUSTER_HEADER)
        self.job_stats_header = list(JOB_HEADER)

    def init(self, infrastructure, legacy=None):
        """legacy=True writes the headers of the Tiresias-era logger (log.py:30-88 of the reference), the
        format the sjf / dlas-gpu loops logged in; default: decided from flags.schedule."""
        self.log_cluster = os.path.join(self.log_path, 'cluster.csv')
        self.log_job = os.path.join(self.log_path, 'job.csv')
        if legacy is None:
            legacy = getattr(self.flags, 'schedule', 'fifo') in ('sjf', 'shortest', 'shortest-gpu', 'dlas-gpu', 'dlas')
        self.legacy = legacy
        if legacy:
            self.cluster_stats_header = list(LEGACY_CLUSTER_HEADER)
            count = getattr(self.flags, 'schedule', '') in ('dlas-gpu', 'dlas') or self.is_count
            self.job_stats_header = list(LEGACY_JOB_HEADER_COUNT if count else LEGACY_JOB_HEADER)
        n_nodes = infrastructure.num_nodes
        n_gpus = infrastructure.num_gpus
        with open(self.log_cluster, 'w+', newline='') as f:
            csv.writer(f).writerow(self.cluster_stats_header)
        if not self.is_count:
            self.log_cpu = os.path.join(self.log_path, 'cpu.csv')
            self.log_gpu = os.path.join(self.log_path, 'gpu.csv')
            self.log_network = os.path.join(self.log_path, 'network.csv')
            self.log_mem = os.path.join(self.log_path, 'memory.csv')
            with open(self.log_cpu, 'w+', newline='') as f:
                csv.writer(f).writerow(['time'] + ['cpu%d' % i for i in range(n_nodes)])
            with open(self.log_gpu, 'w+', newline='') as f:
                csv.writer(f).writerow(['time'] + ['gpu%d' % i for i in range(n_gpus)])
            with open(self.log_mem, 'w+', newline='') as f:
                csv.writer(f).writerow(['time', 'max', '99th', '95th', 'med'])
            with open(self.log_network, 'w+', newline='') as f:
                titles = ['time']
                for i in range(n_nodes):
                    titles += ['in%d' % i, 'out%d' % i]
                csv.writer(f).writerow(titles)
        with open(self.log_job, 'w+', newline='') as f:
            csv.writer(f).writerow(self.job_stats_header)

    def step_cluster(self, loginfo, delta):
        with open(self.log_cluster, 'a+', newline='') as f:
            csv.writer(f).writerow([delta, loginfo.idle_ns, loginfo.busy_ns, loginfo.busy_gs, loginfo.idle_gs,
                                    loginfo.avg_g_utils, loginfo.avg_g_mem, loginfo.avg_pending,
                                    loginfo.median_pending, loginfo.max_pending, loginfo.num_running_jobs,
                                    loginfo.num_queuing_jobs, loginfo.num_finish_jobs])

    def write_cluster_rows(self, rows, cluster, mem_shift, util_mode='sample', seed=None):
        with open(self.log_cluster, 'a+', newline='') as f:
            f.write(format_cluster_csv(rows, cluster, mem_shift, with_header=False, util_mode=util_mode, seed=seed))

    def write_columnar(self, rows, cluster, trace, jobs, mem_shift, get_duration=None, jct=None, util_mode='mean', seed=None):
        """cluster.parquet / job.parquet next to the CSVs: the same columns as typed arrays (with the
        simulation on the device, rendering ~10^5 float rows as text dominates a run).  Float columns carry the same float64
        values the CSV prints; avg_gpu_utilization is the float, not the reference's one-element-array repr."""
        import pyarrow as pa
        import pyarrow.parquet as pq
        c = finish_rows(rows, cluster, mem_shift, util_mode, seed)
        n = c['n']
        busy = np.asarray(c['busy_gpus'])
        mu = rows['util_mu_sum'].astype(np.float64) / 512.0 / cluster.num_gpus
        util = mu if util_mode != 'sample' else np.array([float(u.strip('[]')) for u in c['util']])
        cl = pa.table({'delta': np.arange(1, n + 1, dtype=np.int64), 'num_idle_nodes': np.asarray(c['idle_nodes'], np.int32),
                       'num_busy_nodes': np.asarray(c['busy_nodes'], np.int32), 'num_busy_gpus': busy.astype(np.int32),
                       'num_idle_gpus': np.asarray(c['idle_gpus'], np.int32), 'avg_gpu_utilization': util,
                       'avg_gpu_memory_allocated': np.asarray(c['mem'], np.float64), 'avg_pending_time': np.asarray(c['avg_pending'], np.float64),
                       'median_pending_time': np.asarray(c['median'], np.float64), 'max_pending_time': rows['max_pending'].astype(np.float64),
                       'num_running_jobs': rows['running'].astype(np.int32), 'num_queuing_jobs': rows['queued'].astype(np.int32),
                       'num_finish_jobs': rows['finished'].astype(np.int32)})
        pq.write_table(cl, os.path.join(self.log_path, 'cluster.parquet'))
        fo = np.asarray(jobs['finish_order'], dtype=np.int64)
        st, en = np.asarray(jobs['start'])[fo], np.asarray(jobs['end'])[fo]
        dur = trace.duration[fo]
        act = dur if get_duration is None else np.asarray(get_duration)[fo]
        jb = pa.table({'job_id': trace.label[fo].astype(np.int64), 'num_gpu': trace.used_gpus[fo], 'submit_time': trace.nt[fo].astype(np.int64),
                       'start_time': st.astype(np.int64), 'end_time': en.astype(np.int64), 'original_duration': dur,
                       'actual_duration': np.maximum(act, 0.0), 'jct': ((en - st) if jct is None else np.asarray(jct)[fo]).astype(np.int64),
                       'preempt': np.asarray(jobs['preempt'])[fo].astype(np.int32)})
        pq.write_table(jb, os.path.join(self.log_path, 'job.parquet'))

    def write_legacy(self, rows, cluster, trace, jobs, pending, resume, count_scheme, node_stats=False):
        """cluster.csv / job.csv of the event-driven schedules (LOG.checkpoint log.py:137-258, LOG.job_complete
        log.py:316-330).  rows: _ffi.ROW_DTYPE with the legacy field mapping documented in include/rlgs.h."""
        with open(self.log_cluster, 'a+', newline='') as f:
            f.write(format_legacy_cluster_csv(rows, cluster, count_scheme, node_stats=node_stats, with_header=False))
        with open(self.log_job, 'a+', newline='') as f:
            f.write(format_legacy_job_csv(trace, jobs, pending, resume, count_scheme, with_header=False))

    def jcts(self, finished_jobs):
        """finished_jobs: dict job_id -> object with the reference Job's attributes, or a tuple
        (trace, finish_order, start, end[, preempt]) from the device backend."""
        if isinstance(finished_jobs, tuple):
            has_kw = isinstance(finished_jobs[-1], dict)
            kw = finished_jobs[-1] if has_kw else {}
            args = finished_jobs[:-1] if has_kw else finished_jobs
            text = format_job_csv(*args, with_header=False, **kw)
            assert len(finished_jobs[1]) > 0, ValueError("No finished jobs")
            with open(self.log_job, 'a+', newline='') as f:
                f.write(text)
            return
        assert len(finished_jobs) > 0, ValueError("No finished jobs")
        with open(self.log_job, 'a+', newline='') as f:
            w = csv.writer(f)
            for _, j in finished_jobs.items():
                w.writerow([j.job_id, j.gpus, j.submit_time, j.start_time, j.end_time, j.duration,
                            j.get_duration(), j.time_processed(), j.migration_count])
