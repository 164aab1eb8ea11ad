"""High-level handle over the C ABI: one object = one batch of replicas on one GPU."""
import ctypes as C

import numpy as np

from . import _ffi
from .ingest import Cluster, Trace


class Simulator(object):
    """Replaces the object graph the reference builds in run_sim.py:1716-1735 (Infrastructure,
    JobQueueManager, JobsManager, Scheduler) by a device-resident state; `run()` replaces
    Scheduler.start() (core/scheduling/schedule.py:178-216)."""

    def __init__(self, cluster, schedule='fifo', scheme='yarn', n_replicas=1, rows=True, device=0, slot_cap=0,
                 n_streams=0, ticks_per_launch=0, rows_cap=0, fetch_jobs=False, num_queue=1, queue_limit=(),
                 max_ticks=0, enable_network_costs=False, bandwidth=1250, internode_latency=0.015, num_buffer=5,
                 pack_seed=None, pack_rng=None, rows_format=None, lanes_per_replica=0):
        """rows: True / 'host' = rows copied to the pinned host store inside run(); 'device' = rows stay in
        HBM until asked for; False = no rows.  num_buffer: look-ahead window of the horus schedule (--num_buffer).
        pack_seed: None = utilisation draws of the horus score return their mean (the reference's behaviour on
        zero-spread traces); an int seeds the build-defined counter-based draw.  horus+: pack_seed also seeds the k-means draws
        (core/jobs/utils.py:39,60); pack_rng=False keeps the utilisation draws at their mean while the k-means stays seeded.
        rows_format: 'event4' (default for fifo without network costs: a 4-byte event row per tick — idle nodes + "the queue head
        started"; the host replays the queue to rebuild the job tables and the pending-time columns when asked) / 'event16' (12-byte
        row + the job started at the tick) / 'wire12' / 'wire16' keep compact rows on the device and on the wire and expand them in
        rows(); 'wide' keeps the self-contained 64-byte rows.  lanes_per_replica: 8 / 16 / 32 lanes of a warp per replica, 0 = auto.
        fetch_jobs: True = start / end / finish_order tables copied to the host inside run(); 'end' = end and finish_order only
        (fifo without network costs: a finished job started at end - dur_ticks)."""
        if schedule not in _ffi.SCHED:
            raise NotImplementedError('schedule %r has no device implementation' % (schedule,))
        if scheme not in _ffi.PLACE:
            raise NotImplementedError('placement scheme %r has no device implementation' % (scheme,))
        if rows_format is None:
            rows_format = 'wide'
            if schedule == 'fifo' and cluster.num_nodes <= 4095:
                rows_format = 'wire12' if enable_network_costs else 'event4'
        if rows_format not in _ffi.ROWFMT:
            raise ValueError('rows_format must be one of %s' % sorted(_ffi.ROWFMT))
        self.cluster = cluster
        self.n_replicas = n_replicas
        self.rows_mode = rows
        self._kw = dict(schedule=schedule, scheme=scheme, rows=rows, device=device, n_streams=n_streams,
                        ticks_per_launch=ticks_per_launch, rows_cap=rows_cap, fetch_jobs=fetch_jobs,
                        num_queue=num_queue, queue_limit=tuple(queue_limit), max_ticks=max_ticks,
                        enable_network_costs=enable_network_costs, bandwidth=bandwidth,
                        internode_latency=internode_latency, num_buffer=num_buffer, pack_seed=pack_seed,
                        pack_rng=(pack_seed is not None) if pack_rng is None else bool(pack_rng),
                        rows_format=rows_format, lanes_per_replica=int(lanes_per_replica))
        self._slot_cap = slot_cap
        self._traces = []   # (first, count, Trace)
        self._h = None
        self._create()

    def _create(self):
        L = _ffi.lib()
        k = self._kw
        o = _ffi.Opts()
        o.device = k['device']; o.n_replicas = self.n_replicas
        o.schedule = _ffi.SCHED[k['schedule']]; o.placement = _ffi.PLACE[k['scheme']]
        o.rows_mode = (_ffi.ROWS_DEVICE if k['rows'] == 'device' else _ffi.ROWS_FULL) if k['rows'] else _ffi.ROWS_NONE
        o.slot_cap = self._slot_cap; o.n_streams = k['n_streams']; o.ticks_per_launch = k['ticks_per_launch']
        o.rows_cap = k['rows_cap']; o.fetch_jobs = 2 if k['fetch_jobs'] == 'end' else int(bool(k['fetch_jobs'])); o.num_queue = k['num_queue']
        for i, v in enumerate(k['queue_limit'][:_ffi.MAX_QUEUES]):
            o.queue_limit[i] = int(v)
        o.enable_network_costs = int(bool(k['enable_network_costs']))
        o.bandwidth = float(k['bandwidth']); o.internode_latency = float(k['internode_latency'])
        o.max_ticks = int(k['max_ticks'])
        o.num_buffer = int(k['num_buffer']); o.pack_rng = int(k['pack_rng']); o.pack_seed = int(k['pack_seed'] or 0)
        o.rows_format = _ffi.ROWFMT[k['rows_format']]
        o.lanes_per_replica = k['lanes_per_replica']
        spec = self.cluster.to_ffi()
        h = C.c_void_p()
        _ffi.check(L.rlgs_create(C.byref(spec), C.byref(o), C.byref(h)))
        self._h = h

    def close(self):
        if self._h is not None:
            _ffi.lib().rlgs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_trace(self, trace, first_replica=0, n_replicas=None):
        n_replicas = self.n_replicas - first_replica if n_replicas is None else n_replicas
        rec = np.ascontiguousarray(trace.records)
        net = None
        if self._kw['enable_network_costs']:
            dp = C.POINTER(C.c_double)
            net = _ffi.NetcostInputs(trace.duration.ctypes.data_as(dp), trace.model_mb.ctypes.data_as(dp),
                                     trace.iterations.ctypes.data_as(dp))
        rc = _ffi.lib().rlgs_load_trace(self._h, first_replica, n_replicas, rec.ctypes.data, len(rec),
                                        C.byref(net) if net is not None else None)
        if rc == _ffi.ERR_WIRE and self._kw['rows_format'] != 'wide':
            self._rebuild(rows_format='wide')   # more jobs than the wire row can count: self-contained rows
            return self.load_trace(trace, first_replica, n_replicas)
        _ffi.check(rc)
        if self._kw['schedule'] in ('horus', 'gandiva', 'horus+'):
            pi = trace.pack_inputs()
            dp = C.POINTER(C.c_double)
            inp = _ffi.PackInputs(pi['util_avg'].ctypes.data_as(dp), pi['util_sd'].ctypes.data_as(dp),
                                  pi['task_mem'].ctypes.data_as(C.POINTER(C.c_int64)), pi['heap_cap'].ctypes.data_as(C.POINTER(C.c_int32)),
                                  trace.mem_shift, trace.cap_mib, pi['util_max'].ctypes.data_as(dp), pi['mem_avg_mib'].ctypes.data_as(dp),
                                  pi['used_gpus'].ctypes.data_as(dp))
            _ffi.check(_ffi.lib().rlgs_load_pack_inputs(self._h, first_replica, n_replicas, C.byref(inp), len(rec)))
        self._traces = [x for x in self._traces if (x[0], x[1]) != (first_replica, n_replicas)]
        self._traces.append((first_replica, n_replicas, trace))

    def trace_of(self, replica):
        for f, c, t in self._traces[::-1]:
            if f <= replica < f + c:
                return t
        raise KeyError(replica)

    def _rebuild(self, **changes):
        traces = self._traces
        self.close()
        self._slot_cap = changes.pop('slot_cap', self._slot_cap)
        self._kw.update(changes)
        self._traces = []
        self._create()
        for f, c, t in traces:
            self.load_trace(t, f, c)

    def run(self):
        L = _ffi.lib()
        while True:
            rc = L.rlgs_run(self._h)
            if rc == _ffi.ERR_SLOTS:
                # more jobs ran concurrently than on-chip slots: rebuild with a larger table
                cap = max(64, 2 * (self._slot_cap or 128))
                if cap > 2 * max(self.cluster.num_gpus, 32):
                    _ffi.check(rc)
                self._rebuild(slot_cap=cap)
                continue
            if rc == _ffi.ERR_WIRE and self._kw['rows_format'] != 'wide':
                self._rebuild(rows_format='wide')   # a run longer than 2^24 ticks: self-contained rows
                continue
            _ffi.check(rc)
            return self

    def kernel_ms(self):
        ms, n = C.c_float(0), C.c_int32(0)
        _ffi.check(_ffi.lib().rlgs_last_run_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def summary(self, replica=0):
        s = _ffi.Summary()
        _ffi.check(_ffi.lib().rlgs_get_summary(self._h, replica, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def jobs(self, replica=0, with_nodes=False):
        n = len(self.trace_of(replica))
        fo, st, en, pre = (np.empty(n, np.int32) for _ in range(4))
        fn = np.empty(n, np.int32) if with_nodes else None
        _ffi.check(_ffi.lib().rlgs_read_jobs(self._h, replica, fo.ctypes.data, st.ctypes.data, en.ctypes.data,
                                             pre.ctypes.data, fn.ctypes.data if with_nodes else None))
        k = self.summary(replica)['n_finished']
        out = dict(finish_order=fo[:k], start=st, end=en, preempt=pre)
        if with_nodes:
            out['first_node'] = fn
        return out

    def rows(self, replica=0):
        """All rows of a replica as one numpy array (copied out of the chunk-major store; 'wire16' handles expand the
        16-byte wire rows to full rows here, on the host, see rlgs_row16 in include/rlgs.h)."""
        n = self.summary(replica)['n_ticks']
        out = np.zeros(n, _ffi.ROW_DTYPE)
        if n:
            _ffi.check(_ffi.lib().rlgs_read_rows(self._h, replica, 0, n, out.ctypes.data))
        return out

    rows_view = rows

    def rows_chunk_view(self, replica, chunk):
        """Zero-copy numpy view of one 4096-row chunk of a replica in the pinned host mirror: ROW_DTYPE for 'wide' handles,
        ROW16_DTYPE / ROW12_DTYPE / ROW4_DTYPE (the packed wire rows, see include/rlgs.h) for the other formats."""
        p, n = C.c_void_p(), C.c_int64(0)
        L = _ffi.lib()
        fn, dt = {'wide': (L.rlgs_rows_view, _ffi.ROW_DTYPE), 'wire16': (L.rlgs_rows16_view, _ffi.ROW16_DTYPE),
                  'wire12': (L.rlgs_rows12_view, _ffi.ROW12_DTYPE), 'event16': (L.rlgs_rows16e_view, _ffi.ROW16_DTYPE),
                  'event4': (L.rlgs_rows4e_view, _ffi.ROW4_DTYPE)}[self._kw['rows_format']]
        _ffi.check(fn(self._h, replica, chunk, C.byref(p), C.byref(n)))
        buf = (C.c_char * (n.value * dt.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=n.value)

    def rows_wire(self, replica=0):
        """The packed wire rows of a replica (every format but 'wide'), unexpanded."""
        n = self.summary(replica)['n_ticks']
        L = _ffi.lib()
        fn, dt = {'wire16': (L.rlgs_read_rows16, _ffi.ROW16_DTYPE), 'wire12': (L.rlgs_read_rows12, _ffi.ROW12_DTYPE),
                  'event16': (L.rlgs_read_rows16e, _ffi.ROW16_DTYPE), 'event4': (L.rlgs_read_rows4e, _ffi.ROW4_DTYPE)}[self._kw['rows_format']]
        out = np.zeros(n, dt)
        if n:
            _ffi.check(fn(self._h, replica, 0, n, out.ctypes.data))
        return out

    rows16 = rows_wire

    def durations(self, replica=0):
        """Per-job duration after network costs (enable_network_costs only)."""
        out = np.empty(len(self.trace_of(replica)), np.float64)
        _ffi.check(_ffi.lib().rlgs_read_durations(self._h, replica, out.ctypes.data))
        return out

    def job_plane(self, replica, plane):
        out = np.empty(len(self.trace_of(replica)), np.int32)
        _ffi.check(_ffi.lib().rlgs_read_job_plane(self._h, replica, plane, out.ctypes.data))
        return out

    def returns(self):
        out = np.empty(self.n_replicas, np.int64)
        _ffi.check(_ffi.lib().rlgs_returns(self._h, out.ctypes.data))
        return out

    def returns_device_ptr(self):
        p = C.c_void_p()
        _ffi.check(_ffi.lib().rlgs_returns_device_ptr(self._h, C.byref(p)))
        return p.value


def replay_event_rows(trace, rows):
    """Host-only expansion of one replica's 4-byte event rows (rows_format='event4', e.g. kept from Simulator.rows_wire() or
    rows_chunk_view()): returns dict(start, end, finish_order, max_pending, median_lo, median_hi).  No device call
    (rlgs_replay_rows4e, include/rlgs.h)."""
    rec = np.ascontiguousarray(trace.records)
    w = np.ascontiguousarray(rows['w'] if rows.dtype.names else rows, dtype='<u4')
    J, n = len(rec), len(w)
    st, en, fo = (np.empty(J, np.int32) for _ in range(3))
    pend = np.empty(3 * max(n, 1), np.int32)
    nf = C.c_int32(0)
    _ffi.check(_ffi.lib().rlgs_replay_rows4e(rec.ctypes.data, J, w.ctypes.data, n, st.ctypes.data, en.ctypes.data, fo.ctypes.data,
                                              C.byref(nf), pend.ctypes.data))
    pend = pend[:3 * n].reshape(n, 3)
    return dict(start=st, end=en, finish_order=fo[:nf.value], max_pending=pend[:, 0], median_lo=pend[:, 1], median_hi=pend[:, 2])

