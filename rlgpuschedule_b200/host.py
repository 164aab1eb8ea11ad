"""Host-side mirror of the objects run_sim.py wires together in the reference
(run_sim.py:1716-1738): Infrastructure, JobQueueManager, JobsManager, Scheduler.  They keep the
reference's constructor signatures and public attribute names, but hold no per-job / per-node
Python objects: the state lives on the GPU and Scheduler.start() is one call into librlgs."""
import logging
import os
import sys
import time

from . import algorithm, ingest, log_manager as lm
from .simulator import Simulator


class Infrastructure(object):
    """infra/infrastructure.py:16-44 of the reference: cluster spec from flags or --cluster_spec."""

    def __init__(self, flags):
        self.flags = flags
        self.cluster = ingest.cluster_from_flags(flags)
        c = self.cluster
        self.num_switch = c.num_switch
        self.num_nodes_p_switch = c.num_node_p_switch
        self.num_gpu_p_node = c.num_gpu_p_node
        self.num_cpu_p_node = c.num_cpu_p_node
        self.mem_p_node = c.mem_p_node
        self.gpu_memory_capacity = c.cap_mib
        self.bandwidth = getattr(flags, 'bandwidth', 1250)
        self.internode_latency = getattr(flags, 'internode_latency', 0.015)
        self.enable_network_costs = getattr(flags, 'enable_network_costs', False)
        self.cluster_spec = getattr(flags, 'cluster_spec', None)
        for line in ('num_racks in cluster: %d' % c.num_switch, 'num_node_p_rack in cluster: %d' % c.num_node_p_switch,
                     'num_gpu_p_node in cluster: %d' % c.num_gpu_p_node, 'num_cpu_p_node in cluster: %d' % c.num_cpu_p_node,
                     'mem_p_node in cluster: %d' % c.mem_p_node, 'Total nodes in cluster: %d ' % c.num_nodes,
                     'Total racks in cluster: %d ' % c.num_switch):
            logging.info(line)

    @property
    def num_nodes(self):
        return self.cluster.num_nodes

    @property
    def num_gpus(self):
        return self.cluster.num_gpus

    def get_total_gpus(self):
        return self.cluster.num_gpus


class JobQueueManager(object):
    """core/jobs/job_queue_manager.py:9-20: the queues themselves are device state."""

    def __init__(self, flags, file_path=None):
        self.flags = flags
        self.file_path = file_path
        self.num_queue = getattr(flags, 'num_queue', 1)


class JobsManager(object):
    """core/jobs/jobs_manager.py:13-27: reads the trace (JobTraceReader, job_generator.py:169-196)."""

    def __init__(self, flags, job_queue_manager, cluster=None):
        self.flags = flags
        self.job_queue_manager = job_queue_manager
        if not flags.trace_file:
            raise NotImplementedError('synthetic JobGenerator traces are not wired to the simulator (jobs_manager.py:209-212)')
        if not os.path.exists(flags.trace_file):
            logging.error('file: %s not exist' % flags.trace_file)
            sys.exit(1)
        self.replay_trace = True
        self.cluster = cluster or ingest.cluster_from_flags(flags)
        self.trace = ingest.prepare_trace(flags.trace_file, self.cluster)

    def remaining_jobs(self, delta_time=None):
        return len(self.trace)


def parse_queue_limit(text, default=(30, 60, 150)):
    if text is None or text == '':
        return tuple(default)
    return tuple(int(x) for x in str(text).replace(' ', '').split(',') if x)


class Scheduler(object):
    """core/scheduling/schedule.py:12-28,178-216: start() runs the whole simulation and writes the logs."""

    def __init__(self, infrastructure, jobs_manager, log_manager, enable_migration=False):
        self.infrastructure = infrastructure
        self.jobs_manager = jobs_manager
        self.log_manager = log_manager
        self.placement = infrastructure.flags.scheme
        self.schedule = infrastructure.flags.schedule
        self.enable_migration = enable_migration   # inert in the reference as well
        self.agent = None
        self.simulator = None

    def start(self):
        flags = self.infrastructure.flags
        t0 = time.time()
        sched, place = algorithm.resolve(self.schedule, self.placement)
        if algorithm.is_host_callable(sched):
            return self._start_host_plugin(sched, t0)
        scheme = self.placement
        kw = {}
        if self.schedule in ('dlas-gpu', 'dlas'):
            limits = parse_queue_limit(getattr(flags, 'queue_limit', None))
            kw = dict(num_queue=len(limits) + 1, queue_limit=limits)
            scheme = 'count'   # dlas admits by GPU count (run_sim.py:808-823) whatever --scheme says
        if self.schedule in ('horus', 'horus+', 'gandiva'):
            # horus_score / gandiva_score draw device utilisations from an unseeded normal (infra/device.py:52); here the draw is
            # counter-based: --seed fixes it, --util_mode mean replaces every draw by its mean
            seed = getattr(flags, 'seed', None)
            if seed is None and getattr(flags, 'util_mode', 'sample') != 'mean':
                seed = int.from_bytes(os.urandom(4), 'little')
            kw = dict(num_buffer=int(getattr(flags, 'num_buffer', 5)), pack_seed=seed)
            if self.schedule == 'horus+':   # the k-means init is always drawn (core/jobs/utils.py:39): unseeded in the reference
                if seed is None:
                    seed = int.from_bytes(os.urandom(4), 'little')
                kw.update(num_queue=int(getattr(flags, 'num_queue', 1)), pack_seed=seed,
                          pack_rng=getattr(flags, 'util_mode', 'sample') != 'mean')
        net = bool(getattr(flags, 'enable_network_costs', False))
        if net and self.schedule != 'fifo':
            raise NotImplementedError('--enable_network_costs is implemented for the fifo tick loop only')
        if net:
            kw.update(enable_network_costs=True, bandwidth=self.infrastructure.bandwidth,
                      internode_latency=self.infrastructure.internode_latency)
        cluster = self.infrastructure.cluster
        trace = self.jobs_manager.trace
        sim = Simulator(cluster, self.schedule, scheme, n_replicas=1, rows=True, device=getattr(flags, 'device', 0), **kw)
        self.simulator = sim
        sim.load_trace(trace)
        sim.run()
        took = time.time() - t0
        summ = sim.summary(0)
        legacy = self.schedule not in ('fifo', 'horus', 'horus+', 'gandiva')
        if not legacy:
            util_seed = getattr(flags, 'seed', None)
            if util_seed is None and getattr(flags, 'columnar', False):
                util_seed = int.from_bytes(os.urandom(4), 'little')   # one draw stream for cluster.csv and cluster.parquet
            rows0 = sim.rows(0)
            self.log_manager.write_cluster_rows(rows0, cluster, trace.mem_shift, util_mode=getattr(flags, 'util_mode', 'sample'), seed=util_seed)
            j = sim.jobs(0)
            logging.info('Total Time Taken in seconds: %d' % took)
            extra = {}
            if self.schedule in ('horus', 'horus+', 'gandiva'):
                # Job.get_duration() = original + 5 once a task was de-interfered (jobs_manager.py:184-185); jct = Job.time_processed()
                from . import _ffi
                extra = dict(get_duration=trace.duration + 5.0 * (sim.job_plane(0, _ffi.PLANE_AUX) == 1),
                             jct=sim.job_plane(0, _ffi.PLANE_PREEMPT))
            self.log_manager.jcts((trace, j['finish_order'], j['start'], j['end'], j['preempt'], sim.durations(0) if net else None, extra))
            if getattr(flags, 'columnar', False):
                self.log_manager.write_columnar(rows0, cluster, trace, j, trace.mem_shift, get_duration=extra.get('get_duration'),
                                                jct=extra.get('jct'), util_mode=getattr(flags, 'util_mode', 'sample'), seed=util_seed)
        else:
            from . import _ffi
            j = sim.jobs(0)
            self.log_manager.write_legacy(sim.rows(0), cluster, trace, j, sim.job_plane(0, _ffi.PLANE_AUX),
                                          sim.job_plane(0, _ffi.PLANE_RESUME), count_scheme=(scheme == 'count'))
            logging.info('Total Time Taken in seconds: %d' % took)
        ms, launches = sim.kernel_ms()
        logging.info('device: %d rows, %d jobs finished, %d events, kernel %.3f ms in %d launches' % (
            summ['n_ticks'], summ['n_finished'], summ['events'], ms, launches))
        return summ

    def _start_host_plugin(self, fn, t0):
        """A user-registered scheduling_algorithms entry (a Python callable with the reference's signature,
        core/scheduling/schedule.py:45-47): the tick loop stays on the device, the callable picks the job of each tick."""
        from . import plugin
        flags = self.infrastructure.flags
        cluster, trace = self.infrastructure.cluster, self.jobs_manager.trace
        env, ticks, _tape = plugin.run_host_policy(fn, cluster, trace, flags, scheme=self.placement, k=int(getattr(flags, 'num_buffer', 5)),
                                            device=getattr(flags, 'device', 0))
        sim = env.sim
        self.simulator = sim
        self.log_manager.write_cluster_rows(sim.rows(0), cluster, trace.mem_shift,
                                            util_mode=getattr(flags, 'util_mode', 'sample'), seed=getattr(flags, 'seed', None))
        j = sim.jobs(0)
        logging.info('Total Time Taken in seconds: %d' % (time.time() - t0))
        self.log_manager.jcts((trace, j['finish_order'], j['start'], j['end'], j['preempt'], None, {}))
        summ = sim.summary(0)
        logging.info('device: %d ticks stepped for the host plugin %r, %d jobs finished' % (ticks, self.schedule, summ['n_finished']))
        return summ
