"""Host-callable scheduling plugins over the device-resident simulator.

The reference dispatches one scheduling attempt per tick through a dict of Python callables
(core/scheduling/schedule.py:40-60):

    nodes, job, success = scheduling_algorithms[schedule](scheme, placement_algorithms[scheme], infrastructure,
                                                          jobs_manager, delta, k=num_buffer)

The built-in keys run entirely on the GPU (algorithm.DevicePolicy).  A key the USER registers with a plain Python
callable of that signature is executed here: the tick loop stays on the device (the environment instantiation of the fifo
kernel, one `rlgs_env_step` per tick), and every tick the callable is handed read-only views of the device state

    infrastructure.nodes["1".."N"]  ->  NodeView   (cpu_free(), mem_free(), get_free_devices(), is_free(), rack_id)
    jobs_manager.get_next_job(delta), jobs_manager.window(k)  ->  JobView  (job_id, gpus, task_count, gpu_per_worker,
                                                                            duration, pending_time, ...)

built lazily from the observation the kernel wrote for that tick.  The job it returns (with success true) becomes the
tick's action: the device makes the placement attempt for exactly that job with the yarn fit, starts it if it fits, and
advances the tick.  `placement_algo` is a dry run of ms_yarn_placement on the same node counters (algorithm.yarn_preview):
it has no side effect; the device's own fit in the same tick is authoritative.
"""
import collections
import math

import numpy as np

CPUS_PER_TASK = 12   # core/jobs/job.py:105
MEM_PER_TASK = 60    # core/jobs/job.py:106


class DeviceView(object):
    """Placeholder for an idle device (Node.get_free_devices() returns device objects in the reference)."""
    __slots__ = ('node_id', 'slot')

    def __init__(self, node_id, slot):
        self.node_id, self.slot = node_id, slot


class NodeView(object):
    """Read-only view of one node of the simulated cluster (infra/node.py:7-60 of the reference)."""

    def __init__(self, node_id, rack_id, gpu_count, free_gpus, cpu_free, mem_free):
        self.node_id, self.rack_id, self.gpu_count = node_id, rack_id, gpu_count
        self._free_gpus, self._cpu_free, self._mem_free = int(free_gpus), int(cpu_free), int(mem_free)

    def cpu_free(self):
        return self._cpu_free

    def mem_free(self):
        return self._mem_free

    def is_free(self):                      # node.py:59-60
        return self._cpu_free > 0 or self._mem_free > 0

    def get_free_devices(self):             # node.py:99-107 (yarn: devices without a task)
        return [DeviceView(self.node_id, i) for i in range(self._free_gpus)]

    def can_fit_num_task(self, job):        # node.py:109-127
        return min(self._free_gpus // max(1, int(job.gpu_per_worker)), self._cpu_free // CPUS_PER_TASK, self._mem_free // MEM_PER_TASK)

    def __repr__(self):
        return 'NodeView(%s: %d free gpus, cpu %d, mem %d)' % (self.node_id, self._free_gpus, self._cpu_free, self._mem_free)


class JobView(object):
    """Read-only view of a queued job inside the look-ahead window (core/jobs/job.py:60-110)."""

    def __init__(self, trace, trace_index, window_index, gpus, tasks, dur_ticks, pending):
        self.trace_index, self.window_index = int(trace_index), int(window_index)
        self.job_id = str(int(trace.label[self.trace_index]))
        self.gpus = float(trace.used_gpus[self.trace_index])
        self.task_count = int(tasks)
        self.gpu_per_worker = int(trace.records['gpus_per_task'][self.trace_index])
        self.duration = float(trace.duration[self.trace_index])
        self.dur_ticks = int(dur_ticks)
        self.pending_time = float(pending)
        self.submit_time = int(trace.nt[self.trace_index])
        self.gpus_ceil = int(gpus)

    def is_waiting(self):
        return True

    def __repr__(self):
        return 'JobView(%s: %g gpus, %d tasks, pending %g)' % (self.job_id, self.gpus, self.task_count, self.pending_time)


class InfrastructureView(object):
    """infra/infrastructure.py:16-44: `nodes` is an OrderedDict node id -> NodeView, built on first access."""

    def __init__(self, cluster, flags, obs):
        self.flags = flags
        self.cluster = cluster
        self.num_switch, self.num_nodes_p_switch = cluster.num_switch, cluster.num_node_p_switch
        self.num_gpu_p_node, self.num_cpu_p_node, self.mem_p_node = cluster.num_gpu_p_node, cluster.num_cpu_p_node, cluster.mem_p_node
        self._obs = obs
        self._nodes = None

    @property
    def nodes(self):
        if self._nodes is None:
            N = self.cluster.num_nodes
            o = self._obs
            self._nodes = collections.OrderedDict(
                (str(i + 1), NodeView(str(i + 1), str(i // self.num_nodes_p_switch), self.num_gpu_p_node, o[i], o[N + i], o[2 * N + i]))
                for i in range(N))
        return self._nodes

    def get_free_nodes(self):
        return [n for n in self.nodes.values() if n.is_free()]

    def num_free_nodes(self):
        N = self.cluster.num_nodes
        return int(((self._obs[N:2 * N] > 0) | (self._obs[2 * N:3 * N] > 0)).sum())

    def get_total_gpus(self):
        return self.cluster.num_gpus


class JobsManagerView(object):
    """core/jobs/jobs_manager.py: what a scheduling callable reads; pop() only records the caller's commitment."""

    def __init__(self, trace, obs, n_nodes, window_k):
        base = 3 * n_nodes
        t = obs[base + 5 * window_k: base + 5 * window_k + 4]
        self._queued, self._running, self._finished, self.delta = int(t[0]), int(t[1]), int(t[2]), int(t[3])
        w = obs[base: base + 5 * window_k].reshape(window_k, 5)
        n = min(self._queued, window_k)
        seen = [JobView(trace, w[i, 4], i, w[i, 0], w[i, 1], w[i, 2], w[i, 3]) for i in range(n)]
        # The observation was written at the end of the previous tick; the reference calls the scheduling callable after
        # this tick's arrivals joined the FRONT of the queue (schedule.py:187-190, q1), and the kernel applies the returned
        # index after pushing them too.  Those jobs are known from the trace: every not yet arrived job with arrival_tick <= delta.
        rec = trace.records
        arrived = self._queued + self._running + self._finished
        hi = arrived
        while hi < len(rec) and rec['arrival_tick'][hi] <= self.delta:
            hi += 1
        fresh = [JobView(trace, i, 0, rec['gpus'][i], rec['tasks'][i], rec['dur_ticks'][i], 0) for i in range(arrived, min(hi, arrived + window_k))]
        self._queued += hi - arrived
        self._window = (fresh + seen)[:window_k]
        for i, j in enumerate(self._window):
            j.window_index = i
        self.popped = None

    def window(self, k=None):
        return list(self._window if k is None else self._window[:k])

    def get_next_job(self, delta=None):
        return self._window[0] if self._window else None

    def pop(self, delta=None):
        self.popped = self._window[0] if self._window else None
        return self.popped

    def queuing_jobs(self, delta=None):
        return self._queued

    def total_jobs(self, delta=None):
        return self._queued

    def num_running_jobs(self):
        return self._running

    def total_finished_jobs(self):
        return self._finished


def yarn_preview(infrastructure, job, scheme=None):
    """Dry run of ms_yarn_placement (core/scheduling/algorithm.py:28-32,301-417) on a view: (nodes dict | None, success).
    Counts only (devices, cpu, mem); the device-memory rule of infra/device.py:67-77 is applied by the device."""
    need_g, T, gpc = int(math.ceil(job.gpus)), job.task_count, max(1, job.gpu_per_worker)
    if need_g <= infrastructure.num_gpu_p_node:
        for nid, n in infrastructure.nodes.items():
            if n._free_gpus >= need_g and n._cpu_free >= CPUS_PER_TASK * T and n._mem_free >= MEM_PER_TASK * T:
                return {nid: n}, True
        return None, False
    remaining, used = T, collections.OrderedDict()
    for nid, n in infrastructure.nodes.items():
        cap = max(0, n.can_fit_num_task(job))
        take = min(cap, remaining)
        if take > 0:
            used[nid] = n
            remaining -= take
        if remaining == 0:
            break
    least = int(math.ceil(job.gpus / infrastructure.num_gpu_p_node))
    if remaining > 0 or len(used) < least:
        return None, False
    return used, True


def schedule_fifo_host(scheme, placement_algo, infrastructure, jobs_manager, delta, **kwargs):
    """schedule_fifo (core/scheduling/algorithm.py:189-202) over the views: the host form of the built-in 'fifo' entry."""
    next_job = jobs_manager.get_next_job(delta)
    if next_job is None:
        return None, None, False
    nodes, success = placement_algo(infrastructure, next_job, scheme)
    if success:
        jobs_manager.pop(delta)
    return nodes, next_job, success


def run_host_policy(fn, cluster, trace, flags, scheme='yarn', k=5, device=0, max_ticks=None, rows=True):
    """Scheduler.start() for a user-registered scheduling callable.  Returns (Environment, n_ticks, action tape); rows / job
    tables are read from env.sim afterwards."""
    import torch
    from .env import Environment
    from . import algorithm
    placement = algorithm.placement_algorithms[scheme]
    if not (isinstance(placement, algorithm.DevicePolicy) and placement.name == 'yarn'):
        raise NotImplementedError('host scheduling plugins run over the device yarn placement (got scheme %r)' % (scheme,))
    k = max(1, min(int(k), 32))
    env = Environment(cluster, trace, n_replicas=1, window_k=k, device=device, rows=rows)
    obs = env.reset()
    act = torch.full((1,), -1, dtype=torch.int32, device=env.device)
    N = cluster.num_nodes
    ticks = 0
    tape = []
    while True:
        o = obs[0].cpu().numpy()
        jm = JobsManagerView(trace, o, N, k)
        action = -1
        if jm.queuing_jobs() > 0:
            infra = InfrastructureView(cluster, flags, o)
            if infra.num_free_nodes() >= 1:                       # schedule.py:40-44
                nodes, job, success = fn(scheme, placement, infra, jm, jm.delta, k=k)
                if success and job is not None:
                    if not isinstance(job, JobView):
                        raise TypeError('a scheduling plugin must return one of the JobView objects it was given')
                    action = job.window_index
        act.fill_(action)
        tape.append(action)
        obs, _, done, _ = env.step(act)
        ticks += 1
        if bool(done[0].item()) or (max_ticks is not None and ticks >= max_ticks):
            break
    env.sync()
    return env, ticks, np.asarray(tape, np.int32)
