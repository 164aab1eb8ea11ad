"""Scheduling / placement / post-tick plugin registries with the reference's surface
(core/scheduling/algorithm.py of the reference):

    scheduling_algorithms[name](scheme, placement_algo, infrastructure, jobs_manager, delta, **kwargs{k})
        -> (nodes | None, job | None, success | None)                       algorithm.py:189-298, call site schedule.py:45-47
    placement_algorithms[name](infrastructure, next_job, scheme) -> (nodes, success)      algorithm.py:28-32,182-187
    plugin_algorithms[name](infrastructure, jobs_manager)                   algorithm.py:420-444, call site schedule.py:200-202
    score_fn[name](node, task)                                              algorithm.py:9-13

The entries the hot path implements are DevicePolicy objects: Scheduler.start() hands their integer
id to librlgs and the whole tick / event loop runs on the GPU.  `horus` (schedule_horus + horus_placement
with horus_score), `horus+` (k-means queues + credit pick) and `gandiva` (schedule_fifo + gandiva_score + the
time-slice plugin) are among them, so every key of the reference's tables runs on the device.

Users may register their own scheduling entries, plain callables with the reference's signature:

    def mine(scheme, placement_algo, infrastructure, jobs_manager, delta, **kwargs): ... return nodes, job, success
    algorithm.scheduling_algorithms['mine'] = mine            # then: run_sim.py --schedule mine --scheme yarn

Scheduler.start() then keeps the tick loop on the device and calls `mine` once per tick with read-only views of the device
state (rlgpuschedule_b200/plugin.py); the job it returns gets that tick's placement attempt.  The built-in 'fifo' and 'yarn'
entries are callable over the same views (schedule_fifo and a dry run of ms_yarn_placement), so a plugin can compose them.
"""
from . import _ffi
from . import plugin


class DevicePolicy(object):
    """A policy executed inside the CUDA kernels (rlgpuschedule_b200/csrc)."""

    def __init__(self, name, kind, device_id, reference, host_fn=None):
        self.name, self.kind, self.device_id, self.reference, self.host_fn = name, kind, device_id, reference, host_fn

    def __call__(self, *args, **kwargs):
        """Per-call form over the read-only views of plugin.py (what a user-registered scheduling callable composes with)."""
        if self.host_fn is None:
            raise RuntimeError('%s policy %r runs on the device as part of Scheduler.start(); it has no per-call '
                               'host form' % (self.kind, self.name))
        return self.host_fn(*args, **kwargs)

    def __repr__(self):
        return 'DevicePolicy(%s %r, id %d, restates %s)' % (self.kind, self.name, self.device_id, self.reference)


class HostOnlyPolicy(object):
    """A reference policy that has no device implementation yet ('what comes next')."""

    def __init__(self, name, kind, reference):
        self.name, self.kind, self.reference = name, kind, reference

    def __call__(self, *args, **kwargs):
        raise NotImplementedError('%s policy %r (%s) is not implemented by the device path' % (self.kind, self.name, self.reference))

    def __repr__(self):
        return 'HostOnlyPolicy(%s %r)' % (self.kind, self.name)


scheduling_algorithms = {
    'fifo': DevicePolicy('fifo', 'schedule', _ffi.SCHED['fifo'], 'core/scheduling/algorithm.py:189-202', host_fn=plugin.schedule_fifo_host),
    'sjf': DevicePolicy('sjf', 'schedule', _ffi.SCHED['sjf'], 'run_sim.py:162-287 (dead code, restated)'),
    'dlas-gpu': DevicePolicy('dlas-gpu', 'schedule', _ffi.SCHED['dlas-gpu'], 'run_sim.py:664-947 (dead code, restated)'),
    'dlas': DevicePolicy('dlas', 'schedule', _ffi.SCHED['dlas'], 'run_sim.py:664-947 with gputime=False (dead code, restated)'),
    'shortest': DevicePolicy('shortest', 'schedule', _ffi.SCHED['shortest'], 'run_sim.py:299-431 (dead code, restated)'),
    'shortest-gpu': DevicePolicy('shortest-gpu', 'schedule', _ffi.SCHED['shortest-gpu'], 'run_sim.py:299-431 with gputime (dead code, restated)'),
    'horus': DevicePolicy('horus', 'schedule', _ffi.SCHED['horus'], 'core/scheduling/algorithm.py:204-240'),
    'horus+': DevicePolicy('horus+', 'schedule', _ffi.SCHED['horus+'], 'core/scheduling/algorithm.py:242-290 + core/jobs/utils.py:36-67 (k-means draws are counter-based)'),
    'gandiva': DevicePolicy('gandiva', 'schedule', _ffi.SCHED['gandiva'], 'core/scheduling/algorithm.py:292-298 (schedule_fifo) + time_slice_check :420-440'),
}

placement_algorithms = {
    'yarn': DevicePolicy('yarn', 'placement', _ffi.PLACE['yarn'], 'core/scheduling/algorithm.py:28-32,301-417', host_fn=plugin.yarn_preview),
    'count': DevicePolicy('count', 'placement', _ffi.PLACE['count'], 'run_sim.py:808-823 (free_gpu counting)'),
    # one function under three names in the reference (algorithm.py:182-187); its score table is keyed by the SCHEDULE
    # name (schedule.py:47), so all three behave the same under --schedule horus
    'horus': DevicePolicy('horus', 'placement', _ffi.PLACE['horus'], 'core/scheduling/algorithm.py:34-180'),
    'horus+': DevicePolicy('horus+', 'placement', _ffi.PLACE['horus+'], 'core/scheduling/algorithm.py:34-180'),
    'gandiva': DevicePolicy('gandiva', 'placement', _ffi.PLACE['gandiva'], 'core/scheduling/algorithm.py:34-180'),
}

plugin_algorithms = {
    'gandiva': DevicePolicy('gandiva', 'post-tick plugin', _ffi.SCHED['gandiva'], 'core/scheduling/algorithm.py:420-440'),
}

score_fn = {
    'horus': DevicePolicy('horus', 'score', 0, 'core/scheduling/horus.py:28-56'),
    'horus+': DevicePolicy('horus+', 'score', 0, 'core/scheduling/horus.py:28-56'),
    'gandiva': DevicePolicy('gandiva', 'score', 1, 'core/scheduling/horus.py:6-25'),
}


def is_host_callable(entry):
    """A user-registered scheduling entry: any callable that is not one of this package's policy objects."""
    return callable(entry) and not isinstance(entry, (DevicePolicy, HostOnlyPolicy))


def resolve(schedule, scheme):
    """(schedule name, scheme name) -> (DevicePolicy, DevicePolicy) or raises like the reference would
    (KeyError for unknown keys, NotImplementedError for host-only ones)."""
    sched = scheduling_algorithms[schedule]
    place = placement_algorithms[scheme]
    if is_host_callable(sched):
        if not (isinstance(place, DevicePolicy) and place.name == 'yarn'):
            raise NotImplementedError('a user-registered scheduling callable runs over the device yarn placement (--scheme yarn)')
        return sched, place
    for p in (sched, place):
        if not isinstance(p, DevicePolicy):
            if isinstance(p, HostOnlyPolicy):
                p()  # raises NotImplementedError with the reference location
            raise NotImplementedError('user-registered placement %r: only scheduling entries may be host callables' % (p,))
    packs = place.device_id == _ffi.PLACE['horus']
    pack_schedules = ('horus', 'horus+', 'gandiva')
    if (packs and schedule not in pack_schedules) or (schedule in pack_schedules and not packs and scheme != 'yarn'):
        # fifo + horus: KeyError 'fifo' in the reference's score table (algorithm.py:58)
        raise NotImplementedError('schedule %r with scheme %r is not implemented by the device path' % (schedule, scheme))
    post = plugin_algorithms.get(schedule, None)
    if post is not None and not isinstance(post, DevicePolicy):
        post()
    return sched, place
