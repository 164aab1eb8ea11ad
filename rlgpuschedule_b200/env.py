"""Vectorised RL environment over the device-resident simulator.

Replaces the reference's stub (model/env.py:1-6: `Environment.__init__`, `step(action)` both `pass`;
`Scheduler.agent = None  # TODO: RL agent`, core/scheduling/schedule.py:25-27).  The semantics are
build-defined (DESIGN.md "Environment"): one step = one scheduler tick of every replica; the action is the
index, inside the k-job look-ahead window of the queue (k = --num_buffer, cf. schedule_horus
algorithm.py:204-240), of the job that gets this tick's placement attempt, or -1 for none.
Observations, rewards and done flags are torch CUDA tensors written directly by the kernel.
"""
import ctypes as C

import numpy as np

from . import _ffi
from .simulator import Simulator


class Environment(object):
    def __init__(self, cluster=None, traces=None, n_replicas=1, window_k=5, device=0, slot_cap=0, seed=0, rows=False):
        """traces: one Trace (shared by all replicas) or a list of (Trace, first_replica, count).
        rows=True also records one cluster.csv row per stepped tick (one replica per warp, 64-byte rows on the device)."""
        import torch
        self.torch = torch
        self.window_k = int(window_k)
        self.n_replicas = int(n_replicas)
        self.seed = int(seed)
        self.device = torch.device('cuda', device)
        kw = dict(rows='device', rows_format='wide', lanes_per_replica=32) if rows else dict(rows=False)
        self.sim = Simulator(cluster, 'fifo', 'yarn', n_replicas=n_replicas, device=device, slot_cap=slot_cap, **kw)
        if traces is not None:
            if isinstance(traces, (list, tuple)):
                for tr, first, count in traces:
                    self.sim.load_trace(tr, first, count)
            else:
                self.sim.load_trace(traces)
        dim = C.c_int32(0)
        _ffi.check(_ffi.lib().rlgs_env_obs_dim(self.sim._h, self.window_k, C.byref(dim)))
        self.obs_dim = dim.value
        with torch.cuda.device(self.device):
            self.obs = torch.zeros(self.n_replicas, self.obs_dim, dtype=torch.float32, device=self.device)
            self.reward = torch.zeros(self.n_replicas, dtype=torch.float32, device=self.device)
            self.done = torch.zeros(self.n_replicas, dtype=torch.uint8, device=self.device)
        self._noop = torch.full((self.n_replicas,), -1, dtype=torch.int32, device=self.device)

    def _bind_stream(self):
        # kernels are enqueued on torch's current stream so they order with the policy network's work
        _ffi.check(_ffi.lib().rlgs_set_stream(self.sim._h, C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream), 1))

    def reset(self):
        self._bind_stream()
        _ffi.check(_ffi.lib().rlgs_env_reset(self.sim._h))
        # the first observation is written by the kernel too (a zero-tick launch): the layout has one author
        _ffi.check(_ffi.lib().rlgs_env_observe(self.sim._h, self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(), self.window_k))
        return self.obs

    def step(self, action):
        """action: int32 tensor [n_replicas] on the device (or anything torch.as_tensor accepts).
        Returns (obs, reward, done, info) — the tensors are reused between calls."""
        torch = self.torch
        a = torch.as_tensor(action, dtype=torch.int32, device=self.device).contiguous()
        self._bind_stream()
        _ffi.check(_ffi.lib().rlgs_env_step(self.sim._h, a.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(),
                                            self.done.data_ptr(), 2, self.window_k, self.seed, 1))
        return self.obs, self.reward, self.done, {}

    def rollout(self, policy='random', n_ticks=1 << 30):
        """Runs up to n_ticks ticks of every replica on the device with a built-in policy
        ('head' = fifo, 'random' = uniform pick inside the window from a counter-based RNG)."""
        self._bind_stream()
        pol = {'head': 0, 'random': 1}[policy]
        _ffi.check(_ffi.lib().rlgs_env_step(self.sim._h, None, self.obs.data_ptr(), self.reward.data_ptr(),
                                            self.done.data_ptr(), pol, self.window_k, self.seed, int(min(n_ticks, 1 << 30))))
        return self.obs, self.reward, self.done, {}

    def sync(self):
        rc = _ffi.lib().rlgs_env_sync(self.sim._h)
        if rc == _ffi.ERR_SLOTS:
            raise _ffi.RlgsError(rc, _ffi.lib().rlgs_last_error().decode(errors='replace') +
                                 ' - more jobs ran at once than on-chip slots: create the Environment with a larger slot_cap, or use run_episodes()')
        _ffi.check(rc)
        return self

    def run_episodes(self, policy='random'):
        """reset() + rollout(policy) to the end of every episode + sync().  A policy that packs more jobs at once than the on-chip
        slot table holds (RLGS_ERR_SLOTS) restarts the batch with a table twice as large, like Simulator.run() does."""
        while True:
            self.reset()
            self.rollout(policy)
            rc = _ffi.lib().rlgs_env_sync(self.sim._h)
            if rc == _ffi.ERR_SLOTS:
                cap = max(64, 2 * (self.sim._slot_cap or 128))
                if cap > 2 * max(self.sim.cluster.num_gpus, 32):
                    _ffi.check(rc)
                self.sim._rebuild(slot_cap=cap)
                continue
            _ffi.check(rc)
            return self

    def returns(self):
        """Episode returns so far, -(sum of job completion times) per replica (host numpy int64)."""
        self.sync()
        return self.sim.returns()

    def returns_tensor(self):
        """Zero-copy int64 CUDA tensor over the kernel-written return buffer (the all-gather send buffer)."""
        class _Buf(object):
            pass
        b = _Buf()
        b.__cuda_array_interface__ = {'shape': (self.n_replicas,), 'typestr': '<i8',
                                      'data': (self.sim.returns_device_ptr(), False), 'version': 3}
        return self.torch.as_tensor(b, device=self.device)

    def close(self):
        self.sim.close()
