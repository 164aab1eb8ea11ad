"""Trace and cluster-spec ingestion: CSV -> the flat 32-byte job records of include/rlgs.h.

Host mirror of the reference's ingestion (paths relative to the reference repo):
  JobTraceReader.__init__/prepare_jobs   core/jobs/job_generator.py:169-196   (filter, sort, dropna, shift, /10000)
  JobsManager.gen_jobs -> Job(...)        core/jobs/jobs_manager.py:228-241, core/jobs/job.py:75-110
  Infrastructure._init_from_spec_file    infra/infrastructure.py:71-105
pandas does the filtering / sorting so row order (including how ties are broken) is inherited from
the same library calls the reference makes.  Everything the device compares against is derived
here once, in float64, with the reference's own expressions.
"""
import csv
import math
import os
from dataclasses import dataclass

import numpy as np

from . import _ffi

SPEC_KEYS = ['num_switch', 'num_node_p_switch', 'num_gpu_p_node', 'num_cpu_p_node', 'mem_p_node']


@dataclass
class Cluster:
    num_switch: int = 1
    num_node_p_switch: int = 32
    num_gpu_p_node: int = 8
    num_cpu_p_node: int = 128
    mem_p_node: int = 512
    gpu_memory_capacity: int = 32  # GiB, flag --gpu_memory_capacity

    @property
    def num_nodes(self):
        return self.num_switch * self.num_node_p_switch

    @property
    def num_gpus(self):
        return self.num_nodes * self.num_gpu_p_node

    @property
    def cap_mib(self):
        return self.gpu_memory_capacity * 1024  # infrastructure.py:36

    def to_ffi(self):
        return _ffi.ClusterSpec(self.num_switch, self.num_node_p_switch, self.num_gpu_p_node,
                                self.num_cpu_p_node, self.mem_p_node, 0)


def cluster_from_flags(flags):
    """flags: any object/dict with the run_sim.py cluster flags; honours --cluster_spec like the
    reference (silently ignored when the file is missing or lacks a column, infrastructure.py:73,91-92)."""
    g = (lambda k, d: flags.get(k, d)) if isinstance(flags, dict) else (lambda k, d: getattr(flags, k, d))
    c = Cluster(g('num_switch', 1), g('num_node_p_switch', 32), g('num_gpu_p_node', 8), g('num_cpu_p_node', 128),
                g('mem_p_node', 512), g('gpu_memory_capacity', 32))
    spec = g('cluster_spec', None)
    if spec and os.path.exists(spec):
        with open(spec, 'r') as fh:
            reader = csv.DictReader(fh, delimiter=',')
            if all(k in (reader.fieldnames or []) for k in SPEC_KEYS):
                for row in reader:
                    c.num_switch, c.num_node_p_switch, c.num_gpu_p_node, c.num_cpu_p_node, c.mem_p_node = (
                        int(row[k]) for k in SPEC_KEYS)
    return c


@dataclass
class Trace:
    """Per-job host columns (queue-arrival order) + the packed device records."""
    label: np.ndarray        # original CSV row label -> job_id column
    nt: np.ndarray           # normalized_time in ticks (float64)
    duration: np.ndarray     # minutes * scale_factor (float64)
    used_gpus: np.ndarray    # float64 (num_gpu column prints as float, q7)
    records: np.ndarray      # _ffi.JOB_DTYPE
    mem_shift: int           # mem_term unit = 2**-mem_shift MiB
    cap_mib: int
    model_mb: np.ndarray = None      # network-cost inputs (zeros when the trace has no model_name / iterations)
    iterations: np.ndarray = None
    util_avg: np.ndarray = None      # gpu_utilization_avg / _max and memory_max in MiB: inputs of the pack placement
    util_max: np.ndarray = None
    mem_mib: np.ndarray = None
    mem_avg_mib: np.ndarray = None   # memory_avg in MiB: a k-means feature of horus+ (core/jobs/utils.py:10)

    def pack_inputs(self):
        """Arrays of rlgs_pack_inputs (include/rlgs.h): what horus_score / Device.can_fit read from a Task
        (core/scheduling/horus.py:28-56, infra/device.py:48-77)."""
        mem = self.mem_mib * float(2 ** self.mem_shift)
        if len(mem) and (np.any(mem != np.floor(mem)) or mem.max() >= 2.0 ** 52):
            raise ValueError('memory_max is not an exact multiple of 2**-%d MiB' % self.mem_shift)
        return dict(util_avg=np.ascontiguousarray(self.util_avg, np.float64),
                    util_sd=np.ascontiguousarray((self.util_max - self.util_avg) / 2, np.float64),
                    task_mem=np.ascontiguousarray(mem.astype(np.int64)),
                    heap_cap=np.ascontiguousarray(np.floor(self.used_gpus).astype(np.int32)),
                    util_max=np.ascontiguousarray(self.util_max, np.float64),
                    mem_avg_mib=np.ascontiguousarray(self.mem_avg_mib, np.float64),
                    used_gpus=np.ascontiguousarray(self.used_gpus, np.float64))

    def __len__(self):
        return len(self.records)


def _dyadic_shift(values, limit_bits=62):
    """Smallest s such that every value * 2**s is an integer (values are float64 MiB amounts)."""
    s = 0
    v = np.asarray(values, dtype=np.float64)
    if len(v) == 0:
        return 0
    m, e = np.frexp(v[v != 0]) if np.any(v != 0) else (np.zeros(0), np.zeros(0, int))
    if len(m):
        # m * 2**53 is an integer; count its trailing zero bits
        mi = (np.abs(m) * (1 << 53)).astype(np.int64)
        tz = np.zeros(len(mi), dtype=np.int64)
        x = mi.copy()
        for b in (32, 16, 8, 4, 2, 1):
            z = (x & ((1 << b) - 1)) == 0
            tz += np.where(z, b, 0)
            x = np.where(z, x >> b, x)
        lowest = e.astype(np.int64) - 53 + tz  # exponent of the lowest set bit
        s = int(max(0, -lowest.min()))
    return s


def prepare_trace(trace, cluster, scale_factor=0.5):
    """trace: CSV path or DataFrame in the reference's trace schema."""
    import pandas as pd
    df = pd.read_csv(trace) if isinstance(trace, (str, os.PathLike)) else trace
    df = df[df['type'] == 'noninteractive']
    df = df.sort_values(by='normalized_time')
    df = df.dropna()
    n = len(df)
    nt = df['normalized_time'].to_numpy(dtype=np.float64)
    nt = (nt - (nt.min() if n else 0.0)) / 10000
    duration = df['minutes'].to_numpy(dtype=np.float64) * scale_factor
    used = df['used_gpus'].to_numpy(dtype=np.float64)
    gpc = df['gpu_per_container'].to_numpy()
    mem_mib = df['memory_max'].to_numpy(dtype=np.float64) / 1024 / 1024   # util.convert_bytes(.., "MiB")
    mem_avg_mib = df['memory_avg'].to_numpy(dtype=np.float64) / 1024 / 1024
    ua = df['gpu_utilization_avg'].to_numpy(dtype=np.float64)
    um = df['gpu_utilization_max'].to_numpy(dtype=np.float64)
    if n and (np.any(gpc != np.floor(gpc)) or np.any(gpc < 1)):
        raise ValueError('gpu_per_container must be a positive integer (the reference divides by it, job.py:96)')
    gpc = gpc.astype(np.int64)
    tasks = np.floor_divide(used, gpc.astype(np.float64)).astype(np.int64)      # job.py:96-99
    if n and np.any(tasks < 1):
        raise ValueError('a job has used_gpus < gpu_per_container: zero tasks (the reference raises StopIteration, node.py:118)')
    if n and (not np.all(np.isfinite(duration)) or not np.all(np.isfinite(nt))):
        raise ValueError('non-finite minutes / normalized_time')
    cap = cluster.cap_mib
    rec = np.zeros(n, dtype=_ffi.JOB_DTYPE)
    rec['arrival_tick'] = np.ceil(nt).astype(np.int64)                           # first d with nt <= d
    rec['dur_ticks'] = np.maximum(1, np.ceil(duration)).astype(np.int64)         # cf5
    gpus_cmp = np.ceil(used).astype(np.int64)
    least = np.ceil(used / cluster.num_gpu_p_node).astype(np.int64)              # algorithm.py:310
    fits = (cap - (0 + mem_mib)) > 500                                           # device.py:67-77 on an empty device
    for name, arr, lim in (('gpus', gpus_cmp, 65535), ('tasks', tasks, 32767), ('gpus_per_task', gpc, 65535), ('least', least, 32767)):
        if n and arr.max() > lim:
            raise ValueError('%s out of range for the device record' % name)
    rec['gpus'] = gpus_cmp
    rec['tasks'] = tasks
    rec['gpus_per_task'] = gpc
    rec['least_nodes_fits'] = (least | (fits.astype(np.int64) << 15)).astype(np.uint16)
    # avg_gpu_memory_allocated numerator: sum over busy devices of min(cap, mem_max MiB).  Every term
    # is a dyadic rational, so with a common unit 2**-s MiB the float64 running sum of the reference
    # (schedule.py:109-121) is exact and equals an integer sum as long as it stays below 2**53.
    term = np.where(mem_mib < cap, mem_mib, float(cap))
    shift = max(_dyadic_shift(term), _dyadic_shift(mem_mib))   # the pack placement also adds un-clamped amounts
    scaled = term * float(2 ** shift)
    if n and (shift > 60 or scaled.max() * max(cluster.num_gpus, 1) >= 2.0 ** 53):
        raise ValueError('memory_max values are not exactly summable in 53 bits; cannot reproduce '
                         'avg_gpu_memory_allocated bit-exactly')
    # stored as the job's total over the devices it occupies (tasks * gpus_per_task): the kernel only adds / subtracts it
    rec['mem_term'] = scaled.astype(np.int64) * (tasks * gpc)
    rec['util_mu_q'] = np.clip(np.rint(np.minimum(ua, 100.0) * 512), 0, 65535).astype(np.uint16)
    rec['util_sd_q'] = np.clip(np.rint(np.maximum(um - ua, 0.0) / 2 * 512), 0, 65535).astype(np.uint16)
    rec['index'] = np.arange(n)
    from .model_factory import size_mb
    model_mb = (np.array([size_mb(m) for m in df['model_name']], dtype=np.float64) if 'model_name' in df.columns
                else np.zeros(n, dtype=np.float64))
    iters = df['iterations'].to_numpy(dtype=np.float64) if 'iterations' in df.columns else np.zeros(n, dtype=np.float64)
    return Trace(label=df.index.to_numpy().astype(np.int64), nt=nt, duration=np.ascontiguousarray(duration), used_gpus=used,
                 records=np.ascontiguousarray(rec), mem_shift=shift, cap_mib=cap,
                 model_mb=np.ascontiguousarray(model_mb), iterations=np.ascontiguousarray(iters),
                 util_avg=np.ascontiguousarray(ua), util_max=np.ascontiguousarray(um), mem_mib=np.ascontiguousarray(mem_mib),
                 mem_avg_mib=np.ascontiguousarray(mem_avg_mib))
