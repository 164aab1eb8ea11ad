"""Deterministic synthetic Philly-style trace generators (the bench workload and the test traces).

Writes CSVs in the reference's trace schema (the columns read by
/root/reference/core/jobs/job_generator.py:175-189 and
/root/reference/core/jobs/jobs_manager.py:234-238):
gpu_per_container, gpu_utilization_avg, gpu_utilization_max, memory_avg,
memory_max, minutes, model, type, used_gpus, normalized_time.

`probe100` and `gen` follow SURVEY.md Appendix D draw-for-draw (numpy PCG64 via
default_rng), so the Appendix C.2 hashes can be re-derived.  `frame_*` return
DataFrames; `write` stores them.
"""
import numpy as np
import pandas as pd

COLS = ['gpu_per_container', 'gpu_utilization_avg', 'gpu_utilization_max', 'memory_avg',
        'memory_max', 'minutes', 'model', 'type', 'used_gpus', 'normalized_time']


def frame_probe100():
    """BASELINE config C1 trace (100 jobs, 1 switch x 4 nodes x 8 GPUs)."""
    rng = np.random.default_rng(0)
    n = 100
    df = pd.DataFrame({
        'gpu_per_container': rng.choice([1, 2, 4, 8], n),
        'gpu_utilization_avg': rng.uniform(5, 90, n),
        'gpu_utilization_max': 0,
        'memory_avg': rng.uniform(1e9, 1.2e10, n),
        'memory_max': 0,
        'minutes': rng.uniform(5, 400, n),
        'model': 'V100', 'type': 'noninteractive',
        'used_gpus': 0.0, 'normalized_time': np.sort(rng.uniform(0, 2e6, n))})
    df['gpu_utilization_max'] = np.minimum(100, df.gpu_utilization_avg + rng.uniform(0, 20, n))
    df['memory_max'] = (df.memory_avg * rng.uniform(1, 1.3, n)).astype('int64')
    df['used_gpus'] = (df.gpu_per_container * rng.choice([1, 1, 1, 2], n)).astype(float)
    return df


def frame_gen(n, seed, span_ticks):
    """Philly-style trace: probe2k = gen(2000,1,2000), probe10k = gen(10000,2,10000),
    probe60k = gen(60000,3,60000) (the north-star trace)."""
    rng = np.random.default_rng(seed)
    g = rng.choice([1, 2, 4, 8, 16, 32], n, p=[0.70, 0.10, 0.10, 0.07, 0.02, 0.01])
    gpc = np.minimum(g, rng.choice([1, 2, 4, 8], n, p=[0.6, 0.15, 0.15, 0.1]))
    gpc = np.where(g % gpc == 0, gpc, 1)
    ua = rng.uniform(1, 95, n)
    ma = rng.uniform(5e8, 1.4e10, n)
    df = pd.DataFrame({
        'gpu_per_container': gpc, 'gpu_utilization_avg': ua,
        'gpu_utilization_max': np.minimum(100, ua + rng.uniform(0, 30, n)),
        'memory_avg': ma, 'memory_max': (ma * rng.uniform(1, 1.2, n)).astype('int64'),
        'minutes': np.exp(rng.normal(4.0, 1.5, n)).clip(1, 20000),
        'model': 'V100', 'type': 'noninteractive', 'used_gpus': g.astype(float),
        'normalized_time': np.sort(rng.uniform(0, span_ticks * 1e4, n))})
    return df


def frame_rows(rows):
    """Small hand-written traces: rows = list of dicts with any subset of COLS."""
    base = dict(gpu_per_container=1, gpu_utilization_avg=50.0, gpu_utilization_max=60.0,
                memory_avg=4e9, memory_max=5000000000, minutes=4.0, model='V100',
                type='noninteractive', used_gpus=1.0, normalized_time=0.0)
    out = []
    for r in rows:
        d = dict(base)
        d.update(r)
        out.append(d)
    return pd.DataFrame(out, columns=COLS)


def write(df, fn):
    df.to_csv(fn, index=False)
    return fn
