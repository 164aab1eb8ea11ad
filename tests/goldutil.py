"""Helpers to load tests/golden fixtures (written by oracle/make_golden.py from the real reference)."""
import gzip
import hashlib
import json
import os

import golden_cases
import tracegen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def sha(s):
    return hashlib.sha256(s.encode() if isinstance(s, str) else s).hexdigest()


def case_names(kind='small', schedule='fifo'):
    out = []
    for n, c in golden_cases.CASES.items():
        if not os.path.exists(os.path.join(GOLD, n, 'meta.json')) or c.get('schedule', 'fifo') != schedule:
            continue
        k = 'huge' if c.get('huge') else ('big' if c.get('big') else 'small')
        if kind == 'all' or k == kind:
            out.append(n)
    return out


def load(name):
    d = os.path.join(GOLD, name)
    meta = json.load(open(os.path.join(d, 'meta.json')))
    case = golden_cases.CASES[name]
    flags = dict(case['flags'])
    for k, v in list(flags.items()):
        if isinstance(v, str) and v.startswith('@'):
            flags[k] = os.path.join(ROOT, v[1:])
    out = dict(meta=meta, flags=flags, job=None, cluster=None, schedule=case.get('schedule', 'fifo'), num_buffer=case.get('num_buffer', 5),
               scheme=case.get('scheme', case.get('schedule', 'yarn') if case.get('schedule', 'fifo') != 'fifo' else 'yarn'),
               num_queue=case.get('num_queue', 1), inject_seed=case.get('inject_seed', 0))
    if os.path.exists(os.path.join(d, 'trace.csv')):
        out['trace'] = os.path.join(d, 'trace.csv')
        out['frame'] = None
    else:
        out['trace'] = None
        out['frame'] = case['frame']()
    out['queue_limit'] = tuple(case.get('queue_limit', ()))
    for key, fn in (('job', 'job.csv'), ('cluster', 'cluster_noutil.csv' if out['schedule'] not in golden_cases.LEGACY else 'cluster.csv')):
        p = os.path.join(d, fn)
        if os.path.exists(p):
            out[key] = open(p, newline='').read()
        elif os.path.exists(p + '.gz'):
            out[key] = gzip.open(p + '.gz', 'rb').read().decode()
    return out


def trace_input(g):
    """CSV path when stored, else the regenerated DataFrame round-tripped through CSV text
    (pandas parses floats from text, so the round trip matters for bit-exactness)."""
    import io
    import pandas as pd
    if g['trace']:
        return g['trace']
    buf = io.StringIO()
    g['frame'].to_csv(buf, index=False)
    txt = buf.getvalue()
    assert sha(txt) == g['meta']['trace_sha256'], 'tracegen drifted from the fixture'
    return pd.read_csv(io.StringIO(txt))


def pack_case_names(kinds=('small', 'big', 'huge')):
    """Fixtures of the pack family (horus, horus+, gandiva; over the pack placement or yarn)."""
    return [n for sched in ('horus', 'gandiva', 'horus+') for kind in kinds for n in case_names(kind, sched)]


def legacy_case_names(kinds=('small', 'big', 'huge')):
    """Fixtures of the legacy event loops (sjf family over yarn, dlas family with count admission)."""
    return [n for sched in golden_cases.LEGACY for kind in kinds for n in case_names(kind, sched)]
