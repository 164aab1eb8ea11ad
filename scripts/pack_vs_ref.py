"""Dev aid: oracle_pack (horus / gandiva restatement) vs the live reference on zero-spread traces."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_runner, tracegen, cpu_sim


def zero_spread(df):
    df = df.copy(); df['gpu_utilization_max'] = df['gpu_utilization_avg']; return df


def compare(df, flags, schedule, num_buffer=5, verbose=True, scheme=None, num_queue=1, inject_seed=None):
    d = tempfile.mkdtemp(); p = os.path.join(d, 't.csv'); tracegen.write(df, p)
    t0 = time.time()
    r = ref_runner.run_reference(p, schedule=schedule, scheme=scheme or schedule, num_buffer=num_buffer, num_queue=num_queue, inject_seed=inject_seed, **flags)
    tr = cpu_sim.prepare_trace(p)
    t1 = time.time()
    try:
        o = cpu_sim.run_pack(cpu_sim.make_cluster(**flags), tr, schedule, num_buffer, scheme=scheme, num_queue=num_queue, inject_seed=inject_seed or 0)
    except RuntimeError as e:
        print('oracle raised', e, '| ref rc', r['returncode'], r['stderr'][-300:]); return r['returncode'] != 0
    t2 = time.time()
    oj = cpu_sim.format_job_csv(tr, o); oc = cpu_sim.format_cluster_csv(o)
    rc = ref_runner.strip_util_column(r['cluster_csv'])
    okj, okc = oj == r['job_csv'], oc == rc
    if verbose:
        print(schedule, 'ref %.1fs oracle %.2fs' % (t1 - t0, t2 - t1), 'job', okj, 'cluster', okc, 'ticks', o['n_ticks'], 'fin', len(o['finish_order']),
              'bumped', int((o['actual_duration'] != tr['duration']).sum()), 'rc', r['returncode'])
    if not okj:
        a, b = oj.split('\r\n'), r['job_csv'].split('\r\n')
        for i, (x, y) in enumerate(zip(a, b)):
            if x != y: print(' job line', i, '\n  oracle', x, '\n  ref   ', y); break
        print(' lens', len(a), len(b))
    if not okc:
        a, b = oc.split('\r\n'), rc.split('\r\n')
        for i, (x, y) in enumerate(zip(a, b)):
            if x != y: print(' cluster line', i, '\n  oracle', x, '\n  ref   ', y); break
        print(' lens', len(a), len(b))
    return okj and okc


if __name__ == '__main__':
    small = dict(num_switch=1, num_node_p_switch=4, num_gpu_p_node=8)
    ok = True
    for sched in ('horus', 'gandiva'):
        ok &= compare(zero_spread(tracegen.frame_probe100()), small, sched)
    print('ALL OK' if ok else 'MISMATCH')
