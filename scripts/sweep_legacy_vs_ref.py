"""Randomised sweep: oracle (oracle/cpu_sim.c, legacy section) vs the reference's dead-code event loops run unmodified under the
shim globals (oracle/ref_legacy_runner.py).  Container only (needs /root/reference).

    python scripts/sweep_legacy_vs_ref.py [--big] [first_seed] [n_cases] [workers]

Every case = random trace (20 .. 90 jobs), random cluster shape, one of sjf / shortest / shortest-gpu / dlas-gpu / dlas, 2 .. 4
queues with random limits for the dlas family; compares job.csv and cluster.csv byte for byte."""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import test_oracle_vs_live_reference as live  # noqa: E402  (case generator + runner shared with the test suite)


def big_case(seed):
    """200 .. 600 jobs of the Philly-style generator on 16 .. 64 nodes (queues build up, multi-node jobs, preemption chains)"""
    import numpy as np
    from rlgpuschedule_b200 import synth
    rng = np.random.default_rng(seed)
    n = int(rng.integers(200, 600))
    flags = dict(num_switch=int(rng.integers(1, 5)), num_node_p_switch=int(rng.choice([4, 8, 16])), num_gpu_p_node=int(rng.choice([4, 8])))
    return synth.frame_gen(n, seed, int(n * rng.uniform(0.3, 1.5))), flags


def one(arg):
    try:
        if BIG:
            live._case = big_case
        seed, sched, ref, job, clu = live._run_legacy(arg)
        return arg, job == ref['job_csv'] and clu == ref['cluster_csv'], ''
    except Exception as e:  # the oracle or the reference refused the input
        return arg, None, repr(e)[:200]


BIG = '--big' in sys.argv
if BIG:
    sys.argv.remove('--big')

if __name__ == '__main__':
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    args = [(first + i, live.LEGACY[i % len(live.LEGACY)]) for i in range(n)]
    ok = bad = err = 0
    with ProcessPoolExecutor(max_workers=workers) as ex:
        for arg, same, msg in ex.map(one, args):
            if same is None:
                err += 1; print('ERROR', arg, msg, flush=True)
            elif same:
                ok += 1
            else:
                bad += 1; print('MISMATCH', arg, flush=True)
    print('cases %d: equal %d, mismatches %d, errors %d' % (n, ok, bad, err))
