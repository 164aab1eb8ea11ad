"""Randomised sweep: oracle (oracle/cpu_sim.c, fifo + yarn tick loop) vs the unmodified reference run live (`python run_sim.py`).
Container only (needs /root/reference).

    python scripts/sweep_fifo_vs_ref.py [first_seed] [n_cases] [workers]

Every case = random trace (20 .. 90 jobs), random cluster shape / cpu / memory / GPU-memory capacity (the generator of
tests/test_oracle_vs_live_reference.py); compares job.csv and cluster.csv (minus the unseeded-RNG column) byte for byte."""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import ref_runner  # noqa: E402
import test_oracle_vs_live_reference as live  # noqa: E402


def one(seed):
    try:
        seed, ref, job, clu = live._run(seed)
        if ref['job_csv'] is None:
            return seed, None, ref['stderr'][-200:]
        return seed, job == ref['job_csv'] and clu == ref_runner.strip_util_column(ref['cluster_csv']), ''
    except Exception as e:
        return seed, None, repr(e)[:200]


if __name__ == '__main__':
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    ok = bad = err = 0
    with ProcessPoolExecutor(max_workers=workers) as ex:
        for seed, same, msg in ex.map(one, range(first, first + n)):
            if same is None:
                err += 1; print('ERROR', seed, msg, flush=True)
            elif same:
                ok += 1
            else:
                bad += 1; print('MISMATCH', seed, flush=True)
    print('cases %d: equal %d, mismatches %d, errors %d' % (n, ok, bad, err))
