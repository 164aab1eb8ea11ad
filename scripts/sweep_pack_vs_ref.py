"""Dev aid (container only): random differential sweep of oracle_pack against the live reference, all six schedule x placement
combinations (generator and comparison of tests/test_oracle_vs_live_reference.py).  Round 1: 312 cases, 0 mismatches."""
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/oracle')
import test_oracle_vs_live_reference as T
import ref_runner
from concurrent.futures import ThreadPoolExecutor
args = [(5000 + 41 * i + j, combo) for i, combo in enumerate(T.PACK_COMBOS) for j in range(40)]
bad = 0
with ThreadPoolExecutor(max_workers=7) as ex:
    for seed, sched, scheme, ref, job, clu in ex.map(T._run_pack, args):
        if job is None:
            ok = ref['returncode'] != 0 or ref['job_csv'] is None or 'Error' in ref['stderr']
            print(seed, sched, scheme, 'oracle-raises', 'ref-rc', ref['returncode'], 'OK' if ok else 'MISMATCH', flush=True)
            bad += not ok
            continue
        ok = ref['job_csv'] is not None and job == ref['job_csv'] and clu == ref_runner.strip_util_column(ref['cluster_csv'])
        if not ok: print(seed, sched, scheme, 'MISMATCH', (ref['stderr'] or '')[-200:], flush=True)
        bad += not ok
print('bad', bad, 'of', len(args))
