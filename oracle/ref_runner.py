"""Run the UNMODIFIED reference simulator (read-only at /root/reference) as the parity oracle.

Test infrastructure only: used in THIS container to pin oracle/cpu_sim.c and to generate the
fixtures under tests/golden/ (see oracle/make_golden.py).  /root/reference does not exist on
the GPU box, so nothing on the `-m gpu` / bench path imports this module.

The reference is started exactly as its README says (`python run_sim.py --flags`,
/root/reference/run_sim.py:1742-1757) in a scratch working directory, with
`logging.disable(CRITICAL)` (outputs verified unchanged) so a 10k-job run takes
~70 s instead of ~100 s.
"""
import glob
import os
import subprocess
import sys
import tempfile
import time

REF = os.environ.get('RLGS_REFERENCE_DIR', '/root/reference')

_DRIVER = r'''
import logging, runpy, sys
logging.disable(logging.CRITICAL)
sys.path.insert(0, {ref!r})
sys.argv = ['run_sim.py'] + {argv!r}
{inject}
try:
    runpy.run_path({ref!r} + '/run_sim.py', run_name='__main__')
except SystemExit:
    pass
'''


def available():
    return os.path.exists(os.path.join(REF, 'run_sim.py'))


# Injected randomness for the k-means of horus+ (core/jobs/utils.py:39,60): the reference's code runs unmodified, but
# numpy's module-level randint / choice are replaced by counter-based draws (call number, element) that oracle/cpu_sim.c
# reproduces.  Everything else in np.random stays as it is.
_INJECT = r'''
import numpy as _np
_SEED = {seed}
_calls = [0]
def _mix(z):
    z &= 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)
def _draw(call, elem, n):
    h = _mix(((_SEED << 32) | (call & 0xFFFFFFFF)) + 0x9E3779B97F4A7C15)
    h = _mix(h ^ elem)
    return int((h >> 11) % n)
def _randint(low, high=None, size=None, dtype=int):
    assert high is None
    c = _calls[0]; _calls[0] += 1
    if size is None:
        return _draw(c, 0, low)
    return _np.array([_draw(c, i, low) for i in range(size)])
def _choice(a, size=None, replace=True, p=None):
    assert size is None and p is None
    c = _calls[0]; _calls[0] += 1
    return _draw(c, 0, a)
_np.random.randint = _randint
_np.random.choice = _choice
'''


def run_reference(trace_csv, workdir=None, schedule='fifo', scheme='yarn', inject_seed=None, **flags):
    """Returns dict(job_csv=str, cluster_csv=str, wall_s=float, out_dir=str).  inject_seed: replace np.random.randint /
    np.random.choice by the counter-based draws above (horus+ only)."""
    if not available():
        raise RuntimeError('reference not mounted at %s' % REF)
    workdir = workdir or tempfile.mkdtemp(prefix='rlgs_ref_')
    argv = ['--trace_file', os.path.abspath(trace_csv), '--schedule', schedule,
            '--scheme', scheme, '--log_path', 'oracle']
    for k, v in flags.items():
        argv += ['--' + k, str(v)]
    code = _DRIVER.format(ref=REF, argv=argv, inject=_INJECT.format(seed=int(inject_seed)) if inject_seed is not None else '')
    t0 = time.time()
    p = subprocess.run([sys.executable, '-c', code], cwd=workdir, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    wall = time.time() - t0
    outs = sorted(glob.glob(os.path.join(workdir, 'log', 'oracle', '*')))
    if not outs:
        raise RuntimeError('reference produced no output: rc=%s\n%s' % (p.returncode, p.stderr[-2000:]))
    out_dir = outs[-1]
    res = dict(wall_s=wall, out_dir=out_dir, returncode=p.returncode, stderr=p.stderr[-4000:])
    for name in ('job', 'cluster'):
        fn = os.path.join(out_dir, name + '.csv')
        # newline='' keeps the csv module's '\r\n' terminators: fixtures are byte-exact
        res[name + '_csv'] = open(fn, newline='').read() if os.path.exists(fn) else None
    return res


def strip_util_column(cluster_csv):
    """cluster.csv minus avg_gpu_utilization (column 6): that column is drawn from an unseeded
    numpy RNG in the reference (/root/reference/infra/device.py:52) and cannot be pinned."""
    out = []
    for line in cluster_csv.split('\r\n'):
        if line:
            f = line.split(',')
            out.append(','.join(f[:5] + f[6:]))
    return '\r\n'.join(out) + '\r\n'
