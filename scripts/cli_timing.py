import os, subprocess, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rlgpuschedule_b200 import synth
d = tempfile.mkdtemp()
for n, span in ((10000, 10000), (60000, 60000)):
    fn = os.path.join(d, 't%d.csv' % n)
    synth.write(synth.frame_gen(n, 3 if n == 60000 else 2, span), fn)
    for sched, scheme in (('fifo', 'yarn'), ('dlas-gpu', 'count')):
        t0 = time.time()
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'run_sim.py'), '--trace_file', fn, '--num_switch', '4', '--num_node_p_switch', '32',
                            '--schedule', sched, '--scheme', scheme, '--num_queue', '4', '--log_path', 'x'], cwd=d, capture_output=True, text=True)
        print(n, sched, 'rc', r.returncode, 'wall %.2fs' % (time.time() - t0), [l for l in r.stderr.splitlines() if 'device:' in l][-1:], flush=True)
