"""Device timing of the sjf / dlas-gpu kernels (development aid)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from rlgpuschedule_b200 import synth as tracegen
import rlgpuschedule_b200 as rl
sched = sys.argv[1]; n = int(sys.argv[2]); span = int(sys.argv[3]); reps = [int(x) for x in sys.argv[4].split(',')]
cluster = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
tr = rl.prepare_trace(tracegen.frame_gen(n, 3, span), cluster)
for R in reps:
    kw = dict(num_queue=4, queue_limit=(30, 60, 150)) if sched == 'dlas-gpu' else {}
    sim = rl.Simulator(cluster, sched, 'count' if sched == 'dlas-gpu' else 'yarn', n_replicas=R, rows='device', **kw)
    sim.load_trace(tr)
    for it in range(2):
        t0 = time.time(); sim.run(); wall = time.time() - t0
    ms, nl = sim.kernel_ms(); s = sim.summary(0)
    print('%s n=%d span=%d R=%d kernel=%.1fms events(rows)=%d sim-events=%d sweep_jobs=%d maxM=%d  -> %.3e events/s, %.3e job-updates/s' % (
        sched, n, span, R, ms, s['n_ticks'], s['events'], s['sum_queued'], s['max_queued'], s['events'] * R / (ms / 1e3), s['sum_queued'] * R / (ms / 1e3)), flush=True)
    sim.close()
