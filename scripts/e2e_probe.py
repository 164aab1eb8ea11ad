import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
from rlgpuschedule_b200 import synth as tracegen
import rlgpuschedule_b200 as rl
R = int(sys.argv[1]); NT = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cluster = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
traces = [rl.prepare_trace(tracegen.frame_gen(60000, 3 + i, 60000), cluster) for i in range(NT)]
bounds = [R * i // NT for i in range(NT + 1)]
sim = rl.Simulator(cluster, n_replicas=R, rows='host', fetch_jobs=True)
def attach():
    for i, tr in enumerate(traces):
        sim.load_trace(tr, bounds[i], bounds[i + 1] - bounds[i])
attach()
for it in range(4):
    t0 = time.perf_counter(); attach(); t1 = time.perf_counter(); sim.run(); t2 = time.perf_counter()
    print('R=%d traces=%d attach %.1f ms run %.1f ms kernel %.1f ms ticks %s' % (R, NT, (t1 - t0) * 1e3, (t2 - t1) * 1e3, sim.kernel_ms()[0], [sim.summary(b)['n_ticks'] for b in bounds[:-1]]), flush=True)
