import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rlgpuschedule_b200 import synth
import rlgpuschedule_b200 as rl
sched = sys.argv[1]; n = int(sys.argv[2]); R = int(sys.argv[3])
cluster = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
tr = rl.prepare_trace(synth.frame_gen(n, 3, n), cluster)
kw = dict(num_queue=4, queue_limit=(30, 60, 150)) if sched == 'dlas-gpu' else {}
sim = rl.Simulator(cluster, sched, 'count' if sched == 'dlas-gpu' else 'yarn', n_replicas=R, rows='device', n_streams=1, **kw)
sim.load_trace(tr)
for _ in range(2):
    sim.run()
print(sim.kernel_ms(), sim.summary(0)['n_ticks'], sim.summary(0)['sum_queued'])
