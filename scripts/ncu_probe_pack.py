import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rlgpuschedule_b200 import synth
import rlgpuschedule_b200 as rl
sched = sys.argv[1]; n = int(sys.argv[2]); seed = int(sys.argv[3]); R = int(sys.argv[4])
cluster = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
tr = rl.prepare_trace(synth.frame_gen(n, seed, n), cluster)
sim = rl.Simulator(cluster, sched, sched, n_replicas=R, rows='device', n_streams=1)
sim.load_trace(tr)
sim.run()
s = sim.summary(0)
print(sim.kernel_ms(), s['n_ticks'], s['n_finished'], s['events'])
