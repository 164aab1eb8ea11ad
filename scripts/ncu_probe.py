"""One short run of the hot kernel for ncu capture (development aid; numbers under ncu are not bench values)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from rlgpuschedule_b200 import synth as tracegen
import rlgpuschedule_b200 as rl
n = int(sys.argv[1]); R = int(sys.argv[2]); rows = {'0': False, 'd': 'device'}[sys.argv[3]]
cluster = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
tr = rl.prepare_trace(tracegen.frame_gen(n, 3, n), cluster)
sim = rl.Simulator(cluster, n_replicas=R, rows=rows, n_streams=1)
sim.load_trace(tr)
for _ in range(int(sys.argv[4]) if len(sys.argv) > 4 else 2):
    sim.run()
print(sim.kernel_ms(), sim.summary(0)['n_ticks'])
