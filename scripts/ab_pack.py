"""Dev aid: A/B timing of builds of librlgs on the horus path (RLGS_LIB selects the build; argv = replica counts)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import synth
C = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
tr = rl.prepare_trace(synth.frame_gen(2000, 1, 2000), C)
for R in [int(a) for a in sys.argv[1:]] or [1, 1776]:
    sim = rl.Simulator(C, 'horus', 'horus', n_replicas=R, rows='device'); sim.load_trace(tr)
    sim.run(); a = sim.kernel_ms()[0]; sim.run(); b = sim.kernel_ms()[0]
    print(os.environ.get('RLGS_LIB', 'default'), R, round(a, 1), round(b, 1), 'Mticks/s %.2f' % (R * sim.summary(0)['n_ticks'] / b / 1e3))
    sim.close()
