"""Dev aid: small runs of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import synth
C = rl.Cluster(num_switch=2, num_node_p_switch=4, num_gpu_p_node=8)
df = synth.frame_gen(300, 11, 150)
tr = rl.prepare_trace(df, C)
C16 = rl.Cluster(num_switch=2, num_node_p_switch=3, num_gpu_p_node=16)    # > 8 GPUs per node: the split node-word layout of the fifo kernel
runs = [('fifo', 'yarn', dict(lanes_per_replica=8)), ('fifo', 'yarn', dict(lanes_per_replica=16, rows_format='wide')), ('fifo', 'yarn', dict(lanes_per_replica=32)),
        ('fifo', 'yarn', dict(lanes_per_replica=8, cluster=C16)), ('sjf', 'yarn', {}), ('shortest', 'yarn', {}), ('dlas-gpu', 'count', dict(num_queue=4, queue_limit=(30, 60, 150))),
        ('horus', 'horus', dict(pack_seed=3)), ('gandiva', 'gandiva', {}), ('horus+', 'horus+', dict(num_queue=3, pack_seed=1, pack_rng=False, num_buffer=15)),
        ('horus', 'yarn', {}), ('gandiva', 'yarn', {}), ('horus+', 'yarn', dict(num_queue=2, pack_seed=2, pack_rng=False))]
only = sys.argv[1:] 
for sched, scheme, kw in runs:
    if only and sched not in only: continue
    kw = dict(kw)
    cl = kw.pop('cluster', C)
    sim = rl.Simulator(cl, sched, scheme, n_replicas=7 if sched == 'fifo' else 3, rows=True, ticks_per_launch=97, **kw)
    sim.load_trace(tr if cl is C else rl.prepare_trace(df, cl)); sim.run()
    s = sim.summary(2)
    print(sched, scheme, s['n_ticks'], s['n_finished'], s['status'], flush=True)
    sim.close()
