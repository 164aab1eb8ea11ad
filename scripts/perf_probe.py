"""Quick device timing probe (development aid): kernel ms for 1 replica and for a batch."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
from rlgpuschedule_b200 import synth as tracegen
import rlgpuschedule_b200 as rl

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    reps = [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else ['1', '148', '1184'])]
    rows = {'0': False, '1': True, 'd': 'device'}[sys.argv[3]] if len(sys.argv) > 3 else 'device'
    slot_cap = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    nstreams = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    cluster = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
    t0 = time.time()
    tr = rl.prepare_trace(tracegen.frame_gen(n, 3, n), cluster)
    print('ingest %.2fs' % (time.time() - t0), flush=True)
    for R in reps:
        sim = rl.Simulator(cluster, n_replicas=R, rows=rows, slot_cap=slot_cap, n_streams=nstreams)
        sim.load_trace(tr)
        for it in range(3):
            t0 = time.time(); sim.run(); wall = time.time() - t0
            ms, nl = sim.kernel_ms()
            s = sim.summary(0)
            ev = s['events'] * R
            print('R=%d rows=%s it=%d wall=%.3fs kernel=%.2fms launches=%d ticks=%d events/s(kernel)=%.3e ticks/s=%.3e maxR=%d maxQ=%d' % (
                R, rows, it, wall, ms, nl, s['n_ticks'], ev / (ms / 1e3), s['n_ticks'] * R / (ms / 1e3), s['max_running'], s['max_queued']), flush=True)
        sim.close()

main()
