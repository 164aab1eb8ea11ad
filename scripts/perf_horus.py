"""Dev aid: timings of the horus schedule on the device (single replica and replica batches)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import synth

C = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
out = {}
for name, n, seed, span, reps in (('probe2k', 2000, 1, 2000, (1, 1776, 3552)), ('probe10k', 10000, 2, 10000, (1, 1776))):
    df = synth.frame_gen(n, seed, span)
    tr = rl.prepare_trace(df, C)
    for R in reps:
        for seedp in ((None, 7) if R == 1 else (None,)):
            sim = rl.Simulator(C, 'horus', 'horus', n_replicas=R, rows='device', pack_seed=seedp)
            sim.load_trace(tr)
            t0 = time.time(); sim.run(); wall = time.time() - t0
            ms, nl = sim.kernel_ms()
            s = sim.summary(0)
            ticks = sum(sim.summary(r)['n_ticks'] for r in range(0, R, max(1, R // 8))) / len(range(0, R, max(1, R // 8)))
            ev = s['events']
            key = '%s_R%d_%s' % (name, R, 'rng' if seedp is not None else 'mean')
            out[key] = dict(kernel_ms=ms, wall_s=wall, ticks=s['n_ticks'], finished=s['n_finished'], events=ev,
                            ticks_per_s=R * ticks / (ms / 1e3), events_per_s=R * ev / (ms / 1e3))
            print(key, json.dumps(out[key]), flush=True)
            sim.close()
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/perf_horus.json', 'w'), indent=1)
