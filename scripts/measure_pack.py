"""Measures the pack schedules (horus, gandiva) on one GPU and writes gpurun_out/r01_pack.json (copied to profiles/)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import synth

C = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
REF = {('horus', 'horus', 'probe2k'): 108.96, ('gandiva', 'gandiva', 'probe2k'): 137.85,
       ('horus', 'yarn', 'probe2k'): 36.66, ('gandiva', 'yarn', 'probe2k'): 39.66, ('horus+', 'horus+', 'probe2k'): 94.66}   # tests/golden/*/meta.json reference_wall_s (this container, 1 core)
out = {}
for name, n, seed, reps in (('probe2k', 2000, 1, (1, 1184)), ('probe10k', 10000, 2, (1, 1184))):
    df = synth.frame_gen(n, seed, n).copy(); df['gpu_utilization_max'] = df['gpu_utilization_avg']
    tr = rl.prepare_trace(df, C)
    for sched, scheme in (('horus', 'horus'), ('horus+', 'horus+'), ('gandiva', 'gandiva'), ('horus', 'yarn'), ('horus+', 'yarn'), ('gandiva', 'yarn')):
        for R in reps:
            kw = dict(num_queue=3, num_buffer=15, pack_seed=1, pack_rng=False) if sched == 'horus+' else {}
            sim = rl.Simulator(C, sched, scheme, n_replicas=R, rows='device', **kw)
            sim.load_trace(tr)
            sim.run(); sim.run()
            ms, nl = sim.kernel_ms(); s = sim.summary(0)
            key = '%s+%s %s 4x32x8, %d replica(s)%s' % (sched, scheme, name, R, ', mean draws' if scheme != 'yarn' else '')
            out[key] = dict(kernel_ms=ms, ticks=s['n_ticks'], finished=s['n_finished'], events=s['events'],
                            ticks_per_s=R * s['n_ticks'] / (ms / 1e3), events_per_s=R * s['events'] / (ms / 1e3))
            if R == 1 and (sched, scheme, name) in REF:
                out[key]['reference_python_s'] = REF[(sched, scheme, name)]
                out[key]['speedup_vs_python_reference'] = REF[(sched, scheme, name)] / (ms / 1e3)
            print(key, json.dumps(out[key]), flush=True)
            sim.close()
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'r01_pack.json'), 'w'), indent=1)
