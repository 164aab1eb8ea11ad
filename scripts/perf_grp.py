"""Timing sweep of the fifo tick loop over lanes-per-replica x replica count x row format (development aid).

    python scripts/perf_grp.py [n_jobs] [n_traces] [configs]      configs = "lpr:replicas:rows,..." rows in {0, wide, wire16}

Replicas are attached in contiguous blocks to `n_traces` distinct traces; the kernel puts replicas a quarter of a
launch apart into one warp, so with >= 16 traces the replicas sharing a warp run different traces.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rlgpuschedule_b200 import synth  # noqa: E402
import rlgpuschedule_b200 as rl  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    nt = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    cfgs = sys.argv[3] if len(sys.argv) > 3 else '32:2960:wide,32:2960:wire16,16:2960:wire16,16:4736:wire16,8:2960:wire16,8:7104:wire16,8:9472:wire16'
    cluster = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
    t0 = time.time()
    traces = [rl.prepare_trace(synth.frame_gen(n, 3 + i, n), cluster) for i in range(nt)]
    print('ingest %.2fs' % (time.time() - t0), flush=True)
    out = []
    for cfg in cfgs.split(','):
        lpr, R, rows = cfg.split(':')
        lpr, R = int(lpr), int(R)
        kw = dict(rows=False) if rows == '0' else dict(rows='device', rows_format=rows)
        sim = rl.Simulator(cluster, n_replicas=R, lanes_per_replica=lpr, **kw)
        # contiguous blocks of replicas per trace; the kernel puts replicas a quarter of the launch apart into one warp
        for i in range(nt):
            lo, hi = R * i // nt, R * (i + 1) // nt
            if hi > lo:
                sim.load_trace(traces[i], lo, hi - lo)
        bounds = [R * i // nt for i in range(nt + 1)]
        best = None
        for it in range(3):
            t0 = time.time(); sim.run(); wall = time.time() - t0
            ms, nl = sim.kernel_ms()
            ev = sum(sim.summary(bounds[i])['events'] * (bounds[i + 1] - bounds[i]) for i in range(nt) if bounds[i + 1] > bounds[i])
            tk = sum(sim.summary(bounds[i])['n_ticks'] * (bounds[i + 1] - bounds[i]) for i in range(nt) if bounds[i + 1] > bounds[i])
            rec = dict(lpr=lpr, R=R, rows=rows, it=it, wall_s=round(wall, 3), kernel_ms=round(ms, 2), launches=nl,
                       events_per_s=ev / (ms / 1e3), ticks_per_s=tk / (ms / 1e3))
            print(json.dumps(rec), flush=True)
            if best is None or rec['events_per_s'] > best['events_per_s']:
                best = rec
        out.append(best)
        sim.close()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'perf_grp.json'), 'w'), indent=1)


main()
