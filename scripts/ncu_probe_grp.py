"""One short run of the fifo tick loop for an ncu capture (development aid; numbers under ncu are not bench values).

    python scripts/ncu_probe_grp.py n_jobs lpr replicas rows_format n_traces [runs]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rlgpuschedule_b200 import synth  # noqa: E402
import rlgpuschedule_b200 as rl  # noqa: E402

n, lpr, R, fmt, nt = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
cluster = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
traces = [rl.prepare_trace(synth.frame_gen(n, 3 + i, n), cluster) for i in range(nt)]
kw = dict(rows=False) if fmt == '0' else dict(rows='device', rows_format=fmt)
sim = rl.Simulator(cluster, n_replicas=R, lanes_per_replica=lpr, n_streams=1, rows_cap=n + 8192, **kw)   # one launch per run
for i in range(nt):
    lo, hi = R * i // nt, R * (i + 1) // nt
    if hi > lo:
        sim.load_trace(traces[i], lo, hi - lo)
for _ in range(int(sys.argv[6]) if len(sys.argv) > 6 else 2):
    sim.run()
ticks = sum(sim.summary(R * i // nt)['n_ticks'] * (R * (i + 1) // nt - R * i // nt) for i in range(nt))
print('kernel_ms', sim.kernel_ms(), 'replica_ticks', ticks)
