"""Measures BASELINE.json's five configurations on one GPU and writes profiles/r02_configs.json.
(C5's 8-GPU layout is exercised by bench.py under torchrun; here its per-GPU share of 512 replicas is run.)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import golden_cases
from rlgpuschedule_b200 import synth as tracegen
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200.env import Environment
import torch

def timed(fn, n=3):
    fn(); best = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best

out = {}
C1 = rl.cluster_from_flags(golden_cases.C148); C4 = rl.cluster_from_flags(golden_cases.C4328)
# C1 fifo + yarn, 1x4x8, 100 jobs
tr = rl.prepare_trace(tracegen.frame_probe100(), C1)
sim = rl.Simulator(C1, 'fifo', 'yarn', n_replicas=1, rows=True); sim.load_trace(tr)
w = timed(sim.run); s = sim.summary(0); ms, _ = sim.kernel_ms()
out['C1 fifo+yarn 1x4x8 100 jobs, 1 replica'] = dict(wall_ms=w * 1e3, kernel_ms=ms, ticks=s['n_ticks'], events=s['events'], events_per_s=s['events'] / w, reference_python_s=7.0)
sim.close()
# C2 sjf + yarn 10k
tr = rl.prepare_trace(tracegen.frame_gen(10000, 2, 10000), C4)
for R in (1, 2368):
    sim = rl.Simulator(C4, 'sjf', 'yarn', n_replicas=R, rows='device'); sim.load_trace(tr)
    w = timed(sim.run, 2); s = sim.summary(0); ms, _ = sim.kernel_ms()
    out['C2 sjf+yarn 4x32x8 10k jobs, %d replica(s)' % R] = dict(wall_ms=w * 1e3, kernel_ms=ms, event_rows=s['n_ticks'], events=s['events'], events_per_s=s['events'] * R / w,
                                                               job_updates_per_s=s['sum_queued'] * R / w, algorithmic_GBps=44 * s['sum_queued'] * R / w / 1e9)
    sim.close()
# C3 dlas-gpu 60k
tr60 = rl.prepare_trace(tracegen.frame_gen(60000, 3, 60000), C4)
for R in (1, 2960):
    sim = rl.Simulator(C4, 'dlas-gpu', 'count', n_replicas=R, rows='device', num_queue=4, queue_limit=(30, 60, 150)); sim.load_trace(tr60)
    w = timed(sim.run, 2); s = sim.summary(0); ms, _ = sim.kernel_ms()
    out['C3 dlas-gpu 4 queues 4x32x8 60k jobs, %d replica(s)' % R] = dict(wall_ms=w * 1e3, kernel_ms=ms, event_rows=s['n_ticks'], events=s['events'], events_per_s=s['events'] * R / w,
                                                                        job_updates_per_s=s['sum_queued'] * R / w, algorithmic_GBps=48 * s['sum_queued'] * R / w / 1e9)
    sim.close()
# fifo 60k single replica (the reference's 709 s run)
sim = rl.Simulator(C4, 'fifo', 'yarn', n_replicas=1, rows=True); sim.load_trace(tr60)
w = timed(sim.run); s = sim.summary(0)
out['fifo+yarn 4x32x8 60k jobs, 1 replica (reference: 709 s)'] = dict(wall_ms=w * 1e3, kernel_ms=sim.kernel_ms()[0], ticks=s['n_ticks'], events_per_s=s['events'] / w, speedup_vs_python_reference=709.0 / w)
sim.close()
# C4 env: 512 replicas x 10k-job trace
tr10 = rl.prepare_trace(tracegen.frame_gen(10000, 2, 10000), C4)
env = Environment(C4, tr10, n_replicas=512, window_k=5, seed=1)
def roll():
    env.reset(); env.rollout('random'); env.sync()
w = timed(roll); s = env.sim.summary(0)
out['C4 env rollout (random window policy on device) 512 replicas x 10k jobs'] = dict(wall_ms=w * 1e3, ticks=s['n_ticks'], env_steps_per_s=512 * s['n_ticks'] / w, events_per_s=512 * s['events'] / w)
# per-step API (one launch per tick, actions from the host policy)
env.reset(); a = torch.zeros(512, dtype=torch.int32, device='cuda')
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(2000): env.step(a)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out['C4 env step API (external actions, 1 launch per tick) 512 replicas'] = dict(us_per_step=dt / 2000 * 1e6, env_steps_per_s=512 * 2000 / dt)
env.close()
# C4 with enough environment replicas to fill the GPU (4 per warp)
env = Environment(C4, tr10, n_replicas=8880, window_k=5, seed=1)
w = timed(lambda: (env.reset(), env.rollout('random'), env.sync())); s = env.sim.summary(0)
out['env rollout (random window policy on device) 8880 replicas x 10k jobs'] = dict(wall_ms=w * 1e3, ticks=s['n_ticks'], env_steps_per_s=8880 * s['n_ticks'] / w, events_per_s=8880 * s['events'] / w)
env.close()
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'r02_configs.json'), 'w'), indent=1)
print(json.dumps(out, indent=1))
