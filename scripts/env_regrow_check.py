import sys; sys.path.insert(0, '/root/repo')
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import synth
from rlgpuschedule_b200.env import Environment
cluster = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
traces = [rl.prepare_trace(synth.frame_gen(10000, 1032 + i, 10000), cluster) for i in range(32)]
env = Environment(cluster, [(traces[i], 16 * i, 16) for i in range(32)], n_replicas=512, window_k=5, seed=1)
env.run_episodes('random')
print('slot_cap now', env.sim._slot_cap, 'max_running', max(env.sim.summary(r)['max_running'] for r in range(512)), 'done', bool(env.done.cpu().all()))
env.close()
