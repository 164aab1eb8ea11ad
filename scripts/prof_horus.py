"""Dev aid: per-phase cycle breakdown of the horus kernel.  Build the profiling variant first:
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared --fmad=false -DPACK_PROFILE \\
         -o rlgpuschedule_b200/librlgs_prof.so rlgpuschedule_b200/csrc/rlgs_api.cu
and run with RLGS_LIB=rlgpuschedule_b200/librlgs_prof.so (the counters are printed by rlgs_get_summary)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import synth
C = rl.Cluster(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
n, seed, span = (int(x) for x in sys.argv[1:4])
tr = rl.prepare_trace(synth.frame_gen(n, seed, span), C)
sim = rl.Simulator(C, 'horus', 'horus', n_replicas=1, rows='device', pack_seed=(int(sys.argv[4]) if len(sys.argv) > 4 else None))
sim.load_trace(tr); sim.run()
print(sim.kernel_ms(), {k: v for k, v in sim.summary(0).items() if k in ('n_ticks', 'n_finished', 'n_started', 'max_queued', 'max_running')})
