"""Network-cost model (core/network/network_service.py:3-39) on the device vs the CPU restatement.
PARITY UNPINNED: the reference's own path raises AttributeError when enabled (job.py:199-200), so the
semantics are build-defined; float64 durations must still match the oracle bit for bit."""
import numpy as np
import pytest

import cpu_sim
import tracegen
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import log_manager as lm

pytestmark = pytest.mark.gpu
FLAGS = dict(num_switch=2, num_node_p_switch=4, num_gpu_p_node=8)


def _frame(n=250, seed=12, span=60):
    df = tracegen.frame_gen(n, seed, span)
    rng = np.random.default_rng(seed)
    names = ['resnet50', 'vgg16', 'transformer', 'alexnet', 'unknown_model']
    df['model_name'] = rng.choice(names, n)
    df['iterations'] = rng.integers(0, 400, n).astype(float)
    return df


@pytest.mark.parametrize('bw,lat', [(1250, 0.015), (10, 0.5)])
def test_network_costs_match_oracle_bit_for_bit(bw, lat):
    df = _frame()
    cluster = rl.cluster_from_flags(FLAGS)
    tr = rl.prepare_trace(df, cluster)
    assert tr.model_mb.max() == 1100 and tr.model_mb.min() == 0
    sim = rl.Simulator(cluster, 'fifo', 'yarn', n_replicas=3, rows=True, enable_network_costs=True, bandwidth=bw, internode_latency=lat)
    sim.load_trace(tr)
    sim.run()
    otr = cpu_sim.prepare_trace(df)
    o = cpu_sim.run_fifo_yarn(cpu_sim.make_cluster(**FLAGS), otr,
                              netcost=dict(model_mb=tr.model_mb, iterations=tr.iterations, bandwidth=bw, latency=lat))
    base = cpu_sim.run_fifo_yarn(cpu_sim.make_cluster(**FLAGS), otr)
    for r in (0, 2):
        j = sim.jobs(r)
        assert np.array_equal(j['finish_order'], o['finish_order'])
        assert np.array_equal(j['start'], o['start']) and np.array_equal(j['end'], o['end'])
        dur = sim.durations(r)
        assert np.array_equal(dur, o['actual_duration'])           # float64, bit for bit
        job = lm.format_job_csv(tr, j['finish_order'], j['start'], j['end'], j['preempt'], actual_duration=dur)
        assert job == cpu_sim.format_job_csv(otr, o)
        assert lm.format_cluster_csv(sim.rows(r), cluster, tr.mem_shift, with_util=False) == cpu_sim.format_cluster_csv(o)
    assert (o['actual_duration'] >= otr['duration']).all() and (o['actual_duration'] > otr['duration']).any()
    assert not np.array_equal(o['end'], base['end'])                 # the costs change the schedule
    sim.close()


def test_network_costs_off_by_default_and_fifo_only():
    cluster = rl.cluster_from_flags(FLAGS)
    with pytest.raises(rl._ffi.RlgsError):
        rl.Simulator(cluster, 'sjf', 'yarn', enable_network_costs=True)
