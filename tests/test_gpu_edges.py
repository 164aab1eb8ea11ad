"""Edge cases of the boundary and of the cluster shapes: empty / single-job traces, one node, 32 GPUs per node
(full device mask), a 512-node cluster (16 node words per lane), slot-table growth, row-store growth."""
import numpy as np
import pytest

import cpu_sim
import tracegen
from rlgpuschedule_b200 import synth
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import _ffi, log_manager as lm

pytestmark = pytest.mark.gpu


def _cmp(df, flags, **kw):
    cluster = rl.cluster_from_flags(flags)
    tr = rl.prepare_trace(df, cluster)
    sim = rl.Simulator(cluster, 'fifo', 'yarn', n_replicas=2, rows=True, **kw)
    sim.load_trace(tr)
    sim.run()
    otr = cpu_sim.prepare_trace(df)
    o = cpu_sim.run_fifo_yarn(cpu_sim.make_cluster(**flags), otr)
    for r in (0, 1):
        j = sim.jobs(r)
        assert np.array_equal(j['finish_order'], o['finish_order'])
        assert np.array_equal(j['start'], o['start']) and np.array_equal(j['end'], o['end'])
        assert lm.format_cluster_csv(sim.rows(r), cluster, tr.mem_shift, with_util=False) == cpu_sim.format_cluster_csv(o)
    return sim, o


def test_empty_trace_is_rejected_like_the_reference_asserts():
    cluster = rl.Cluster()
    df = tracegen.frame_rows([dict(type='interactive')])          # every row filtered out
    tr = rl.prepare_trace(df, cluster)
    assert len(tr) == 0
    sim = rl.Simulator(cluster)
    with pytest.raises(_ffi.RlgsError) as e:
        sim.load_trace(tr)
    assert e.value.code == _ffi.ERR_BAD_ARG
    with pytest.raises(_ffi.RlgsError) as e:
        sim.run()                                                  # run before load_trace
    assert e.value.code == _ffi.ERR_STATE
    sim.close()


def test_single_job_single_node():
    sim, o = _cmp(tracegen.frame_rows([dict(minutes=3.0, used_gpus=2.0, gpu_per_container=1)]), dict(num_switch=1, num_node_p_switch=1, num_gpu_p_node=2))
    assert sim.summary(0)['n_ticks'] == 2 and len(o['finish_order']) == 1
    sim.close()


def test_32_gpus_per_node_full_mask():
    rng = np.random.default_rng(2)
    rows = [dict(normalized_time=float(t), minutes=float(m), used_gpus=float(g), gpu_per_container=int(min(c, g)))
            for t, m, g, c in zip(np.sort(rng.uniform(0, 4e5, 80)), rng.uniform(2, 60, 80), rng.choice([1, 8, 16, 32, 40, 64], 80), rng.choice([1, 8], 80))]
    sim, _ = _cmp(tracegen.frame_rows(rows), dict(num_switch=1, num_node_p_switch=3, num_gpu_p_node=32, num_cpu_p_node=512, mem_p_node=2048))
    sim.close()


def test_512_node_cluster_and_slot_table_growth():
    # 8 switches x 64 nodes x 2 GPUs, thousands of one-GPU jobs running at once: more running jobs than the
    # default 128 on-chip slots, so the handle is rebuilt with a larger table
    df = tracegen.frame_gen(2500, 21, 40)
    df['used_gpus'] = 1.0; df['gpu_per_container'] = 1
    df['minutes'] = np.random.default_rng(5).uniform(300, 900, len(df))   # long jobs: one start per tick fills > 128 slots
    sim, o = _cmp(df, dict(num_switch=8, num_node_p_switch=64, num_gpu_p_node=2))
    assert sim.summary(0)['max_running'] > 128
    sim.close()


def test_row_store_grows_across_chunks():
    # a long makespan (one long job) with few arrivals: rows far beyond max_arrival + 4096
    rows = [dict(normalized_time=0, minutes=30000.0, used_gpus=1.0, gpu_per_container=1),
            dict(normalized_time=50000, minutes=2.0, used_gpus=1.0, gpu_per_container=1)]
    for mode in (True, 'device'):
        cluster = rl.Cluster(num_switch=1, num_node_p_switch=1, num_gpu_p_node=2)
        tr = rl.prepare_trace(tracegen.frame_rows(rows), cluster)
        sim = rl.Simulator(cluster, n_replicas=3, rows=mode)
        sim.load_trace(tr); sim.run()
        assert sim.summary(2)['n_ticks'] == 15000
        r = sim.rows(2)
        assert len(r) == 15000 and r['finished'][-1] == 2 and r['running'][7000] == 1 and r['finished'][5] == 1
        sim.run()                                                   # second run sizes the pipeline from the first
        assert np.array_equal(sim.rows(1), r)
        sim.close()


@pytest.mark.parametrize('mode', [True, 'device'])
def test_replicas_with_different_lengths_across_row_chunks(mode):
    """Three traces whose makespans straddle the 4096-row chunks differently, reloaded and rerun (the e2e pattern)."""
    flags = dict(num_switch=1, num_node_p_switch=2, num_gpu_p_node=4)
    cluster = rl.cluster_from_flags(flags)

    def frame(long_minutes, n_extra, seed):
        rng = np.random.default_rng(seed)
        rows = [dict(normalized_time=0, minutes=float(long_minutes), used_gpus=2.0, gpu_per_container=1)]
        rows += [dict(normalized_time=float(t), minutes=float(m), used_gpus=float(g), gpu_per_container=1)
                 for t, m, g in zip(np.sort(rng.uniform(1e4, 2e7, n_extra)), rng.uniform(1, 300, n_extra), rng.choice([1, 2, 4], n_extra))]
        return synth.frame_rows(rows)
    frames = [frame(6000, 40, 1), frame(10000, 60, 2), frame(18500, 80, 3)]
    traces = [rl.prepare_trace(f, cluster) for f in frames]
    sim = rl.Simulator(cluster, n_replicas=5, rows=mode, fetch_jobs=True)
    layout = [(0, 2, 0), (2, 1, 1), (3, 2, 2)]
    for rep in range(2):
        for first, count, t in layout:
            sim.load_trace(traces[t], first, count)
        sim.run()
        for first, count, t in layout:
            o = cpu_sim.run_fifo_yarn(cpu_sim.make_cluster(**flags), cpu_sim.prepare_trace(frames[t]))
            for r in (first, first + count - 1):
                assert sim.summary(r)['n_ticks'] == o['n_ticks']
                j = sim.jobs(r)
                assert np.array_equal(j['end'], o['end']) and np.array_equal(j['finish_order'], o['finish_order'])
                assert lm.format_cluster_csv(sim.rows(r), cluster, traces[t].mem_shift, with_util=False) == cpu_sim.format_cluster_csv(o)
    sim.close()
