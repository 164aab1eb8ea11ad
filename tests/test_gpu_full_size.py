"""BASELINE.json's full-size configurations on the device, checked through size-independent properties
(the oracle comparison at these sizes is covered for one replica by the golden hashes):
conservation of jobs per row, capacity never exceeded, end - start = runtime, rows consistent with
the job table, identical replicas give identical results, sortedness of the finish order."""
import numpy as np
import pytest

import golden_cases
import tracegen
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import _ffi
from rlgpuschedule_b200.env import Environment

pytestmark = pytest.mark.gpu
C = rl.cluster_from_flags(golden_cases.C4328)


def _check_fifo_properties(sim, tr, r):
    j = sim.jobs(r)
    rows = sim.rows(r)
    rec = tr.records
    n = len(rec)
    st, en, fo = j['start'].astype(np.int64), j['end'].astype(np.int64), j['finish_order']
    assert len(fo) == n and len(np.unique(fo)) == n                          # every job finished exactly once
    assert np.array_equal(en - st, rec['dur_ticks'])                          # cf5: runtime = max(1, ceil(duration))
    assert (st >= rec['arrival_tick']).all()
    assert (np.diff(en[fo]) >= 0).all()                                       # job.csv is ordered by finish tick
    same = np.diff(en[fo]) == 0
    assert (np.diff(st[fo])[same] >= 0).all()                                 # ... and by start order inside a tick
    d = np.arange(1, len(rows) + 1)
    arrived = np.searchsorted(rec['arrival_tick'], d - 1, side='right')       # jobs with arrival_tick <= d-1
    assert np.array_equal(rows['queued'] + rows['running'] + rows['finished'], arrived)   # conservation
    assert np.array_equal(rows['finished'], np.searchsorted(np.sort(en), d, side='right'))
    assert np.array_equal(rows['running'], np.searchsorted(np.sort(st), d - 1, side='right') - rows['finished'])
    assert rows['busy_gpus'].max() <= C.num_gpus and (rows['idle_nodes'] >= 0).all()
    assert (np.diff(rows['idle_nodes']) <= 0).all()                            # q3: busy nodes are sticky
    ndev = rec['tasks'].astype(np.int64) * rec['gpus_per_task']
    delta = np.zeros(len(rows) + 2, np.int64)
    np.add.at(delta, st + 1, ndev); np.add.at(delta, en, -ndev)                # busy from row start+1 to row end-1
    assert np.array_equal(rows['busy_gpus'], np.cumsum(delta)[1:len(rows) + 1])
    q = rows['queued'] > 0
    assert (rows['median_lo'][q] <= rows['median_hi'][q]).all() and (rows['median_hi'][q] <= rows['max_pending'][q]).all()
    assert (rows['sum_pending'][q] >= rows['max_pending'][q]).all() and (rows['sum_pending'][~q] == 0).all()
    assert len(rows) == int(en.max())                                          # the run ends when the last job finishes


def test_fifo_60k_trace_many_replicas_properties():
    """config C3-sized trace (60k jobs), 592 replicas over two traces."""
    trs = [rl.prepare_trace(tracegen.frame_gen(60000, s, 60000), C) for s in (3, 4)]
    sim = rl.Simulator(C, 'fifo', 'yarn', n_replicas=592, rows='device')
    sim.load_trace(trs[0], 0, 296); sim.load_trace(trs[1], 296, 296)
    sim.run()
    for r, tr in ((0, trs[0]), (295, trs[0]), (296, trs[1]), (591, trs[1])):
        _check_fifo_properties(sim, tr, r)
    ret = sim.returns()
    assert (ret[:296] == ret[0]).all() and (ret[296:] == ret[296]).all() and ret[0] != ret[296]
    s0 = sim.summary(0)
    assert s0['events'] == 3 * 60000 and s0['n_ticks'] == 63167               # the pinned probe60k makespan
    sim.close()


def test_dlas_gpu_60k_trace_properties():
    """config C3: dlas-gpu, 4-queue MLFQ, 60k-job trace."""
    tr = rl.prepare_trace(tracegen.frame_gen(60000, 3, 60000), C)
    sim = rl.Simulator(C, 'dlas-gpu', 'count', n_replicas=64, rows='device', num_queue=4, queue_limit=(30, 60, 150))
    sim.load_trace(tr)
    sim.run()
    for r in (0, 63):
        j = sim.jobs(r); rows = sim.rows(r); rec = tr.records
        st, en = j['start'].astype(np.int64), j['end'].astype(np.int64)
        assert len(j['finish_order']) == 60000
        pend = sim.job_plane(r, _ffi.PLANE_AUX).astype(np.int64)
        assert np.array_equal(en - rec['arrival_tick'], rec['dur_ticks'] + pend)  # JCT = executed + pending
        assert (st >= rec['arrival_tick']).all() and (en - st >= rec['dur_ticks']).all()
        res, pre = sim.job_plane(r, _ffi.PLANE_RESUME), j['preempt']
        assert np.array_equal(res, pre + 1)                                       # every finished job resumed once more than it was preempted
        t = rows['median_lo']
        assert (np.diff(t) > 0).all()                                             # event times strictly increase
        assert (rows['busy_gpus'] <= C.num_gpus).all() and rows['finished'][-1] == 60000
        arrived = np.searchsorted(rec['arrival_tick'], t, side='right')
        assert np.array_equal(rows['queued'] + rows['running'] + rows['finished'], arrived)
    assert sim.returns()[0] == sim.returns()[63]
    sim.close()


def test_sjf_10k_trace_properties():
    """config C2: sjf + yarn, 10k-job trace."""
    tr = rl.prepare_trace(tracegen.frame_gen(10000, 2, 10000), C)
    sim = rl.Simulator(C, 'sjf', 'yarn', n_replicas=32, rows='device')
    sim.load_trace(tr)
    sim.run()
    j = sim.jobs(5); rows = sim.rows(5); rec = tr.records
    en = j['end'].astype(np.int64)
    assert len(j['finish_order']) == 10000 and (np.diff(en[j['finish_order']]) >= 0).all()
    pend = sim.job_plane(5, _ffi.PLANE_AUX).astype(np.int64)
    assert np.array_equal(en - rec['arrival_tick'], rec['dur_ticks'] + pend)
    assert (np.diff(rows['median_lo']) > 0).all() and (rows['busy_gpus'] <= C.num_gpus).all()
    assert (rows['idle_nodes'] + rows['median_hi'] <= C.num_nodes).all()
    sim.close()


def test_env_512_replicas_10k_trace():
    """config C4: 512 environment replicas of a 10k-job trace; the head policy reproduces fifo in every replica,
    the random-window policy gives replica-specific but reproducible returns."""
    tr = rl.prepare_trace(tracegen.frame_gen(10000, 2, 10000), C)
    env = Environment(C, tr, n_replicas=512, window_k=5, seed=1)
    env.reset(); env.rollout('head'); env.sync()
    head = env.sim.returns().copy()
    assert (head == head[0]).all() and env.sim.summary(7)['n_ticks'] == 13381   # the pinned probe10k makespan
    env.close()
    # a loaded cluster (queues build up), so that the choice inside the window matters
    tr2 = rl.prepare_trace(tracegen.frame_gen(10000, 4, 2500), C)
    env = Environment(C, tr2, n_replicas=512, window_k=5, seed=1)
    env.reset(); env.rollout('random'); env.sync()
    r1 = env.sim.returns().copy()
    env.reset(); env.rollout('random'); env.sync()
    assert np.array_equal(env.sim.returns(), r1)                                # same seed -> same episodes
    # the cluster is start-rate bound here (one start per tick), so the return -(sum JCT) can coincide between
    # replicas; the schedules themselves differ because every replica draws its own picks
    s0, s1, s2 = (env.sim.jobs(r)['start'] for r in (0, 1, 511))
    assert not np.array_equal(s0, s1) and not np.array_equal(s1, s2)
    assert bool(env.done.cpu().all())
    env.close()
