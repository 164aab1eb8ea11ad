"""BASELINE.json's full-size configurations on the device: one replica against the reference's own files (golden hashes) and
the oracle, every other replica through size-independent properties:
conservation of jobs per row, capacity never exceeded, end - start = runtime, rows consistent with
the job table, identical replicas give identical results, sortedness of the finish order."""
import numpy as np
import pytest

import cpu_sim
import golden_cases
import goldutil
import tracegen
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import _ffi
from rlgpuschedule_b200 import log_manager as lm
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


def _legacy_csvs(sim, tr, r, count):
    j = sim.jobs(r)
    return (lm.format_legacy_job_csv(tr, j, sim.job_plane(r, _ffi.PLANE_AUX), sim.job_plane(r, _ffi.PLANE_RESUME), count),
            lm.format_legacy_cluster_csv(sim.rows(r), C, count))


def test_dlas_gpu_60k_trace_matches_reference_and_oracle():
    """config C3: dlas-gpu, 4-queue MLFQ, 60k-job trace.  One replica is compared with the files the reference's dead-code
    loop wrote (tests/golden/dlasgpu_probe60k, 212 s of reference time) and with the oracle; the others through properties."""
    g = goldutil.load('dlasgpu_probe60k')
    df = goldutil.trace_input(g)
    tr = rl.prepare_trace(df, C)
    sim = rl.Simulator(C, 'dlas-gpu', 'count', n_replicas=64, rows='device', num_queue=4, queue_limit=(30, 60, 150))
    sim.load_trace(tr)
    sim.run()
    job, clu = _legacy_csvs(sim, tr, 63, True)
    assert goldutil.sha(job) == g['meta']['job_sha256'] and goldutil.sha(clu) == g['meta']['cluster_sha256']
    ores = cpu_sim.run_dlas_gpu(cpu_sim.make_cluster(**golden_cases.C4328), cpu_sim.prepare_trace(df), (30, 60, 150))
    j = sim.jobs(0)
    assert np.array_equal(j['finish_order'], ores['finish_order']) and np.array_equal(j['end'], ores['end'])
    assert np.array_equal(j['start'], ores['start']) and np.array_equal(j['preempt'], ores['preempt'])
    assert np.array_equal(sim.job_plane(0, _ffi.PLANE_AUX), ores['pending'])
    assert sim.summary(0)['events'] == ores['counters']['events'] and sim.summary(0)['n_ticks'] == ores['n_events']
    for r in (0, 63):
        j = sim.jobs(r); rows = sim.rows(r); rec = tr.records
        st, en = j['start'].astype(np.int64), j['end'].astype(np.int64)
        assert len(j['finish_order']) == 60000
        pend = sim.job_plane(r, _ffi.PLANE_AUX).astype(np.int64)
        res, pre = sim.job_plane(r, _ffi.PLANE_RESUME), j['preempt']
        normal = res == pre + 1            # jobs completed through a stale 'end_jobs' list (run_sim.py:706-717) end while preempted
        assert normal.mean() > 0.95 and ((res == pre) | normal).all()
        assert np.array_equal((en - rec['arrival_tick'])[normal], (rec['dur_ticks'] + pend)[normal])  # JCT = executed + pending
        assert (st >= rec['arrival_tick']).all()
        t = rows['median_lo']
        assert (np.diff(t) > 0).all()                                             # event times strictly increase
        assert (rows['busy_gpus'] <= C.num_gpus).all() and rows['finished'][-1] == 60000
        arrived = np.searchsorted(rec['arrival_tick'], t, side='right')
        assert np.array_equal(rows['queued'] + rows['running'] + rows['finished'], arrived)
    assert sim.returns()[0] == sim.returns()[63]
    sim.close()


def test_sjf_10k_trace_matches_reference_and_oracle():
    """config C2: sjf + yarn, 10k-job trace: byte-exact vs the reference's dead-code loop (tests/golden/sjf_probe10k) and the oracle."""
    g = goldutil.load('sjf_probe10k')
    df = goldutil.trace_input(g)
    tr = rl.prepare_trace(df, C)
    sim = rl.Simulator(C, 'sjf', 'yarn', n_replicas=32, rows='device')
    sim.load_trace(tr)
    sim.run()
    job, clu = _legacy_csvs(sim, tr, 5, False)
    assert goldutil.sha(job) == g['meta']['job_sha256'] and goldutil.sha(clu) == g['meta']['cluster_sha256']
    ores = cpu_sim.run_sjf_yarn(cpu_sim.make_cluster(**golden_cases.C4328), cpu_sim.prepare_trace(df))
    j = sim.jobs(31); rows = sim.rows(31); rec = tr.records
    assert np.array_equal(j['finish_order'], ores['finish_order']) and np.array_equal(j['end'], ores['end'])
    assert np.array_equal(rows['idle_nodes'], ores['rows']['idle_nodes']) and np.array_equal(rows['busy_gpus'], ores['rows']['busy_gpus'])
    en = j['end'].astype(np.int64)
    assert len(j['finish_order']) == 10000 and (np.diff(en[j['finish_order']]) >= 0).all()
    pend = sim.job_plane(31, _ffi.PLANE_AUX).astype(np.int64)
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
