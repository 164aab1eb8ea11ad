"""The legacy event loops (sjf / shortest / shortest-gpu over yarn, dlas-gpu / dlas with count admission) on the device.
(1) against the fixtures tests/golden/{sjf,shortest,shortestgpu,dlasgpu,dlas}_*: cluster.csv / job.csv written by the
reference's own log._Log while its dead-code loops (run_sim.py:162-287, :299-431, :664-947) ran UNMODIFIED under shim
globals (oracle/ref_legacy_runner.py; the live simulator raises NotImplementedError for these schedules) - byte-exact;
(2) against the CPU restatement oracle/cpu_sim.c (itself pinned on the same fixtures) on more traces - bit-exact on
every integer output: per-job start / end / pending / preempt / resume, finish order, and every per-event row."""
import numpy as np
import pytest

import cpu_sim
import golden_cases
import goldutil
import tracegen
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import _ffi
from rlgpuschedule_b200 import log_manager as lm

pytestmark = pytest.mark.gpu

SMALL = dict(num_switch=2, num_node_p_switch=4, num_gpu_p_node=8)
CASES = {
    'dense': (lambda: tracegen.frame_gen(300, 5, 30), SMALL),
    'dense2': (lambda: tracegen.frame_gen(500, 6, 60), dict(num_switch=1, num_node_p_switch=6, num_gpu_p_node=4)),
    'light': (lambda: tracegen.frame_gen(400, 8, 400), SMALL),
    'ties': (golden_cases.CASES['ties']['frame'], golden_cases.CASES['ties']['flags']),
    'multi_node': (golden_cases.CASES['multi_node']['frame'], golden_cases.CASES['multi_node']['flags']),
    'probe2k': (lambda: tracegen.frame_gen(2000, 1, 2000), golden_cases.C4328),
    'loaded': (lambda: tracegen.frame_gen(3000, 4, 500), golden_cases.C4328),
}


def compare(sim, tr, ores, check_rows=True):
    j = sim.jobs(0)
    assert np.array_equal(j['finish_order'], ores['finish_order'])
    assert np.array_equal(j['start'], ores['start'])
    assert np.array_equal(j['end'], ores['end'])
    assert np.array_equal(j['preempt'], ores['preempt'])
    assert np.array_equal(sim.job_plane(0, _ffi.PLANE_AUX), ores['pending'])
    assert np.array_equal(sim.job_plane(0, _ffi.PLANE_RESUME), ores['resume'])
    s = sim.summary(0)
    assert s['n_ticks'] == ores['n_events']
    assert s['sum_queued'] == ores['counters']['sweep_jobs']
    assert s['events'] == ores['counters']['events']
    if check_rows:
        rows = sim.rows(0)
        orow = ores['rows']
        assert np.array_equal(rows['median_lo'], orow['time'])
        assert np.array_equal(rows['idle_nodes'], orow['idle_nodes'])
        assert np.array_equal(rows['median_hi'], orow['full_nodes'])
        assert np.array_equal(rows['busy_gpus'], orow['busy_gpus'])
        assert np.array_equal(rows['queued'], orow['pending'])
        assert np.array_equal(rows['running'], orow['running'])
        assert np.array_equal(rows['finished'], orow['completed'])


@pytest.mark.parametrize('name', list(CASES))
@pytest.mark.parametrize('sched,mode', [('sjf', 0), ('shortest', 1), ('shortest-gpu', 2)])
def test_sjf_family_matches_restated_oracle(name, sched, mode):
    frame, flags = CASES[name]
    df = frame()
    cluster = rl.cluster_from_flags(flags)
    tr = rl.prepare_trace(df, cluster)
    sim = rl.Simulator(cluster, sched, 'yarn', n_replicas=3, rows=True)
    sim.load_trace(tr)
    sim.run()
    ores = cpu_sim.run_sjf_yarn(cpu_sim.make_cluster(**flags), cpu_sim.prepare_trace(df), sort_mode=mode)
    compare(sim, tr, ores)
    assert ores['preempt'].sum() > 0 or name not in ('dense', 'dense2')   # the dense cases must exercise preemption
    sim.close()


@pytest.mark.parametrize('name', list(CASES))
@pytest.mark.parametrize('limits', [(30, 60, 150), (8,), (5, 9, 14, 20, 33)])
def test_dlas_gpu_matches_restated_oracle(name, limits):
    frame, flags = CASES[name]
    df = frame()
    cluster = rl.cluster_from_flags(flags)
    tr = rl.prepare_trace(df, cluster)
    for sched, gputime in (('dlas-gpu', True), ('dlas', False)):
        sim = rl.Simulator(cluster, sched, 'count', n_replicas=2, rows=True, num_queue=len(limits) + 1, queue_limit=limits)
        sim.load_trace(tr)
        sim.run()
        ores = cpu_sim.run_dlas_gpu(cpu_sim.make_cluster(**flags), cpu_sim.prepare_trace(df), limits, gputime=gputime)
        compare(sim, tr, ores)
        assert sim.summary(0)['sum_running'] == ores['counters']['demotions']
        sim.close()


def test_legacy_bounded_launches_resume_exactly():
    frame, flags = CASES['dense']
    df = frame()
    cluster = rl.cluster_from_flags(flags)
    tr = rl.prepare_trace(df, cluster)
    for sched, scheme, kw in (('sjf', 'yarn', {}), ('dlas-gpu', 'count', dict(num_queue=4, queue_limit=(30, 60, 150)))):
        sim = rl.Simulator(cluster, sched, scheme, n_replicas=1, rows=True, ticks_per_launch=5, **kw)
        sim.load_trace(tr)
        sim.run()
        oc = cpu_sim.make_cluster(**flags)
        ot = cpu_sim.prepare_trace(df)
        ores = cpu_sim.run_sjf_yarn(oc, ot) if sched == 'sjf' else cpu_sim.run_dlas_gpu(oc, ot, (30, 60, 150))
        compare(sim, tr, ores)
        sim.close()


@pytest.mark.parametrize('name', goldutil.legacy_case_names(('small', 'big')))
def test_legacy_device_matches_reference_files(name):
    g = goldutil.load(name)
    cluster = rl.cluster_from_flags(g['flags'])
    tr = rl.prepare_trace(goldutil.trace_input(g), cluster)
    count = g['schedule'] in ('dlas-gpu', 'dlas')
    kw = dict(num_queue=len(g['queue_limit']) + 1, queue_limit=g['queue_limit']) if count else {}
    sim = rl.Simulator(cluster, g['schedule'], 'count' if count else 'yarn', n_replicas=2, rows=True, **kw)
    sim.load_trace(tr)
    sim.run()
    j = sim.jobs(1)
    assert lm.format_legacy_job_csv(tr, j, sim.job_plane(1, _ffi.PLANE_AUX), sim.job_plane(1, _ffi.PLANE_RESUME), count) == g['job']
    assert lm.format_legacy_cluster_csv(sim.rows(1), cluster, count) == g['cluster']
    sim.close()
