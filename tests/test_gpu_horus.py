"""horus schedule + horus placement on the device vs the CPU restatement (oracle_pack), which is pinned byte-for-byte
against the real reference on zero-spread traces (tests/golden/horus_*, tests/test_oracle_golden.py)."""
import numpy as np
import pytest

import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import _ffi, log_manager as lm
from oracle import cpu_sim, tracegen
import goldutil

pytestmark = pytest.mark.gpu


def zero_spread(df):
    df = df.copy()
    df['gpu_utilization_max'] = df['gpu_utilization_avg']
    return df


CASES = {
    'probe100_1x4x8': (lambda: zero_spread(tracegen.frame_probe100()), dict(num_switch=1, num_node_p_switch=4, num_gpu_p_node=8), 5),
    'probe100_2x2x8_k3': (lambda: zero_spread(tracegen.frame_probe100()), dict(num_switch=2, num_node_p_switch=2, num_gpu_p_node=8), 3),
    'gen300_2x4x8': (lambda: zero_spread(tracegen.frame_gen(300, 11, 150)), dict(num_switch=2, num_node_p_switch=4, num_gpu_p_node=8), 5),
    'gen300_3x2x4': (lambda: zero_spread(tracegen.frame_gen(300, 12, 60)), dict(num_switch=3, num_node_p_switch=2, num_gpu_p_node=4), 5),
    'gen2000_4x8x8_spread': (lambda: tracegen.frame_gen(2000, 13, 1000), dict(num_switch=4, num_node_p_switch=8, num_gpu_p_node=8), 5),
}


def run_device(df, flags, k, seed=None, schedule='horus', scheme=None, **kw):
    cluster = rl.cluster_from_flags(flags)
    tr = rl.prepare_trace(df, cluster)
    sim = rl.Simulator(cluster, schedule, scheme or schedule, n_replicas=2, rows=True, num_buffer=k, pack_seed=seed, max_ticks=400000, **kw)
    sim.load_trace(tr)
    sim.run()
    return sim, cluster, tr


def check(sim, cluster, tr, o, otr, replica):
    j = sim.jobs(replica)
    aux = sim.job_plane(replica, _ffi.PLANE_AUX)
    dur = tr.duration + 5.0 * (aux == 1)
    assert np.array_equal(j['finish_order'], o['finish_order'])
    assert np.array_equal(j['start'], o['start']) and np.array_equal(j['end'], o['end'])
    fo = o['finish_order']
    assert np.array_equal(dur[fo], o['actual_duration'][fo])   # Job.get_duration() is written when the job finishes
    jct = sim.job_plane(replica, _ffi.PLANE_PREEMPT)
    assert np.array_equal(jct[j['finish_order']], o['jct'][o['finish_order']]) and np.array_equal(j['preempt'][j['finish_order']], o['preempt'][o['finish_order']])
    got = lm.format_job_csv(tr, j['finish_order'], j['start'], j['end'], j['preempt'], get_duration=dur, jct=jct)
    assert got == cpu_sim.format_job_csv(otr, o)
    assert lm.format_cluster_csv(sim.rows(replica), cluster, tr.mem_shift, with_util=False) == cpu_sim.format_cluster_csv(o)


@pytest.mark.parametrize('schedule', ['horus', 'gandiva'])
@pytest.mark.parametrize('name', list(CASES))
def test_horus_matches_oracle_mean_draws(name, schedule):
    frame, flags, k = CASES[name]
    df = frame()
    sim, cluster, tr = run_device(df, flags, k, schedule=schedule)
    otr = cpu_sim.prepare_trace(df)
    o = cpu_sim.run_pack(cpu_sim.make_cluster(**flags), otr, schedule, k)
    for r in range(2):
        check(sim, cluster, tr, o, otr, r)
    sim.close()


@pytest.mark.parametrize('schedule', ['horus', 'gandiva'])
@pytest.mark.parametrize('name', ['probe100_1x4x8', 'gen300_2x4x8', 'gen2000_4x8x8_spread'])
def test_horus_matches_oracle_seeded_draws(name, schedule):
    frame, flags, k = CASES[name]
    df = tracegen.frame_probe100() if name.startswith('probe') else (tracegen.frame_gen(300, 11, 150) if name.startswith('gen300') else frame())
    sim, cluster, tr = run_device(df, flags, k, seed=1234, schedule=schedule)
    otr = cpu_sim.prepare_trace(df)
    for r in range(2):
        o = cpu_sim.run_pack(cpu_sim.make_cluster(**flags), otr, schedule, k, seed=1234, replica=r)
        check(sim, cluster, tr, o, otr, r)
    sim.close()


@pytest.mark.parametrize('schedule', ['horus', 'gandiva'])
def test_horus_bounded_launches_resume(schedule):
    frame, flags, k = CASES['gen300_2x4x8']
    df = frame()
    sim, cluster, tr = run_device(df, flags, k, ticks_per_launch=37, schedule=schedule)
    otr = cpu_sim.prepare_trace(df)
    o = cpu_sim.run_pack(cpu_sim.make_cluster(**flags), otr, schedule, k)
    check(sim, cluster, tr, o, otr, 1)
    sim.close()


@pytest.mark.parametrize('name', goldutil.pack_case_names())
def test_horus_matches_the_reference_golden(name):
    """Device outputs vs the files the UNMODIFIED reference wrote for `--schedule horus --scheme horus` (tests/golden)."""
    g = goldutil.load(name)
    cluster = rl.cluster_from_flags(g['flags'])
    tr = rl.prepare_trace(goldutil.trace_input(g), cluster)
    kw = dict(num_queue=g['num_queue'], pack_seed=g['inject_seed'], pack_rng=False) if g['schedule'] == 'horus+' else {}
    sim = rl.Simulator(cluster, g['schedule'], g['scheme'], n_replicas=3, rows=True, num_buffer=g['num_buffer'], max_ticks=400000, **kw)
    sim.load_trace(tr)
    sim.run()
    for r in (0, 2):
        j = sim.jobs(r)
        dur = tr.duration + 5.0 * (sim.job_plane(r, _ffi.PLANE_AUX) == 1)
        job = lm.format_job_csv(tr, j['finish_order'], j['start'], j['end'], j['preempt'], get_duration=dur,
                                jct=sim.job_plane(r, _ffi.PLANE_PREEMPT))
        clu = lm.format_cluster_csv(sim.rows(r), cluster, tr.mem_shift, with_util=False)
        if g['job'] is not None:
            assert job == g['job'] and clu == g['cluster']
        assert goldutil.sha(job) == g['meta']['job_sha256'] and goldutil.sha(clu) == g['meta']['cluster_noutil_sha256']
    sim.close()


@pytest.mark.parametrize('schedule', ['horus', 'gandiva'])
@pytest.mark.parametrize('name', ['probe100_1x4x8', 'gen300_3x2x4', 'gen2000_4x8x8_spread'])
def test_pack_schedules_over_yarn_match_oracle(name, schedule):
    """--schedule horus|gandiva --scheme yarn: no utilisation draw anywhere, traces keep their spread."""
    frame, flags, k = CASES[name]
    df = tracegen.frame_probe100() if name.startswith('probe') else (tracegen.frame_gen(300, 12, 60) if name.startswith('gen300') else frame())
    sim, cluster, tr = run_device(df, flags, k, schedule=schedule, scheme='yarn', ticks_per_launch=(0 if name.startswith('probe') else 53))
    otr = cpu_sim.prepare_trace(df)
    o = cpu_sim.run_pack(cpu_sim.make_cluster(**flags), otr, schedule, k, scheme='yarn')
    for r in range(2):
        check(sim, cluster, tr, o, otr, r)
    sim.close()


def test_pack_limits_are_reported():
    """More than 32 tasks per job is outside the pack kernels' lane-per-task layout: refused at load time, not at run time."""
    df = tracegen.frame_rows([dict(normalized_time=0, minutes=4, used_gpus=40.0, gpu_per_container=1)])
    cluster = rl.Cluster(num_switch=1, num_node_p_switch=8, num_gpu_p_node=8)
    tr = rl.prepare_trace(df, cluster)
    sim = rl.Simulator(cluster, 'horus', 'horus', n_replicas=1)
    with pytest.raises(_ffi.RlgsError) as e:
        sim.load_trace(tr)
    assert e.value.code == _ffi.ERR_UNSUPPORTED and '32' in str(e.value)
    sim.close()
    with pytest.raises(_ffi.RlgsError):
        rl.Simulator(cluster, 'fifo', 'horus')          # KeyError 'fifo' in the reference's score table
    sim = rl.Simulator(cluster, 'gandiva', 'yarn', n_replicas=1)
    with pytest.raises(_ffi.RlgsError):
        sim.run()                                        # no trace
    sim.close()


@pytest.mark.parametrize('scheme', ['horus+', 'yarn'])
@pytest.mark.parametrize('name,kq,seed', [('probe100_1x4x8', 3, 1), ('probe100_1x4x8', 1, 4), ('probe100_2x2x8_k3', 2, 5), ('gen300_2x4x8', 4, 2), ('gen300_3x2x4', 5, 3),
                                          ('gen2000_4x8x8_spread', 3, 7)])
def test_horus_plus_matches_oracle(name, kq, seed, scheme):
    """--schedule horus+: k-means queues + credit pick.  The oracle is pinned against the reference run with the same
    injected k-means draws (oracle/ref_runner.py _INJECT, tests/golden/horusplus_*)."""
    frame, flags, k = CASES[name]
    df = frame()
    cluster = rl.cluster_from_flags(flags)
    tr = rl.prepare_trace(df, cluster)
    sim = rl.Simulator(cluster, 'horus+', scheme, n_replicas=2, rows=True, num_buffer=k, num_queue=kq, pack_seed=seed, pack_rng=False,
                       max_ticks=400000, ticks_per_launch=(0 if kq != 4 else 41))
    sim.load_trace(tr)
    sim.run()
    otr = cpu_sim.prepare_trace(df)
    o = cpu_sim.run_pack(cpu_sim.make_cluster(**flags), otr, 'horus+', k, scheme=('yarn' if scheme == 'yarn' else None), num_queue=kq, inject_seed=seed)
    for r in range(2):
        check(sim, cluster, tr, o, otr, r)
    sim.close()


def test_pack_trace_reload_replaces_the_previous_one():
    """load_trace on the same replica range twice (the e2e pattern): the second trace is the one that runs."""
    frame, flags, k = CASES['probe100_2x2x8_k3']
    cluster = rl.cluster_from_flags(flags)
    sim = rl.Simulator(cluster, 'horus', 'horus', n_replicas=2, rows=True, num_buffer=k, max_ticks=400000)
    for df in (zero_spread(tracegen.frame_gen(300, 12, 60)), frame()):
        tr = rl.prepare_trace(df, cluster)
        sim.load_trace(tr)
        sim.run()
        otr = cpu_sim.prepare_trace(df)
        check(sim, cluster, tr, cpu_sim.run_pack(cpu_sim.make_cluster(**flags), otr, 'horus', k), otr, 1)
    sim.close()


def test_horus_plus_with_seeded_utilisation_draws():
    """horus+ on a trace with a utilisation spread: the k-means draws and the score's utilisation draws share opts.pack_seed."""
    frame, flags, k = CASES['gen2000_4x8x8_spread']
    df = frame()
    cluster = rl.cluster_from_flags(flags)
    tr = rl.prepare_trace(df, cluster)
    sim = rl.Simulator(cluster, 'horus+', 'horus+', n_replicas=2, rows=True, num_buffer=k, num_queue=3, pack_seed=5, pack_rng=True, max_ticks=400000)
    sim.load_trace(tr)
    sim.run()
    otr = cpu_sim.prepare_trace(df)
    for r in range(2):
        o = cpu_sim.run_pack(cpu_sim.make_cluster(**flags), otr, 'horus+', k, seed=5, replica=r, num_queue=3, inject_seed=5)
        check(sim, cluster, tr, o, otr, r)
    sim.close()
