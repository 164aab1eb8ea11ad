"""CPU tests of the host-side mirror: flags, plugin registries, CSV formatting, ingestion."""
import numpy as np
import pytest

import cpu_sim
import goldutil
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import _ffi, algorithm, flags, log_manager as lm


def test_flags_follow_the_reference_conventions():
    flags.reset_for_tests()
    for name, default, kind in (('t_int', 3, flags.DEFINE_integer), ('t_str', 'a', flags.DEFINE_string), ('t_f', 0.5, flags.DEFINE_float)):
        kind(name, default, 'doc')
    flags.DEFINE_boolean('t_bool', False, 'doc')
    unknown = flags.FLAGS.parse(['--t_int', '7', '--t_bool', 'True', '--not_a_flag', '1'])
    assert flags.FLAGS.t_int == 7 and flags.FLAGS.t_bool is True and flags.FLAGS.t_str == 'a' and flags.FLAGS.t_f == 0.5
    assert unknown == ['--not_a_flag', '1']            # parse_known_args: unknown flags are ignored (core/flags.py:22)
    flags.reset_for_tests()
    flags.FLAGS.parse(['--t_bool', 'False'])
    assert flags.FLAGS.t_bool is False                  # str2bool: only true/t/1 are True (core/flags.py:91-98)
    flags.reset_for_tests()
    flags.FLAGS.parse(['--t_bool'])
    assert flags.FLAGS.t_bool is True
    flags.reset_for_tests()
    flags.FLAGS.parse(['--t_bool', '--not_bool'])
    assert flags.FLAGS.t_bool is False                  # --no<flag> negation (core/flags.py:100-104)
    flags.FLAGS.t_int = 11
    assert flags.FLAGS.t_int == 11
    flags.reset_for_tests()


def test_plugin_registries_keep_the_reference_keys():
    assert {'fifo', 'horus', 'horus+', 'gandiva'} <= set(algorithm.scheduling_algorithms)   # algorithm.py:292-298
    assert {'yarn', 'horus', 'horus+', 'gandiva'} <= set(algorithm.placement_algorithms)    # algorithm.py:182-187
    assert set(algorithm.plugin_algorithms) == {'gandiva'} and set(algorithm.score_fn) == {'horus', 'horus+', 'gandiva'}
    s, p = algorithm.resolve('fifo', 'yarn')
    assert s.device_id == _ffi.SCHED['fifo'] and p.device_id == _ffi.PLACE['yarn']
    assert algorithm.resolve('dlas-gpu', 'count')[0].device_id == 2 and algorithm.resolve('dlas', 'count')[0].device_id == 3
    assert algorithm.resolve('horus', 'gandiva')[0].device_id == _ffi.SCHED['horus']   # three names, one placement
    assert algorithm.resolve('horus+', 'horus+')[0].device_id == _ffi.SCHED['horus+']
    with pytest.raises(NotImplementedError):
        algorithm.resolve('fifo', 'horus')              # KeyError 'fifo' in the reference (algorithm.py:58)
    assert algorithm.resolve('horus', 'yarn')[1].device_id == _ffi.PLACE['yarn'] and algorithm.resolve('gandiva', 'yarn')[0].device_id == _ffi.SCHED['gandiva']
    assert algorithm.resolve('gandiva', 'gandiva')[0].device_id == _ffi.SCHED['gandiva']
    with pytest.raises(KeyError):
        algorithm.resolve('lpjf', 'yarn')
    algorithm.scheduling_algorithms['mine'] = lambda *a, **k: (None, None, False)   # a user-registered callable: executable over yarn
    assert algorithm.resolve('mine', 'yarn')[0] is algorithm.scheduling_algorithms['mine']
    with pytest.raises(NotImplementedError):
        algorithm.resolve('mine', 'gandiva')
    del algorithm.scheduling_algorithms['mine']


@pytest.mark.parametrize('name', ['probe100', 'big_mem_leak', 'dense'])
def test_row_finishing_reproduces_the_reference_floats(name):
    """Feeds the oracle's per-tick state through the product's integer row format and formatter:
    the float columns must come out byte-identical to the reference's cluster.csv."""
    g = goldutil.load(name)
    ti = goldutil.trace_input(g)
    cluster = rl.cluster_from_flags(g['flags'])
    tr = rl.prepare_trace(ti, cluster)
    otr = cpu_sim.prepare_trace(ti)
    o = cpu_sim.run_fifo_yarn(cpu_sim.make_cluster(**g['flags']), otr)
    # rebuild integer sufficient statistics from the oracle run
    n = o['n_ticks']
    rows = np.zeros(n, _ffi.ROW_DTYPE)
    arr = np.ceil(otr['nt']).astype(np.int64)
    start, end = o['start'].astype(np.int64), o['end'].astype(np.int64)
    ndev = tr.records['tasks'].astype(np.int64) * tr.records['gpus_per_task']
    for d in range(1, n + 1):
        queued = np.nonzero((arr < d) & ((start < 0) | (start >= d)))[0]
        running = (start >= 0) & (start < d) & ((end < 0) | (end > d))
        pend = np.sort(d - arr[queued])
        r = rows[d - 1]
        r['queued'] = len(queued); r['running'] = running.sum(); r['finished'] = ((end >= 0) & (end <= d)).sum()
        r['busy_gpus'] = ndev[running].sum(); r['mem_sum'] = tr.records['mem_term'][running].sum()
        r['sum_pending'] = pend.sum()
        if len(pend):
            r['median_lo'] = pend[(len(pend) - 1) // 2]; r['median_hi'] = pend[len(pend) // 2]; r['max_pending'] = pend[-1]
        r['idle_nodes'] = o['rows']['idle_nodes'][d - 1]
    assert np.array_equal(rows['busy_gpus'], o['rows']['busy_gpus']) and np.array_equal(rows['queued'], o['rows']['queued'])
    assert lm.format_cluster_csv(rows, cluster, tr.mem_shift, with_util=False) == g['cluster']
    assert lm.format_job_csv(tr, o['finish_order'], o['start'], o['end']) == g['job']


def test_log_manager_creates_the_six_files_with_reference_headers(tmp_path):
    class F(object):
        scheme = 'yarn'; schedule = 'fifo'
    cluster = rl.Cluster(num_switch=1, num_node_p_switch=2, num_gpu_p_node=2)
    m = lm.LogManager(str(tmp_path), F())
    m.init(cluster)
    assert open(tmp_path / 'cluster.csv', newline='').read() == ','.join(lm.CLUSTER_HEADER) + '\r\n'
    assert open(tmp_path / 'job.csv', newline='').read() == ','.join(lm.JOB_HEADER) + '\r\n'
    assert open(tmp_path / 'cpu.csv', newline='').read() == 'time,cpu0,cpu1\r\n'
    assert open(tmp_path / 'gpu.csv', newline='').read() == 'time,gpu0,gpu1,gpu2,gpu3\r\n'
    assert open(tmp_path / 'memory.csv', newline='').read() == 'time,max,99th,95th,med\r\n'
    assert open(tmp_path / 'network.csv', newline='').read() == 'time,in0,out0,in1,out1\r\n'
    m.step_cluster(lm.LogInfo(1, 1, 2, 2, 0.0, 0.5, 0.0, float('nan'), 0, 1, 0, 0), 1)
    assert open(tmp_path / 'cluster.csv', newline='').read().split('\r\n')[1] == '1,1,1,2,2,0.0,0.5,0.0,nan,0,1,0,0'


def test_ingest_rejects_what_the_reference_would_raise_on():
    import pandas as pd
    import tracegen
    cluster = rl.Cluster()
    df = tracegen.frame_rows([dict(used_gpus=1.0, gpu_per_container=2)])
    with pytest.raises(ValueError):
        rl.prepare_trace(df, cluster)                    # zero tasks: StopIteration at node.py:118 in the reference
    df = tracegen.frame_rows([dict(used_gpus=2.0, gpu_per_container=1, memory_max=5000000001.5)])
    t = rl.prepare_trace(df, cluster)
    assert t.mem_shift == 21 and int(t.records['mem_term'][0]) == 2 * int(5000000001.5 * 2)
    df = tracegen.frame_rows([dict(memory_max=1e9 / 3)])
    with pytest.raises(ValueError):
        rl.prepare_trace(df, cluster)                    # not exactly summable in 53 bits


def test_pack_inputs_follow_the_task_fields_the_scores_read():
    """Trace.pack_inputs: utilisation mean / half spread, un-clamped task memory in the trace's dyadic unit, floor(used_gpus)."""
    from rlgpuschedule_b200 import synth
    df = synth.frame_rows([dict(used_gpus=4.0, gpu_per_container=2, gpu_utilization_avg=40.0, gpu_utilization_max=50.0, memory_max=5 * 2 ** 30, memory_avg=3 * 2 ** 30),
                           dict(normalized_time=10000, used_gpus=3.0, gpu_per_container=2, gpu_utilization_avg=0.0, gpu_utilization_max=0.0, memory_max=40 * 10 ** 9, memory_avg=1e9)])
    tr = rl.prepare_trace(df, rl.Cluster(num_gpu_p_node=8))
    pi = tr.pack_inputs()
    assert pi['util_sd'].tolist() == [5.0, 0.0] and pi['heap_cap'].tolist() == [4, 3]
    unit = 2.0 ** tr.mem_shift
    assert pi['task_mem'].tolist() == [int(5 * 1024 * unit), int(40e9 / 2 ** 20 * unit)]        # not clamped to the 32 GiB capacity
    assert 40e9 / 2 ** 20 * unit == float(int(40e9 / 2 ** 20 * unit))                             # exact in the unit
    assert tr.records['tasks'].tolist() == [2, 1] and pi['used_gpus'].tolist() == [4.0, 3.0]
    assert np.allclose(pi['mem_avg_mib'], [3 * 1024, 1e9 / 2 ** 20])


def test_plugin_views_and_host_forms_of_the_builtin_entries():
    """The read-only views handed to a user-registered scheduling callable (rlgpuschedule_b200/plugin.py), built from an
    observation vector laid out like the kernel writes it; no GPU involved."""
    import numpy as np
    import tracegen
    import rlgpuschedule_b200 as rl
    from rlgpuschedule_b200 import algorithm, plugin
    cluster = rl.Cluster(num_switch=2, num_node_p_switch=2, num_gpu_p_node=4)
    tr = rl.prepare_trace(tracegen.frame_gen(20, 3, 20), cluster)
    N, K = 4, 3
    obs = np.zeros(3 * N + 5 * K + 4, np.float32)
    obs[0:N] = [0, 2, 4, 1]; obs[N:2 * N] = [128 - 48, 128 - 24, 128, 0]; obs[2 * N:3 * N] = [512 - 240, 512 - 120, 512, 0]
    for i, (job, pend) in enumerate(((7, 3), (5, 9))):
        rec = tr.records[job]
        obs[3 * N + 5 * i: 3 * N + 5 * i + 5] = [rec['gpus'], rec['tasks'], rec['dur_ticks'], pend, job]
    obs[3 * N + 5 * 2 + 4] = -1
    arrived = int((tr.records['arrival_tick'] <= 17).sum())          # jobs the trace has delivered by tick 17
    assert arrived >= 7
    obs[3 * N + 5 * K:] = [2, 3, arrived - 5, 17]
    infra = plugin.InfrastructureView(cluster, None, obs)
    jm = plugin.JobsManagerView(tr, obs, N, K)
    assert list(infra.nodes) == ['1', '2', '3', '4'] and infra.nodes['3'].cpu_free() == 128 and infra.nodes['2'].rack_id == '0'
    assert len(infra.nodes['2'].get_free_devices()) == 2 and not infra.nodes['4'].is_free() and infra.num_free_nodes() == 3
    assert jm.delta == 17 and jm.queuing_jobs() == 2 and [j.trace_index for j in jm.window()] == [7, 5]
    head = jm.get_next_job(17)
    assert head.job_id == str(int(tr.label[7])) and head.pending_time == 3 and head.gpus == tr.used_gpus[7]
    nodes, job, ok = algorithm.scheduling_algorithms['fifo']('yarn', algorithm.placement_algorithms['yarn'], infra, jm, 17, k=K)
    need = int(np.ceil(head.gpus))
    if need <= 4:
        first = next(nid for nid, n in infra.nodes.items() if len(n.get_free_devices()) >= need and n.cpu_free() >= 12 * head.task_count and n.mem_free() >= 60 * head.task_count)
        assert ok and job is head and list(nodes) == [first] and jm.popped is head
    # jobs that arrive at this tick join the FRONT of the window before the callable runs (schedule.py:187-190, q1)
    obs2 = obs.copy(); obs2[3 * N + 5 * K + 2] -= 2
    jm2 = plugin.JobsManagerView(tr, obs2, N, K)
    assert jm2.queuing_jobs() == 4 and [j.trace_index for j in jm2.window()] == [arrived - 2, arrived - 1, 7]
    assert [j.window_index for j in jm2.window()] == [0, 1, 2] and jm2.window()[0].pending_time == 0
    with pytest.raises(RuntimeError):
        algorithm.scheduling_algorithms['horus']('horus', None, infra, jm, 17)
    algorithm.scheduling_algorithms['user'] = lambda *a, **k: (None, None, False)
    try:
        assert algorithm.resolve('user', 'yarn')[0] is algorithm.scheduling_algorithms['user']
        with pytest.raises(NotImplementedError):
            algorithm.resolve('user', 'horus')
    finally:
        del algorithm.scheduling_algorithms['user']


@pytest.mark.parametrize('name', goldutil.case_names('small'))
def test_start_bits_and_the_trace_determine_the_run(name):
    """The rule behind RLGS_ROWFMT_EVENT4 (rlgs_row4e, include/rlgs.h), checked against the reference's own files: given only
    idle_nodes and "a job started at this tick" per tick, a replay of the queue (arrivals pushed to the FRONT in trace order,
    the attempt pops the front) names every started job, and end = start + dur_ticks, finish order = (end, start) and the
    pending times at fixed queue positions reproduce job.csv and cluster.csv byte for byte."""
    g = goldutil.load(name)
    if g['flags'].get('enable_network_costs'):
        pytest.skip('network costs change the duration at placement time')
    ti = goldutil.trace_input(g)
    cluster = rl.cluster_from_flags(g['flags'])
    tr = rl.prepare_trace(ti, cluster)
    o = cpu_sim.run_fifo_yarn(cpu_sim.make_cluster(**g['flags']), cpu_sim.prepare_trace(ti))
    n, rec = o['n_ticks'], tr.records
    J = len(rec)
    bits = np.zeros(n, bool)
    bits[o['start'][o['start'] >= 0]] = True                      # all that the rows say about starts
    arr = rec['arrival_tick'].astype(np.int64)
    ndev = rec['tasks'].astype(np.int64) * rec['gpus_per_task']
    start, end = np.full(J, -1, np.int64), np.full(J, -1, np.int64)
    queue, cursor, back_arr = [], 0, 0
    rows = np.zeros(n, _ffi.ROW_DTYPE)
    for i in range(n):
        k = 0
        while cursor + k < J and arr[cursor + k] <= i:
            k += 1
        if k:
            if not queue:
                back_arr = i
            queue[0:0] = list(range(cursor, cursor + k))
            cursor += k
        if bits[i]:
            j = queue.pop(0)
            start[j] = i
            if i + rec['dur_ticks'][j] <= n:
                end[j] = i + rec['dur_ticks'][j]
        r, d, Q = rows[i], i + 1, len(queue)
        r['queued'] = Q
        if Q:
            r['max_pending'] = d - back_arr
            r['median_lo'] = d - arr[queue[(Q - 1) // 2]]; r['median_hi'] = d - arr[queue[Q // 2]]
            r['sum_pending'] = sum(d - arr[q] for q in queue)
        running = (start >= 0) & ((end < 0) | (end > d))
        r['running'] = running.sum(); r['finished'] = ((end >= 0) & (end <= d)).sum()
        r['busy_gpus'] = ndev[running].sum(); r['mem_sum'] = rec['mem_term'][running].sum()
        r['idle_nodes'] = o['rows']['idle_nodes'][i]
    fin = sorted((int(end[j]), int(start[j]), j) for j in range(J) if end[j] >= 0)
    order = np.array([j for _, _, j in fin], np.int32)
    assert np.array_equal(start, o['start']) and np.array_equal(end, o['end']) and np.array_equal(order, o['finish_order'])
    assert lm.format_cluster_csv(rows, cluster, tr.mem_shift, with_util=False) == g['cluster']
    assert lm.format_job_csv(tr, order, start.astype(np.int32), end.astype(np.int32)) == g['job']
