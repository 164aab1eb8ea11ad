"""The fifo tick loop with 8 / 16 / 32 lanes per replica (4 / 2 / 1 replicas per warp) and both row formats.
Every variant must give the same bytes: against the reference's golden files, against the oracle on replicas that share a
warp while following DIFFERENT traces (groups diverge), across bounded launches (state save / restore), and the 16-byte wire
rows expanded on the host must equal the self-contained 64-byte rows the device writes in 'wide' mode."""
import numpy as np
import pytest

import cpu_sim
import goldutil
import tracegen
import rlgpuschedule_b200 as rl
from rlgpuschedule_b200 import _ffi, log_manager as lm

pytestmark = pytest.mark.gpu

LPRS = (8, 16, 32)
FORMATS = ('event4', 'event16', 'wire12', 'wire16', 'wide')


def _csvs(sim, cluster, tr, r):
    j = sim.jobs(r)
    return (lm.format_job_csv(tr, j['finish_order'], j['start'], j['end'], j['preempt']),
            lm.format_cluster_csv(sim.rows(r), cluster, tr.mem_shift, with_util=False))


@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('lpr', LPRS)
@pytest.mark.parametrize('name', ['kat6', 'multi_node', 'big_mem_leak', 'ties', 'dense', 'gpu_cap16', 'cluster_spec', 'nondivisible'])
def test_every_variant_matches_the_reference_files(name, lpr, fmt):
    g = goldutil.load(name)
    cluster = rl.cluster_from_flags(g['flags'])
    tr = rl.prepare_trace(goldutil.trace_input(g), cluster)
    sim = rl.Simulator(cluster, 'fifo', 'yarn', n_replicas=5, rows=True, lanes_per_replica=lpr, rows_format=fmt)
    sim.load_trace(tr)
    sim.run()
    for r in (0, 4):
        job, clu = _csvs(sim, cluster, tr, r)
        assert job == g['job'] and clu == g['cluster']
    sim.close()


@pytest.mark.parametrize('lpr', LPRS)
def test_replicas_of_one_warp_follow_different_traces(lpr):
    """11 replicas over 5 traces of different lengths: the groups of a warp diverge, finish at different ticks, and one warp is
    partly empty.  Bounded launches on top (state leaves and re-enters shared memory every 37 ticks)."""
    flags = dict(num_switch=2, num_node_p_switch=5, num_gpu_p_node=8)
    cluster = rl.cluster_from_flags(flags)
    frames = [tracegen.frame_gen(150 + 40 * i, 30 + i, 60 + 50 * i) for i in range(5)]
    traces = [rl.prepare_trace(f, cluster) for f in frames]
    layout = [(0, 3, 0), (3, 2, 1), (5, 1, 2), (6, 3, 3), (9, 2, 4)]
    oc = cpu_sim.make_cluster(**flags)
    ores = [cpu_sim.run_fifo_yarn(oc, cpu_sim.prepare_trace(f)) for f in frames]
    for kw in (dict(), dict(ticks_per_launch=37, rows_cap=64)):
        sim = rl.Simulator(cluster, 'fifo', 'yarn', n_replicas=11, rows=True, lanes_per_replica=lpr, **kw)
        for first, count, t in layout:
            sim.load_trace(traces[t], first, count)
        sim.run()
        for first, count, t in layout:
            for r in range(first, first + count):
                j = sim.jobs(r)
                o = ores[t]
                assert np.array_equal(j['finish_order'], o['finish_order']) and np.array_equal(j['start'], o['start']) and np.array_equal(j['end'], o['end'])
                assert lm.format_cluster_csv(sim.rows(r), cluster, traces[t].mem_shift, with_util=False) == cpu_sim.format_cluster_csv(o)
                s = sim.summary(r)
                assert s['sum_queued'] == o['counters']['sum_queued'] and s['sum_running'] == o['counters']['sum_running']
        sim.close()


@pytest.mark.parametrize('lpr', LPRS)
def test_wire_rows_expand_to_the_wide_rows(lpr):
    flags = dict(num_switch=2, num_node_p_switch=4, num_gpu_p_node=8)
    cluster = rl.cluster_from_flags(flags)
    tr = rl.prepare_trace(tracegen.frame_gen(600, 9, 90), cluster)
    rows = {}
    for fmt in FORMATS:
        sim = rl.Simulator(cluster, 'fifo', 'yarn', n_replicas=6, rows='device', lanes_per_replica=lpr, rows_format=fmt)
        sim.load_trace(tr)
        sim.run()
        rows[fmt] = sim.rows(3)
        if fmt == 'wire16':
            w = sim.rows_wire(3)['w']
            assert np.array_equal(w[:, 0] & 0xfff, rows[fmt]['idle_nodes']) and np.array_equal(w[:, 0] >> 12, rows[fmt]['finished'])
            assert np.array_equal(sim.rows_chunk_view(3, 0)['w'], w[:4096])
        if fmt == 'event4':
            w = sim.rows_wire(3)['w']
            assert w.ndim == 1 and np.array_equal(w & 0xfff, rows[fmt]['idle_nodes']) and np.array_equal(w >> 13, rows[fmt]['queued'] & 0x7ffff)
            assert np.array_equal(sim.rows_chunk_view(3, 0)['w'], w[:4096])
        if fmt == 'wire12':
            w = sim.rows_wire(3)['w']
            assert w.shape[1] == 3 and np.array_equal(w[:, 0] & 0xffffff, rows[fmt]['max_pending']) and np.array_equal(w[:, 2], rows[fmt]['median_hi'])
            part = np.zeros(50, _ffi.ROW_DTYPE)          # a sub-range: the cumulative counts must not depend on where reading starts
            _ffi.check(_ffi.lib().rlgs_read_rows(sim._h, 3, 100, 50, part.ctypes.data))
            assert np.array_equal(part, rows[fmt][100:150])
        sim.close()
    assert all(rows[fmt].dtype == _ffi.ROW_DTYPE for fmt in FORMATS)
    for f in _ffi.ROW_DTYPE.names:
        for fmt in FORMATS[:-1]:
            assert np.array_equal(rows['wide'][f], rows[fmt][f]), (fmt, f)
    assert rows['wide']['busy_gpus'].max() > 0 and rows['wide']['sum_pending'].max() > 0 and rows['wide']['util_var_sum'].max() > 0


def test_trace_reloads_small_large_small_and_overlapping_ranges():
    """A reload may only reuse device buffers the replicas still point at (round-1 advisor finding)."""
    flags = dict(num_switch=1, num_node_p_switch=4, num_gpu_p_node=8)
    cluster = rl.cluster_from_flags(flags)
    frames = {n: tracegen.frame_gen(n, 50 + n, max(20, n // 2)) for n in (100, 200, 50, 10)}
    traces = {n: rl.prepare_trace(f, cluster) for n, f in frames.items()}
    oc = cpu_sim.make_cluster(**flags)
    ores = {n: cpu_sim.run_fifo_yarn(oc, cpu_sim.prepare_trace(f)) for n, f in frames.items()}

    def check(sim, expect):
        sim.run()
        for r, n in enumerate(expect):
            j = sim.jobs(r)
            assert len(j['start']) == n
            assert np.array_equal(j['end'], ores[n]['end']) and np.array_equal(j['finish_order'], ores[n]['finish_order'])
            assert lm.format_cluster_csv(sim.rows(r), cluster, traces[n].mem_shift, with_util=False) == cpu_sim.format_cluster_csv(ores[n])
    sim = rl.Simulator(cluster, 'fifo', 'yarn', n_replicas=4, rows=True)
    sim.load_trace(traces[100]); check(sim, [100] * 4)
    sim.load_trace(traces[200]); check(sim, [200] * 4)
    sim.load_trace(traces[50]); check(sim, [50] * 4)                       # small after large on the same range
    sim.load_trace(traces[100], 0, 4); sim.load_trace(traces[10], 0, 2)     # [0,2) re-pointed at a short trace
    check(sim, [10, 10, 100, 100])
    sim.load_trace(traces[100], 0, 4); check(sim, [100] * 4)                # the full range again: must not reuse the 10-job buffers
    sim.load_trace(traces[200], 1, 2); check(sim, [100, 200, 200, 100])
    sim.close()


def test_env_step_before_reset_is_a_state_error():
    import torch
    cluster = rl.Cluster(num_switch=1, num_node_p_switch=2, num_gpu_p_node=4)
    tr = rl.prepare_trace(tracegen.frame_gen(40, 3, 40), cluster)
    sim = rl.Simulator(cluster, 'fifo', 'yarn', n_replicas=2, rows=False)
    sim.load_trace(tr)
    obs = torch.zeros(2, 3 * 2 + 5 * 5 + 4, device='cuda'); rew = torch.zeros(2, device='cuda'); done = torch.zeros(2, dtype=torch.uint8, device='cuda')
    rc = _ffi.lib().rlgs_env_step(sim._h, None, obs.data_ptr(), rew.data_ptr(), done.data_ptr(), 0, 5, 0, 1)
    assert rc == _ffi.ERR_STATE
    assert _ffi.lib().rlgs_env_reset(sim._h) == _ffi.OK
    assert _ffi.lib().rlgs_env_step(sim._h, None, obs.data_ptr(), rew.data_ptr(), done.data_ptr(), 0, 5, 0, 1) == _ffi.OK
    assert _ffi.lib().rlgs_env_sync(sim._h) == _ffi.OK
    sim.load_trace(tr)                                                          # a new trace invalidates the environment state
    assert _ffi.lib().rlgs_env_step(sim._h, None, obs.data_ptr(), rew.data_ptr(), done.data_ptr(), 0, 5, 0, 1) == _ffi.ERR_STATE
    sim.close()


def test_slot_overflow_has_its_own_status_code():
    cluster = rl.Cluster(num_switch=1, num_node_p_switch=8, num_gpu_p_node=8)
    df = tracegen.frame_gen(300, 4, 20)
    df['used_gpus'] = 1.0; df['gpu_per_container'] = 1
    df['minutes'] = 800.0
    tr = rl.prepare_trace(df, cluster)
    sim = rl.Simulator(cluster, 'fifo', 'yarn', n_replicas=1, rows=False, slot_cap=32, max_ticks=100000)
    sim.load_trace(tr)
    assert _ffi.lib().rlgs_run(sim._h) == _ffi.ERR_SLOTS                        # also with max_ticks set (round-1 advisor finding)
    sim.run()                                                                   # the wrapper regrows the table from the code
    assert sim.summary(0)['max_running'] > 32
    sim.close()


@pytest.mark.parametrize('fmt', FORMATS)
def test_end_only_job_tables_derive_the_start_ticks(fmt):
    """fetch_jobs='end' moves end ticks and the finish order to the host inside run(); a finished fifo job started at
    end - dur_ticks, which jobs() and the row expansion then use.  Must equal the three-table path and the oracle."""
    flags = dict(num_switch=2, num_node_p_switch=4, num_gpu_p_node=8)
    cluster = rl.cluster_from_flags(flags)
    df = tracegen.frame_gen(500, 13, 70)
    tr = rl.prepare_trace(df, cluster)
    o = cpu_sim.run_fifo_yarn(cpu_sim.make_cluster(**flags), cpu_sim.prepare_trace(df))
    sim = rl.Simulator(cluster, 'fifo', 'yarn', n_replicas=5, rows=True, rows_format=fmt, fetch_jobs='end')
    sim.load_trace(tr)
    sim.run()
    for r in (0, 4):
        j = sim.jobs(r)
        assert np.array_equal(j['start'], o['start']) and np.array_equal(j['end'], o['end']) and np.array_equal(j['finish_order'], o['finish_order'])
        assert lm.format_cluster_csv(sim.rows(r), cluster, tr.mem_shift, with_util=False) == cpu_sim.format_cluster_csv(o)
    sim.close()


@pytest.mark.parametrize('fmt', ['event16', 'event4'])
@pytest.mark.parametrize('lpr', LPRS)
def test_event_rows_rebuild_the_job_tables(lpr, fmt):
    """RLGS_ROWFMT_EVENT16: every row names the job that started at its tick; RLGS_ROWFMT_EVENT4: every row says whether the
    queue head started at its tick and the host replays the queue.  Either way the row stream is the event log: the tables
    rebuilt from it on the host (start, end = start + dur_ticks, finish order = (end, start)) must equal the tables the device
    wrote itself, on a trace with many same-tick finishes and jobs that never start, and the expanded rows (EVENT4: pending
    times out of the replayed queue) must print the oracle's cluster.csv."""
    flags = dict(num_switch=2, num_node_p_switch=4, num_gpu_p_node=8)
    cluster = rl.cluster_from_flags(flags)
    df = tracegen.frame_gen(700, 21, 400)
    df.loc[df.index[::3], 'minutes'] = 6.0                      # many equal durations: several jobs finish at the same tick
    df.loc[df.index[-5], 'used_gpus'] = 128.0; df.loc[df.index[-5], 'gpu_per_container'] = 8   # wider than the cluster: blocks the queue for good
    tr = rl.prepare_trace(df, cluster)
    o = cpu_sim.run_fifo_yarn(cpu_sim.make_cluster(**flags), cpu_sim.prepare_trace(df))
    assert (np.diff(o['end'][o['finish_order']]) == 0).sum() > 20 and (o['start'] < 0).sum() > 20   # the trace does what the docstring says
    tables = {}
    for fetch in (False, True):
        sim = rl.Simulator(cluster, 'fifo', 'yarn', n_replicas=5, rows=True, lanes_per_replica=lpr, rows_format=fmt, fetch_jobs=fetch)
        sim.load_trace(tr)
        sim.run()
        tables[fetch] = sim.jobs(4)
        w = sim.rows_wire(4)['w']
        if fmt == 'event16':
            ticks = np.nonzero(w[:, 3])[0]
            assert np.array_equal(o['start'][w[ticks, 3] - 1], ticks)
        else:
            ticks = np.nonzero(w & 0x1000)[0]
            assert np.array_equal(np.sort(o['start'][o['start'] >= 0]), ticks)
        assert len(ticks) == (o['start'] >= 0).sum()
        assert lm.format_cluster_csv(sim.rows(4), cluster, tr.mem_shift, with_util=False) == cpu_sim.format_cluster_csv(o)
        part = np.zeros(300, _ffi.ROW_DTYPE)                      # a range that starts mid-run: the cumulative counts catch up
        _ffi.check(_ffi.lib().rlgs_read_rows(sim._h, 4, 1000, 300, part.ctypes.data))
        assert np.array_equal(part, sim.rows(4)[1000:1300])
        assert np.array_equal(sim.jobs(4)['start'], tables[fetch]['start'])   # the expansion left the tables alone
        sim.close()
    for k in ('start', 'end', 'finish_order', 'preempt'):
        assert np.array_equal(tables[False][k], tables[True][k]), k
    assert np.array_equal(tables[False]['finish_order'], o['finish_order']) and np.array_equal(tables[False]['end'], o['end'])
