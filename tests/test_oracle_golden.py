"""Pins oracle/cpu_sim.c (the CPU restatement) against the UNMODIFIED reference's own outputs
(tests/golden, written by oracle/make_golden.py): job.csv byte-for-byte and cluster.csv minus the
unseeded-RNG column byte-for-byte.  The legacy event loops (sjf / shortest / shortest-gpu / dlas-gpu / dlas)
are pinned against the files the reference's own log._Log wrote while its dead-code loops ran unmodified
under shim globals (oracle/ref_legacy_runner.py)."""
import pytest

import cpu_sim
import goldutil


def _run(name):
    g = goldutil.load(name)
    tr = cpu_sim.prepare_trace(goldutil.trace_input(g))
    res = cpu_sim.run_fifo_yarn(cpu_sim.make_cluster(**g['flags']), tr)
    return g, tr, res


@pytest.mark.parametrize('name', goldutil.case_names('small'))
def test_oracle_matches_reference_small(name):
    g, tr, res = _run(name)
    assert cpu_sim.format_job_csv(tr, res) == g['job']
    assert cpu_sim.format_cluster_csv(res) == g['cluster']
    assert res['n_ticks'] == g['meta']['n_ticks']


@pytest.mark.parametrize('name', goldutil.case_names('big') + goldutil.case_names('huge'))
def test_oracle_matches_reference_big(name):
    g, tr, res = _run(name)
    assert goldutil.sha(cpu_sim.format_job_csv(tr, res)) == g['meta']['job_sha256']
    assert goldutil.sha(cpu_sim.format_cluster_csv(res)) == g['meta']['cluster_noutil_sha256']


# ---- horus schedule + horus_placement (oracle_pack) on zero-spread traces
def _run_pack(name):
    g = goldutil.load(name)
    tr = cpu_sim.prepare_trace(goldutil.trace_input(g))
    res = cpu_sim.run_pack(cpu_sim.make_cluster(**g['flags']), tr, g['schedule'], g['num_buffer'], scheme=g['scheme'],
                           num_queue=g['num_queue'], inject_seed=g['inject_seed'])
    return g, tr, res


@pytest.mark.parametrize('name', goldutil.pack_case_names(('small',)))
def test_pack_oracle_matches_reference_small(name):
    g, tr, res = _run_pack(name)
    assert cpu_sim.format_job_csv(tr, res) == g['job']
    assert cpu_sim.format_cluster_csv(res) == g['cluster']
    assert res['n_ticks'] == g['meta']['n_ticks']


@pytest.mark.parametrize('name', goldutil.pack_case_names(('big', 'huge')))
def test_pack_oracle_matches_reference_big(name):
    g, tr, res = _run_pack(name)
    assert goldutil.sha(cpu_sim.format_job_csv(tr, res)) == g['meta']['job_sha256']
    assert goldutil.sha(cpu_sim.format_cluster_csv(res)) == g['meta']['cluster_noutil_sha256']


# ---- legacy event loops: run_sim.py:162-287, :299-431, :664-947 executed unmodified (oracle/ref_legacy_runner.py)
def _run_legacy(name):
    g = goldutil.load(name)
    tr = cpu_sim.prepare_trace(goldutil.trace_input(g))
    cl = cpu_sim.make_cluster(**g['flags'])
    res, count = cpu_sim.run_legacy(cl, tr, g['schedule'], g['queue_limit'] or (30, 60, 150))
    return g, cpu_sim.format_legacy_job_csv(tr, res, count), cpu_sim.format_legacy_cluster_csv(res, cl, count), res


@pytest.mark.parametrize('name', goldutil.legacy_case_names(('small', 'big')))
def test_legacy_oracle_matches_reference(name):
    g, job, clu, res = _run_legacy(name)
    assert job == g['job']
    assert clu == g['cluster']
    assert res['n_events'] == g['meta']['n_events']


@pytest.mark.parametrize('name', goldutil.legacy_case_names(('huge',)))
def test_legacy_oracle_matches_reference_huge(name):
    g, job, clu, res = _run_legacy(name)
    assert goldutil.sha(job) == g['meta']['job_sha256']
    assert goldutil.sha(clu) == g['meta']['cluster_sha256']


def test_dlas_fixture_exercises_the_attached_end_list():
    """run_sim.py:706-717: a start event keeps the 'end_jobs' list written into it when a queue jump replaced the event; the
    fixture must contain a job completed through that stale list while it was PENDING (a job that ends normally has
    resume == preempt + 1)."""
    g = goldutil.load('dlasgpu_multi_node')
    rows = [r.split(',') for r in g['job'].split('\r\n')[1:] if r]
    assert any(int(r[10]) >= int(r[11]) for r in rows)   # preempt >= resume: completed while preempted
