"""Pins oracle/cpu_sim.c (the CPU restatement) against the UNMODIFIED reference's own outputs
(tests/golden, written by oracle/make_golden.py): job.csv byte-for-byte and cluster.csv minus the
unseeded-RNG column byte-for-byte."""
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
