#!/usr/bin/env python
"""Sequential experiment sweep over run_sim.py, the role execute.py plays in the reference
(execute.py:5-55 there): every (placement, schedule, queues, look-ahead) combination is run `--repeats`
times as its own child process and logs under log/thesis_fitted_<k>_nodes_p_s<N>_job_<trace>/<scheme>_<schedule>/.

Default sweep = one run of every schedule family of the device path (the reference's horus+ / horus / gandiva / fifo plus
sjf and dlas-gpu).  `--reference-sweep` replays the reference's own list (horus+ with 3, 4, 5 queues, horus, gandiva,
yarn/fifo, each with look-ahead 15, 15, 15, 1, 1, 1).
"""
import argparse
import os
import subprocess
import sys
from collections import namedtuple

Run = namedtuple('Run', 'scheme schedule num_queue num_buffer')
HERE = os.path.dirname(os.path.abspath(__file__))

DEVICE_SWEEP = [Run('horus+', 'horus+', 3, 15), Run('horus', 'horus', 1, 15), Run('horus', 'horus', 1, 1), Run('gandiva', 'gandiva', 1, 1), Run('yarn', 'fifo', 1, 1),
                Run('yarn', 'sjf', 1, 1), Run('count', 'dlas-gpu', 4, 1)]
REFERENCE_SWEEP = [Run(s, s, q, b) for s, q in (('horus+', 3), ('horus+', 4), ('horus+', 5), ('horus', 1), ('gandiva', 1))
                   for b in (15, 15, 15, 1, 1, 1)] + [Run('yarn', 'fifo', 1, b) for b in (15, 15, 15, 1, 1, 1)]


def log_dir_for(run, trace_name, nodes_per_switch):
    parts = ['thesis_fitted_%d_nodes_p_s%d_job_%s' % (run.num_buffer, nodes_per_switch, trace_name),
             '%s_%s' % (run.scheme, run.schedule)]
    if run.schedule == 'horus+':
        parts.append('k%d' % run.num_queue)
    return os.path.join(*parts)


def launch(run, opts):
    argv = [sys.executable, os.path.join(HERE, 'run_sim.py'),
            '--scheme', run.scheme, '--schedule', run.schedule,
            '--num_queue', str(run.num_queue), '--num_buffer', str(run.num_buffer),
            '--num_switch', str(opts.num_switch), '--num_node_p_switch', str(opts.num_node_p_switch),
            '--trace_file', os.path.join(opts.data_dir, opts.trace + '.csv'),
            '--enable_network_costs', 'False', '--enable_migration', 'True',
            '--log_path', log_dir_for(run, opts.trace, opts.num_node_p_switch)]
    child = subprocess.Popen(argv)
    print('process pid %d: %s' % (child.pid, ' '.join(argv[2:])), flush=True)
    try:
        return child.wait()
    except KeyboardInterrupt:
        child.kill()
        raise


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('--reference-sweep', action='store_true', help="the reference's own combination list")
    ap.add_argument('--trace', default='month', help='trace name: <data-dir>/<trace>.csv')
    ap.add_argument('--data-dir', default='data')
    ap.add_argument('--repeats', type=int, default=3)
    ap.add_argument('--num_switch', type=int, default=4)
    ap.add_argument('--num_node_p_switch', type=int, default=32)
    opts = ap.parse_args(argv)
    failures = 0
    for run in (REFERENCE_SWEEP if opts.reference_sweep else DEVICE_SWEEP):
        for _ in range(opts.repeats):
            failures += launch(run, opts) != 0
    return failures


if __name__ == '__main__':
    sys.exit(1 if main() else 0)
