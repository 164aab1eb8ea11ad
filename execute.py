#!/usr/bin/env python
"""Experiment sweep driver with the reference's execute.py shape (execute.py:5-55): one run_sim.py
child process per (scheme, schedule, num_queue, num_buffer, repeat), strictly sequential.

The reference sweeps horus+/horus/gandiva/yarn; here the default sweep covers the schedules the device
path implements (fifo, sjf, dlas-gpu).  `--reference-sweep` uses the reference's own list: the
horus / gandiva combinations then exit with "not implemented by the device path".
"""
import argparse
import os
import sys
from subprocess import Popen


def do_once(scheme, schedule, num_queue, num_buffer, trace_file='month', num_nodes_p_switch=32, num_switch=4, data_dir='data'):
    migrate = True
    log_sub_dir = 'thesis_fitted_' + str(num_buffer) + '_nodes_p_s' + str(num_nodes_p_switch) + '_job_' + trace_file
    log_path = os.path.join(log_sub_dir, '%s_%s' % (scheme, schedule))
    if schedule == 'horus+':
        log_path = os.path.join(log_path, 'k' + str(num_queue))
    cmd = [sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'run_sim.py'),
           '--num_node_p_switch', str(num_nodes_p_switch), '--num_switch', str(num_switch), '--scheme', scheme,
           '--trace_file', os.path.join(data_dir, trace_file + '.csv'), '--num_queue', str(num_queue),
           '--num_buffer', str(num_buffer), '--schedule', schedule, '--enable_network_costs', 'False',
           '--enable_migration', str(migrate), '--log_path', log_path]
    p = Popen(cmd)
    print('process pid %d: ' % p.pid)
    try:
        return p.wait()
    except KeyboardInterrupt:
        p.kill()
        return -1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference-sweep', action='store_true')
    ap.add_argument('--trace', default='month')
    ap.add_argument('--data-dir', default='data')
    ap.add_argument('--repeats', type=int, default=3)
    a = ap.parse_args()
    if a.reference_sweep:
        schemes = ['horus+', 'horus+', 'horus+', 'horus', 'gandiva', 'yarn']
        queues = [3, 4, 5, 1, 1, 1]
        schedules = ['horus+', 'horus+', 'horus+', 'horus', 'gandiva', 'fifo']
        buffers = [15, 15, 15, 1, 1, 1]
    else:
        schemes, queues, schedules, buffers = ['yarn', 'yarn', 'count'], [1, 1, 4], ['fifo', 'sjf', 'dlas-gpu'], [1]
    for scheme, schedule, queue in zip(schemes, schedules, queues):
        for buff in buffers:
            for _ in range(a.repeats):
                do_once(scheme, schedule, queue, buff, trace_file=a.trace, data_dir=a.data_dir)


if __name__ == '__main__':
    main()
