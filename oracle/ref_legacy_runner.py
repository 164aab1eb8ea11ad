"""Run the reference's DEAD-CODE legacy event loops UNMODIFIED, as the parity oracle for sjf / dlas-gpu.

    smallest_first_sim_jobs   /root/reference/run_sim.py:162-287   (--schedule sjf)
    shortest_first_sim_jobs   /root/reference/run_sim.py:299-431   (--schedule shortest / shortest-gpu)
    dlas_sim_jobs             /root/reference/run_sim.py:664-947   (--schedule dlas-gpu / dlas)

Those functions are unreachable from the reference's main(): they read four module globals that
run_sim.py never defines (JOBS, CLUSTER, LOG, scheduler) and two of the modules behind them import
`core.job`, a file the repository does not ship.  This runner supplies exactly those missing pieces
at run time (the reference's files are not touched) and then calls the functions as they are:

  real reference code that runs                                   what had to be supplied
  -----------------------------------------------------------     -------------------------------------
  run_sim.<loop>()              run_sim.py (imported as module)   the four globals, set as module attributes
  CLUSTER = infra.cluster.CLUSTER  (_Cluster: set_spec,           an empty stub module `core.job` so that
     empty_infra :88-95, release_job_res :1442-1483, free_gpu)      infra/switch.py:3 and log.py:7 import
  LOG = log._Log(dir)  (init_log :31-89, checkpoint :137-258,
     job_complete :316-330 -> the legacy cluster.csv / job.csv)
  core.scheduling.algorithm.ms_yarn_placement on a live            scheduler.try_get_job_res(CLUSTER, JOBS, job):
     infra.infrastructure.Infrastructure + core.jobs.job.Job        builds the live Job of the trace row and
     (sjf family: "placement = the live yarn fit")    calls ms_yarn_placement on an Infrastructure
                                                                    that CLUSTER.empty_infra() re-creates
  -                                                               JOBS: the Tiresias `_TFJobs` container the
                                                                    reference forked from but did not ship
                                                                    (class _ShimJobs below: job dicts, job_events
                                                                    grouped by submit time, runnable_jobs, queues,
                                                                    queue_limit, num_queue, move_to_runnable)

What the JOBS shim defines (the only reconstructed semantics; everything else is reference code):
  * job dict fields and their initial values (job_idx = position in arrival order, job_id = CSV row label,
    status 'ADDED', start_time = sys.maxsize as run_sim.py:251 / :833 test for it, counters 0);
  * the build-defined trace conversion of SURVEY 8(d): submit_time = ceil(normalized_time) ticks,
    duration = max(1, ceil(minutes * 0.5)) ticks, num_gpu = ceil(used_gpus);
  * job_events = [{'time': t, 'start_jobs': [...]}] sorted by time, jobs in arrival order;
  * move_to_runnable(job): status 'PENDING', last_check_time = submit_time, the time counters zeroed,
    appended to runnable_jobs.

Test infrastructure only (runs in THIS container; /root/reference does not exist on the GPU box).
Fixtures are written by oracle/make_golden.py from the files the reference's own _Log wrote.
"""
import json
import os
import subprocess
import sys
import tempfile
import time

REF = os.environ.get('RLGS_REFERENCE_DIR', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))

_DRIVER = r'''
import logging, os, sys, types, math
logging.disable(logging.CRITICAL)
REF = {ref!r}
sys.path.insert(0, REF)
sys.path.insert(1, {here!r})
sys.argv = ['run_sim.py'] + {argv!r}
_real_stdout = sys.stdout
sys.stdout = open(os.devnull, 'w')          # the loops print() per demotion / event

import core
_stub = types.ModuleType('core.job')        # imported by infra/switch.py:3 and log.py:7, absent from the repository
sys.modules['core.job'] = _stub
core.job = _stub

import numpy as np
import run_sim                               # flag definitions + the dead-code loops, unmodified
from infra import cluster as legacy_cluster  # _Cluster, CLUSTER
import log as legacy_log                     # _Log
FLAGS = run_sim.FLAGS
import cpu_sim                               # prepare_trace only: the same pandas calls as JobTraceReader.prepare_jobs

tr = cpu_sim.prepare_trace({trace!r})
n = len(tr['nt'])


class _ShimJobs(object):
    """Stand-in for the Tiresias job container the reference's loops expect in the global JOBS."""
    def __init__(self, num_queue, queue_limit):
        self.job_list, self.job_events, self.runnable_jobs = [], [], []
        self.pending_jobs, self.running_jobs, self.completed_jobs = [], [], []
        self.num_queue = num_queue
        self.queues = [list() for _ in range(num_queue)]
        self.queue_limit = list(queue_limit)
        for i in range(n):
            j = dict(job_idx=i, job_id=str(int(tr['label'][i])), num_gpu=int(math.ceil(tr['used_gpus'][i])),
                     submit_time=int(math.ceil(tr['nt'][i])), duration=max(1, int(math.ceil(tr['duration'][i]))),
                     status='ADDED', start_time=sys.maxsize, end_time=0, last_check_time=0, total_executed_time=0,
                     executed_time=0, pending_time=0, last_pending_time=0, q_id=0, preempt=0, resume=0, promote=0,
                     rank=sys.maxsize, placements=list(), _row=i)
            self.job_list.append(j)
            if self.job_events and self.job_events[-1]['time'] == j['submit_time']:
                self.job_events[-1]['start_jobs'].append(j)
            else:
                assert not self.job_events or self.job_events[-1]['time'] < j['submit_time']
                self.job_events.append(dict(time=j['submit_time'], start_jobs=[j]))

    def move_to_runnable(self, job):
        job['status'] = 'PENDING'
        job['start_time'] = sys.maxsize
        job['last_check_time'] = job['submit_time']
        job['total_executed_time'] = 0
        job['executed_time'] = 0
        job['pending_time'] = 0
        job['last_pending_time'] = 0
        self.runnable_jobs.append(job)


class _ShimScheduler(object):
    """scheduler.try_get_job_res for --scheme yarn: the live yarn fit on a live Infrastructure."""
    def __init__(self):
        from core.jobs import base_factory
        base_factory.BASE_OBJ = base_factory.BaseJobFactory(FLAGS)      # run_sim.py:1716 ("must do this first")
        from core.jobs.job import Job
        from core.scheduling.algorithm import ms_yarn_placement
        from infra.infrastructure import Infrastructure
        self.Job, self.place, self.Infrastructure = Job, ms_yarn_placement, Infrastructure
        self.reset()

    def reset(self):
        self.infra = self.Infrastructure(FLAGS)

    def try_get_job_res(self, cluster, jobs, rjob):
        i = rjob['_row']
        # the Job the live JobsManager.gen_jobs would build from this trace row (jobs_manager.py:233-238)
        j = self.Job(int(tr['label'][i]), float(tr['duration'][i]), float(tr['nt'][i]), int(tr['gpc'][i]),
                     gpu_utilization_avg=float(tr['util_avg'][i]), gpu_utilization_max=float(tr['util_max'][i]),
                     gpu_memory_max=float(tr['mem_mib'][i]), gpu_memory_avg=float(tr['mem_avg_mib'][i]),
                     total_gpus=float(tr['used_gpus'][i]))
        nodes, ok = self.place(self.infra, j, 'yarn')
        return bool(ok)


out_dir = {out_dir!r}
os.makedirs(out_dir, exist_ok=True)
CLUSTER = legacy_cluster.CLUSTER
CLUSTER.set_spec(FLAGS.num_switch, FLAGS.num_node_p_switch, FLAGS.num_gpu_p_node, FLAGS.num_cpu_p_node, FLAGS.mem_p_node)
LOG = legacy_log._Log(out_dir)
LOG.init_log()
JOBS = _ShimJobs({num_queue!r}, {queue_limit!r})
run_sim.JOBS, run_sim.CLUSTER, run_sim.LOG = JOBS, CLUSTER, LOG
if FLAGS.scheme != 'count':
    sched = _ShimScheduler()
    run_sim.scheduler = sched
    _real_empty = CLUSTER.empty_infra
    def _empty_infra():
        _real_empty()
        sched.reset()
    CLUSTER.empty_infra = _empty_infra

schedule = FLAGS.schedule
if schedule == 'sjf':
    run_sim.smallest_first_sim_jobs()
elif schedule == 'shortest':
    run_sim.shortest_first_sim_jobs(False)
elif schedule == 'shortest-gpu':
    run_sim.shortest_first_sim_jobs(True)
elif schedule == 'dlas-gpu':
    run_sim.dlas_sim_jobs(True)
elif schedule == 'dlas':
    run_sim.dlas_sim_jobs(False)
elif schedule == 'gittins' or schedule == 'dlas-gpu-gittins':
    JOBS.job_dist_data = run_sim.parse_job_dist() if hasattr(run_sim, 'parse_job_dist') else None
    run_sim.gittins_sim_jobs(JOBS.job_dist_data, True, True)
else:
    raise SystemExit('unknown legacy schedule ' + schedule)
'''


def available():
    return os.path.exists(os.path.join(REF, 'run_sim.py'))


def run_legacy(trace_csv, schedule, workdir=None, queue_limit=(30, 60, 150), **flags):
    """Returns dict(job_csv, cluster_csv, wall_s).  scheme = 'count' for the dlas family (admission by GPU count,
    run_sim.py:808-823), 'yarn' for the sjf family (scheduler.try_get_job_res)."""
    if not available():
        raise RuntimeError('reference not mounted at %s' % REF)
    workdir = workdir or tempfile.mkdtemp(prefix='rlgs_refleg_')
    dlas = schedule in ('dlas-gpu', 'dlas')
    scheme = 'count' if dlas else 'yarn'
    argv = ['--trace_file', os.path.abspath(trace_csv), '--schedule', schedule, '--scheme', scheme, '--log_path', 'oracle']
    for k, v in flags.items():
        argv += ['--' + k, str(v)]
    out_dir = os.path.join(workdir, 'legacy_out')
    code = _DRIVER.format(ref=REF, here=HERE, argv=argv, trace=os.path.abspath(trace_csv), out_dir=out_dir,
                          num_queue=(len(queue_limit) + 1) if dlas else 1, queue_limit=list(queue_limit) if dlas else [])
    t0 = time.time()
    p = subprocess.run([sys.executable, '-c', code], cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError('legacy reference run failed: rc=%s\n%s' % (p.returncode, p.stderr[-4000:]))
    res = dict(wall_s=wall, out_dir=out_dir, stderr=p.stderr[-2000:])
    for name in ('job', 'cluster'):
        fn = os.path.join(out_dir, name + '.csv')
        res[name + '_csv'] = open(fn, newline='').read() if os.path.exists(fn) else None
    return res


if __name__ == '__main__':
    r = run_legacy(sys.argv[1], sys.argv[2], **json.loads(sys.argv[3]) if len(sys.argv) > 3 else {})
    print(r['wall_s'], len(r['job_csv']), len(r['cluster_csv']))
    print(r['job_csv'][:600])
    print(r['cluster_csv'][:600])
