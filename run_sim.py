#!/usr/bin/env python
"""GPU-cluster scheduling simulator — same command line and outputs as the reference's run_sim.py
(flags run_sim.py:19-94, bootstrap :1710-1757), with the tick / event loop executed on a B200 through
librlgs (include/rlgs.h).

    python run_sim.py --num_switch 4 --num_node_p_switch 32 --num_gpu_p_node 8 \
        --scheme yarn --schedule fifo --trace_file data/month.csv --log_path my_run

writes log/<log_path>/<timestamp>/{cluster.csv, job.csv, cpu.csv, gpu.csv, network.csv, memory.csv, output.log}.
"""
import datetime
import logging
import os
import time

from rlgpuschedule_b200 import flags
from rlgpuschedule_b200 import log_manager as lm
from rlgpuschedule_b200.host import Infrastructure, JobQueueManager, JobsManager, Scheduler

flags.DEFINE_string('trace_file', 'tf_job.csv', 'job trace file (*.csv) in the Philly-style schema')
flags.DEFINE_string('log_path', 'result-' + time.strftime('%Y%m%d-%H-%M-%S', time.localtime()),
                    'simulation output folder under log/; default result-[time]')
flags.DEFINE_string('scheme', 'yarn', 'job placement scheme: yarn | count | horus (= horus+ = gandiva, with --schedule horus | horus+ | gandiva)')
flags.DEFINE_string('schedule', 'fifo', 'job schedule: fifo | horus | horus+ | gandiva | sjf | shortest | shortest-gpu | dlas | dlas-gpu (all on the device)')
flags.DEFINE_boolean('pack', False, 'enable packing for gpu jobs (stored, not consulted by yarn)')
flags.DEFINE_integer('num_switch', 1, 'cluster spec: number of switches')
flags.DEFINE_integer('num_node_p_switch', 32, 'cluster spec: nodes under one switch')
flags.DEFINE_boolean('enable_network_costs', False, 'network costs when communicating over more nodes')
flags.DEFINE_boolean('enable_migration', False, 'preempt-and-migrate scan (inert in the reference)')
flags.DEFINE_integer('bandwidth', 1250, 'bandwidth per rack in MB/s')
flags.DEFINE_float('internode_latency', 0.015, 'inter-node latency in seconds')
flags.DEFINE_integer('gpu_memory_capacity', 32, 'GPU memory capacity in GiB')
flags.DEFINE_integer('num_queue', 1, 'number of queues within the job manager')
flags.DEFINE_integer('num_buffer', 5, 'number of jobs considered within the job manager (look-ahead k)')
flags.DEFINE_integer('num_gpu_p_node', 8, 'cluster spec: GPUs per node')
flags.DEFINE_integer('num_cpu_p_node', 128, 'cluster spec: CPUs per node')
flags.DEFINE_integer('mem_p_node', 512, 'cluster spec: memory per node')
flags.DEFINE_string('cluster_spec', None, 'cluster spec csv: num_switch,num_node_p_switch,num_gpu_p_node,num_cpu_p_node,mem_p_node')
flags.DEFINE_boolean('print', False, 'print debug information')
flags.DEFINE_boolean('flush_stdout', True, 'flush stdout')
# additions of this implementation
flags.DEFINE_string('queue_limit', '30,60,150', 'dlas-gpu: MLFQ demotion thresholds in GPU-ticks (README.md:57-62), comma separated')
flags.DEFINE_string('util_mode', 'sample', "avg_gpu_utilization column: 'sample' (seedable normal draw) or 'mean'")
flags.DEFINE_integer('seed', None, 'seed of the utilisation draws: the avg_gpu_utilization column and the horus score (the reference draws unseeded)')
flags.DEFINE_integer('device', 0, 'CUDA device ordinal')
flags.DEFINE_string('backend', 'cuda', "execution backend (SURVEY 8b): 'cuda' = librlgs.so on the GPU named by --device.  'python' is accepted for "
                    "interface compatibility and refused: this build has no CPU path; run the reference itself for that")
flags.DEFINE_boolean('columnar', False, 'also write cluster.parquet / job.parquet (typed columns, no float formatting)')
flags.DEFINE_version('0.1')

FLAGS = flags.FLAGS


def main(log_manager):
    if FLAGS.backend != 'cuda':
        raise SystemExit("--backend %s: only 'cuda' is built here (the product path has no CPU fallback); the Python backend is the "
                         "reference's own run_sim.py" % FLAGS.backend)
    infrastructure = Infrastructure(FLAGS)
    log_manager.init(infrastructure)
    jq_manager = JobQueueManager(FLAGS, os.path.abspath(FLAGS.trace_file))
    jobs_manager = JobsManager(FLAGS, jq_manager, infrastructure.cluster)
    scheduler = Scheduler(infrastructure, jobs_manager, log_manager, enable_migration=FLAGS.enable_migration)
    return scheduler.start()


if __name__ == '__main__':
    if FLAGS.backend != 'cuda':
        main(None)   # refuses before a log directory is created
    logging.basicConfig(format='%(asctime)s p%(process)s {%(module)s:%(lineno)d} %(levelname)s: %(message)s', level=logging.DEBUG)
    execution_id = datetime.datetime.now().strftime('%Y-%m-%d-%H-%M-%S-%f')
    output_dir = os.path.join('log', FLAGS.log_path, execution_id)
    os.makedirs(output_dir, exist_ok=True)
    filehandler = logging.FileHandler(filename=os.path.join(output_dir, 'output.log'), mode='w')
    filehandler.setFormatter(logging.Formatter('%(asctime)s %(levelname)s: %(message)s'))
    logging.getLogger().addHandler(filehandler)
    main(lm.LogManager(output_dir, FLAGS))
    logging.getLogger().removeHandler(filehandler)
