#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--replicas R]

Workload (BASELINE.json metric "simulated events/sec on 60k-job trace"): R independent replicas per
GPU of the 60 000-job Philly-style trace (rlgpuschedule_b200/synth.py frame_gen(60000, seed, 60000); seed 3 is the
trace whose reference output is pinned in tests/golden/probe60k) on the 4 switches x 32 nodes x 8 GPUs
simulated cluster under fifo + yarn.  One "step" = every replica simulated to completion.
An event = arrival | start | finish: 3 per finished job under non-preemptive fifo.

Prints ONE JSON line (rank 0).  `value` = events/s with traces resident in HBM (rows written to the
device-resident row store); `e2e` = the same through the C ABI with host buffers: trace upload,
simulation, rows + job tables copied back to the host, all inside the timed region.
`--impl reference` times the reference's algorithm on the host cores (oracle/cpu_sim.c, the C port
validated byte-for-byte against the real Python reference; the Python reference itself cannot
travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_JOBS = 60000
CLUSTER_FLAGS = dict(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)
N_TRACES = 8          # distinct traces per GPU, replicas are spread over them round-robin in blocks
METRIC = 'simulated_events_per_sec'
UNIT = 'events/s'
WORKLOAD = 'fifo+yarn, 4x32x8 simulated cluster, 60k-job Philly-style trace (gen(60000, seed, 60000)), %d replicas/GPU over %d seeds'


def frames(rank):
    from rlgpuschedule_b200 import synth
    return [synth.frame_gen(N_JOBS, 3 + rank * N_TRACES + i, N_JOBS) for i in range(N_TRACES)]


def algorithmic_bytes(summ, n_nodes, n_gpus):
    """bytes_tick = 8Q + 12R + 12N + 8D + 64, summed over the ticks of one replica-run."""
    return 8 * summ['sum_queued'] + 12 * summ['sum_running'] + summ['n_ticks'] * (12 * n_nodes + 8 * n_gpus + 64)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(',')])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace('.', '').isdigit())
        mx = [int(float(s[1])) for s in self.samples if len(s) > 1 and s[1].replace('.', '').isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(self.samples))


def run_reference(args, rank, world):
    """CPU arm: the oracle port on all host cores, a bounded sample of the same workload per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))   # the only legs that touch oracle/: the CPU baseline
    import cpu_sim
    cores = os.cpu_count() or 1
    cl = cpu_sim.make_cluster(**CLUSTER_FLAGS)
    traces = [cpu_sim.prepare_trace(f) for f in frames(0)[:min(N_TRACES, 4)]]
    cpu_sim.lib()
    per_step = 2 * cores  # replica-runs per step: two per hardware thread (~0.25 s each)
    k = [0]

    def step():
        k[0] += 1
        return cpu_sim.run_fifo_yarn_batch(cl, traces[k[0] % len(traces)], per_step, cores, rows_cap=70000)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    ev = 0
    for _ in range(args.steps):
        ev += step()
    dt = time.perf_counter() - t0
    val = ev / dt
    sample = '%d replica-runs of the 60k-job trace per step on %d pthreads (oracle/cpu_sim.c, gcc -O2)' % (per_step, cores)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'int32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD % (per_step, min(N_TRACES, 4)), 'note': 'CPU arm: bounded sample, host cores only'},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def cpu_baseline_sample():
    """Bounded cpu_baseline for the cuda arm's JSON line: a few seconds of the oracle on all cores."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import cpu_sim
    cores = os.cpu_count() or 1
    cl = cpu_sim.make_cluster(**CLUSTER_FLAGS)
    tr = cpu_sim.prepare_trace(frames(0)[0])
    cpu_sim.lib()
    cpu_sim.run_fifo_yarn_batch(cl, tr, cores, cores, rows_cap=70000)
    n_runs = 4 * cores
    t0 = time.perf_counter()
    ev = cpu_sim.run_fifo_yarn_batch(cl, tr, n_runs, cores, rows_cap=70000)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    ev1 = cpu_sim.run_fifo_yarn_batch(cl, tr, 1, 1, rows_cap=70000)
    dt1 = time.perf_counter() - t1
    return {'value': ev / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '%d replica-runs of the 60k-job trace on %d pthreads, %.1f s (oracle/cpu_sim.c, gcc -O2)' % (n_runs, cores, dt),
            'single_core_value': ev1 / dt1}


def bind_to_gpu_numa_node(index):
    """Pins this rank to the CPUs next to its GPU (NVML's ideal affinity) so that the pinned host mirrors it
    allocates are first-touched on the GPU's NUMA node: with several ranks per box the device->host copies
    otherwise cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--replicas', type=int, default=2960, help='replicas per GPU (default 20 resident warps x 148 SMs)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import rlgpuschedule_b200 as rl
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl cuda needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    R = args.replicas
    cluster = rl.Cluster(**CLUSTER_FLAGS)
    traces = [rl.prepare_trace(f, cluster) for f in frames(rank)]
    bounds = [R * i // N_TRACES for i in range(N_TRACES + 1)]

    def attach(sim):
        for i, tr in enumerate(traces):
            if bounds[i + 1] > bounds[i]:
                sim.load_trace(tr, bounds[i], bounds[i + 1] - bounds[i])

    # ---------------- device-resident arm: `value`
    sim = rl.Simulator(cluster, 'fifo', 'yarn', n_replicas=R, rows='device', device=local_rank)
    attach(sim)
    ret_dev = None
    if world > 1:
        class _Buf(object):
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {'shape': (n,), 'typestr': '<i8', 'data': (ptr, False), 'version': 3}
        ret_dev = torch.as_tensor(_Buf(sim.returns_device_ptr(), R), device='cuda')
        gathered = torch.empty(world * R, dtype=torch.int64, device='cuda')

    def step():
        sim.run()
        if world > 1:  # the one collective of the path: episode returns of every replica of every GPU
            dist.all_gather_into_tensor(gathered, ret_dev)

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    kernel_ms = 0.0
    launches = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
        ms, nl = sim.kernel_ms()
        kernel_ms += ms
        launches += nl
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.finish()
    summ = [sim.summary(bounds[i]) for i in range(N_TRACES) if bounds[i + 1] > bounds[i]]
    events_step = sum(s['events'] * (bounds[i + 1] - bounds[i]) for i, s in enumerate(summ))
    jobs_step = sum(s['n_finished'] * (bounds[i + 1] - bounds[i]) for i, s in enumerate(summ))
    ticks_step = sum(s['n_ticks'] * (bounds[i + 1] - bounds[i]) for i, s in enumerate(summ))
    alg_bytes_step = sum(algorithmic_bytes(s, cluster.num_nodes, cluster.num_gpus) * (bounds[i + 1] - bounds[i]) for i, s in enumerate(summ))
    hbm_stream_bytes_step = sum((s['n_jobs'] * (32 + 32 + 32 + 12 + 8) + s['n_ticks'] * 64) * (bounds[i + 1] - bounds[i]) for i, s in enumerate(summ))
    t = torch.tensor([dt, kernel_ms / 1e3], dtype=torch.float64, device='cuda')
    tot = torch.tensor([events_step, jobs_step, ticks_step], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        if rank == 0:
            g = gathered.cpu().numpy().reshape(world, R)
            assert (g[0] == sim.returns()).all(), 'all-gathered returns do not match the local ones'
    dt_max, kern_s = t.tolist()
    events_all, jobs_all, ticks_all = tot.tolist()
    value = events_all * args.steps / dt_max
    sim.close()

    # ---------------- end-to-end arm through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        sim2 = rl.Simulator(cluster, 'fifo', 'yarn', n_replicas=R, rows='host', fetch_jobs=True, device=local_rank)
        attach(sim2)

        def step2():
            attach(sim2)          # host -> device: the step's input records
            sim2.run()            # simulate; rows + job tables -> pinned host store, overlapped per stream
            return int(sim2.summary(0)['n_finished'])
        for _ in range(max(args.warmup, 3)):
            step2()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step2()
        barrier()
        dt2 = time.perf_counter() - t0
        t2 = torch.tensor([dt2], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        rows0 = sim2.rows_view(0)
        j0 = sim2.jobs(0)
        assert len(rows0) == summ[0]['n_ticks'] and int(rows0['finished'][-1]) == len(j0['finish_order'])
        h2d = sum(len(tr.records) * 32 for tr in traces)
        from rlgpuschedule_b200 import _ffi
        n_chunks = -(-max(s_['n_ticks'] for s_ in summ) // _ffi.ROWS_PER_CHUNK)        # whole chunks travel (chunk-major row store)
        jmax = max(len(tr.records) for tr in traces)
        d2h = int(n_chunks * R * _ffi.ROWS_PER_CHUNK * 64 + 3 * 4 * R * jmax + R * 264)
        e2e = {'value': events_all * args.steps / t2.item(), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
               'ms_per_step': 1e3 * t2.item() / args.steps}
        sim2.close()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        per_launch_s = kern_s / args.steps
        achieved = alg_bytes_step / per_launch_s / 1e9
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of one step's kernel work, from the committed ncu capture
        try:
            tj = json.load(open(os.path.join(ROOT, 'profiles', 'r01_bench_traffic.json')))
            traffic = tj['traffic_bytes_per_launch'] * R / tj.get('replicas', 2368)   # rows + queue stack + job tables scale with the replica count
        except Exception:
            pass
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': 1e3 * dt_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'int32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD % (R, N_TRACES), 'schedule': 'fifo', 'scheme': 'yarn', 'replicas_per_gpu': R,
                       'jobs_per_replica': N_JOBS, 'l2': 'per-step working set (queue stacks + row store + job tables, %.1f GB/GPU) >> 126 MB L2; no explicit flush'
                       % ((sum(len(tr.records) for tr in traces) / N_TRACES * 32 * R + ticks_step * 64) / 1e9),
                       'parallelism': 'replicas: %d GPU x %d warps (1 warp = 1 replica)' % (world, R)},
            'jobs_per_sec': jobs_all * args.steps / dt_max, 'ticks_per_sec': ticks_all * args.steps / dt_max,
            'gpu_launches': launches, 'kernel_ms_per_step': 1e3 * per_launch_s,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                         'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s',
                         'launch': 'one step = %d concurrent launches of fifo_yarn_kernel<false> (one per replica group / CUDA stream); bytes and time are per step' % (launches // max(args.steps, 1)),
                         'note': 'achieved = SURVEY 8(d) algorithmic bytes (8Q+12R+12N+8D+64 per tick) / kernel time; that state is held in '
                                 'shared memory / registers, so it is not DRAM traffic. hbm_stream_GBps = bytes this layout must move through HBM '
                                 '(records in, queue stack write+read, job tables, 64 B row per tick) / kernel time',
                         'hbm_stream_GBps': hbm_stream_bytes_step / per_launch_s / 1e9},
            'clocks': clocks,
        }
        if e2e:
            out['e2e'] = e2e
        if not args.no_cpu and world == 1:   # the CPU baseline is reported at N=1 only, on every host core
            os.sched_setaffinity(0, all_cpus)
            out['cpu_baseline'] = cpu_baseline_sample()
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
