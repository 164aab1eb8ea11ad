/*
 * oracle/cpu_sim.c — CPU restatement of the reference simulator's hot path.  TEST INFRASTRUCTURE.
 *
 * This file is the parity checker (tests/, __graft_entry__.smoke()) and the "port" CPU baseline
 * (bench.py cpu_baseline / --impl reference).  It is never linked into, imported by or called
 * from the product library (rlgpuschedule_b200/csrc, include/rlgs.h).
 *
 * It restates, function by function, the simulator of matthewygf/RLGPUSchedule, keeping the reference's
 * per-tick whole-state sweeps and all quirks of SURVEY.md Appendix A.  Citations are path:line under
 * /root/reference.  Sections and their parity status:
 *
 *   1. live tick loop, `--schedule fifo --scheme yarn` (oracle_env_yarn / oracle_fifo_yarn): PINNED — byte for byte
 *      against job.csv / cluster.csv written by the unmodified reference (tests/golden, oracle/make_golden.py,
 *      tests/test_oracle_golden.py) and on random traces run live (tests/test_oracle_vs_live_reference.py).
 *      Its network-cost option and the window policies of the RL environment are build-defined: UNPINNED.
 *   2. legacy event-driven schedules sjf / shortest / shortest-gpu / dlas / dlas-gpu (oracle_sjf_yarn, oracle_dlas):
 *      PINNED GIVEN SHIMS — the loops are dead code in the reference (undefined globals JOBS / CLUSTER / LOG / scheduler),
 *      but oracle/ref_legacy_runner.py executes them UNMODIFIED (run_sim.py:162-287, :299-431, :664-947, with the
 *      reference's own infra.cluster._Cluster and log._Log) after supplying the missing pieces at run time: the job
 *      container JOBS (reconstructed; its semantics are listed in that file's header), scheduler.try_get_job_res = the live
 *      ms_yarn_placement, and an empty stub for the missing module core.job.  This section matches the cluster.csv /
 *      job.csv those runs wrote byte for byte on 26 fixtures (tests/golden/{sjf,shortest,shortestgpu,dlasgpu,dlas}_*,
 *      up to the 10 000-job sjf trace = BASELINE config C2 and the 60 000-job dlas-gpu trace = C3).  What stays
 *      build-defined: the trace -> (submit, duration, num_gpu) conversion in ticks and queue_limit = [30, 60, 150].
 *   3. pack family, `--schedule horus | horus+ | gandiva` over horus_placement or yarn (oracle_pack): PINNED on traces
 *      without utilisation spread (and on any trace over yarn); horus+ with its k-means draws injected into the
 *      reference run.  The counter-based utilisation draw used when there is a spread is build-defined: UNPINNED.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CPUS_PER_TASK 12 /* core/jobs/job.py:105 */
#define MEM_PER_TASK 60  /* core/jobs/job.py:106 */
#define MAX_TASKS_PER_DEVICE 4 /* infra/device.py:72 */

typedef struct {
    int32_t num_switch;
    int32_t num_node_p_switch;
    int32_t num_gpu_p_node;
    int32_t num_cpu_p_node;
    int32_t mem_p_node;
    int32_t gpu_memory_capacity_mib; /* flags.gpu_memory_capacity * 1024, infra/infrastructure.py:36 */
} oracle_cluster;

/* One row of cluster.csv as the reference computes it (core/scheduling/schedule.py:95-133). */
typedef struct {
    int32_t idle_nodes, busy_nodes, busy_gpus, idle_gpus;
    int32_t running, queued, finished;
    int32_t max_is_int_zero;  /* max_pending_time printed as the int 0 (empty queue) */
    double avg_gpu_memory_allocated;
    double avg_pending, median_pending, max_pending;
    double util_mu_sum;   /* sum over busy devices of gpu_utilization_avg: mean of the RNG column */
} oracle_row;

typedef struct {
    int32_t cpu_used, mem_used;
    int32_t placed_jobs;   /* len(node.placed_jobs): never popped on completion (q3) */
    int32_t placed_tasks;  /* len(node.placed_tasks) */
    int32_t running_tasks; /* len(node.running_tasks) */
} onode;

typedef struct {
    int32_t ntasks;        /* len(device.running_tasks) */
    int32_t owner_job[MAX_TASKS_PER_DEVICE];
    int32_t owner_task[MAX_TASKS_PER_DEVICE];
} odevice;

typedef struct {
    /* immutable */
    double nt, duration, gpus, mem_mib, util_avg;
    int32_t gpc, task_count, submit_time;
    /* state */
    double pending_time;
    int32_t time_processed; /* all tasks of a job step together (core/jobs/job.py:183-188) */
    int32_t running, finished, start_time, end_time, migration_count, tasks_finished;
    int32_t *task_node;     /* tasks_running_on: node index per task, -1 = not placed */
} ojob;

typedef struct {
    const oracle_cluster *c;
    int32_t N, G;
    onode *nodes;
    odevice *devs; /* N*G */
    ojob *jobs;
    int32_t n;
    int32_t *queue; int32_t qlen;      /* job_queue_manager.queues[0]: python list */
    int32_t *running; int32_t rlen;    /* jobs_manager.running_jobs: dict in insertion order */
    int32_t *fin; int32_t flen;        /* jobs_manager.finished_jobs */
    int32_t cursor;                    /* rows of the sorted trace already generated */
    int32_t *task_pool;
    /* network costs (core/network/network_service.py:3-39); enabled when net_model != NULL */
    const double *net_model, *net_iters; double net_bandwidth, net_latency;
} osim;

/* python float // float (Objects/floatobject.c float_floor_div) for the worker_count of job.py:96 */
static double py_floordiv(double vx, double wx) {
    double mod = fmod(vx, wx), div = (vx - mod) / wx, fl;
    if (mod != 0.0 && ((wx < 0) != (mod < 0))) div -= 1.0;
    if (div != 0.0) { fl = floor(div); if (div - fl > 0.5) fl += 1.0; }
    else fl = copysign(0.0, vx / wx);
    return fl;
}

static int node_cpu_free(const osim *s, int i) { return s->c->num_cpu_p_node - s->nodes[i].cpu_used; } /* infra/node.py:51 */
static int node_mem_free(const osim *s, int i) { return s->c->mem_p_node - s->nodes[i].mem_used; }     /* infra/node.py:55 */
static int node_is_free(const osim *s, int i) { return node_cpu_free(s, i) > 0 || node_mem_free(s, i) > 0; } /* node.py:59 */
static int node_is_idle(const osim *s, int i) { /* infra/node.py:93-97 */
    const onode *nd = &s->nodes[i];
    return nd->running_tasks + nd->placed_tasks + nd->placed_jobs == 0;
}
static int node_free_devices(const osim *s, int i) { /* infra/node.py:99-107, pack=False */
    int cnt = 0;
    for (int g = 0; g < s->G; ++g) cnt += s->devs[i * s->G + g].ntasks == 0;
    return cnt;
}
static double device_current_memory(const osim *s, const odevice *d) { /* infra/device.py:56-62 */
    double cap = s->c->gpu_memory_capacity_mib, mem = 0;
    for (int k = 0; k < d->ntasks; ++k) {
        double m = s->jobs[d->owner_job[k]].mem_mib;
        mem += (m < cap) ? m : cap;
        mem = (mem < cap) ? mem : cap;
    }
    return mem;
}
static int device_can_fit(const osim *s, const odevice *d, const ojob *j) { /* infra/device.py:67-77 */
    double cur = device_current_memory(s, d);
    if (d->ntasks >= MAX_TASKS_PER_DEVICE) return 0;
    return ((double)s->c->gpu_memory_capacity_mib - (cur + j->mem_mib)) > 500;
}
static int device_add_task(osim *s, odevice *d, int job, int task) { /* infra/device.py:19-43, pack=False */
    if (!device_can_fit(s, d, &s->jobs[job])) return 0;
    if (d->ntasks > 0) return 0; /* `not pack and len(running_tasks) > 0` */
    d->owner_job[d->ntasks] = job; d->owner_task[d->ntasks] = task; d->ntasks++;
    return 1;
}

/* infra/node.py:109-127 (pack=False). `ntasks` = len(tasks) passed in. */
static int node_can_fit_num_task(const osim *s, int i, const ojob *j, int ntasks) {
    /* python // on ints floors toward -inf (cpu_free can be negative after the q8 leak) */
    int fd = node_free_devices(s, i);
    int a = fd / j->gpc;
    int cf = node_cpu_free(s, i), mf = node_mem_free(s, i);
    int b = (int)floor((double)cf / CPUS_PER_TASK), m = (int)floor((double)mf / MEM_PER_TASK);
    int g_fit = (a - ntasks >= 0) ? ntasks : a;
    int c_fit = (b - ntasks >= 0) ? ntasks : b;
    int m_fit = (m - ntasks >= 0) ? ntasks : m;
    int r = c_fit < m_fit ? c_fit : m_fit;
    return r < g_fit ? r : g_fit;
}
static int node_can_fit(const osim *s, int i, const ojob *j) { /* infra/node.py:146-171, pack=False */
    int cpu_off = node_cpu_free(s, i) - CPUS_PER_TASK, mem_off = node_mem_free(s, i) - MEM_PER_TASK;
    if (cpu_off < 0 || mem_off < 0) return 0;
    return node_free_devices(s, i) - j->gpc >= 0;
}
/* infra/node.py:200-221: cpu/mem are charged BEFORE the device loop and not undone (q8). */
static int node_try_reserve_and_placed_task(osim *s, int i, int job, int task) {
    ojob *j = &s->jobs[job];
    if (!node_can_fit(s, i, j)) return 0;
    s->nodes[i].cpu_used += CPUS_PER_TASK;
    s->nodes[i].mem_used += MEM_PER_TASK;
    int should = j->gpc;
    for (int g = 0; g < s->G; ++g) {
        if (should <= 0) break;
        if (device_add_task(s, &s->devs[i * s->G + g], job, task)) should--;
    }
    if (should == 0) s->nodes[i].placed_tasks++;
    return should == 0;
}
/* infra/node.py:71-91 */
static void node_release_allocated_resources(osim *s, int i, int job, int task) {
    s->nodes[i].cpu_used -= CPUS_PER_TASK;
    s->nodes[i].mem_used -= MEM_PER_TASK;
    for (int g = 0; g < s->G; ++g) {
        odevice *d = &s->devs[i * s->G + g];
        for (int k = 0; k < d->ntasks; ++k)
            if (d->owner_job[k] == job && d->owner_task[k] == task) {
                for (int q = k; q + 1 < d->ntasks; ++q) { d->owner_job[q] = d->owner_job[q + 1]; d->owner_task[q] = d->owner_task[q + 1]; }
                d->ntasks--; break;
            }
    }
}

/* core/scheduling/algorithm.py:396-417 + infra/node.py:245-275 (try_alloc_job, is_single=True) */
static int try_single_node_alloc_ms(osim *s, int job) {
    ojob *j = &s->jobs[job];
    int N = s->N;
    uint8_t *isfree = (uint8_t *)malloc(N);
    for (int i = 0; i < N; ++i) isfree[i] = node_is_free(s, i); /* get_free_nodes(): list built up front */
    int allocated = 0;
    for (int i = 0; i < N && !allocated; ++i) {
        if (!isfree[i]) continue;
        if (!((double)node_free_devices(s, i) >= j->gpus &&
              node_cpu_free(s, i) >= CPUS_PER_TASK * j->task_count &&
              node_mem_free(s, i) >= MEM_PER_TASK * j->task_count)) continue;
        /* try_alloc_job */
        int worker_tasks = node_can_fit_num_task(s, i, j, j->task_count);
        if (worker_tasks < j->task_count) continue;
        int placed = 0;
        for (int t = 0; t < j->task_count; ++t)
            if (node_try_reserve_and_placed_task(s, i, job, t)) { j->task_node[t] = i; placed++; }
        if (placed == 0) continue; /* result False: nothing recorded, cpu/mem leak stays (q8) */
        /* try_reserve_and_placed_job(job, True): first task must be in placed_tasks */
        if (j->task_node[0] != i) { free(isfree); return -1; } /* reference would hit node.py:264 (q10) */
        s->nodes[i].placed_jobs++;
        allocated = 1;
    }
    free(isfree);
    return allocated;
}

/* core/scheduling/algorithm.py:301-393 */
static int try_cross_node_alloc_ms(osim *s, int job) {
    ojob *j = &s->jobs[job];
    int least = (int)ceil(j->gpus / (double)s->c->num_gpu_p_node);
    int num_full = j->task_count, assigned = 0, nodes_assigned = 0;
    uint8_t *node_used = (uint8_t *)calloc(s->N, 1);
    for (int i = 0; i < s->N; ++i) {
        if (!node_is_free(s, i)) continue;
        if (assigned == num_full) break;
        int fit = node_can_fit_num_task(s, i, j, num_full - assigned);
        if (fit == 0) continue;
        int worker_count = 0, check_next = 0;
        for (int t = 0; t < j->task_count; ++t) {
            if (j->task_node[t] >= 0) continue;        /* k in assigned_task */
            if (!(worker_count <= fit)) continue;      /* q11: tries fit+1 tasks */
            worker_count++;
            if (!node_try_reserve_and_placed_task(s, i, job, t)) { worker_count--; check_next = 1; break; }
            j->task_node[t] = i; assigned++;
        }
        if (worker_count > 0) { if (!node_used[i]) { node_used[i] = 1; nodes_assigned++; s->nodes[i].placed_jobs++; } }
        if (check_next) continue;
        if (nodes_assigned >= least && num_full == assigned) break;
    }
    if (assigned < num_full || nodes_assigned < least) {
        for (int i = 0; i < s->N; ++i) {
            if (!node_used[i]) continue;
            s->nodes[i].placed_jobs--;
            for (int t = 0; t < j->task_count; ++t)
                if (j->task_node[t] == i) { s->nodes[i].placed_tasks--; node_release_allocated_resources(s, i, job, t); }
        }
        for (int t = 0; t < j->task_count; ++t) j->task_node[t] = -1; /* stale entries are overwritten on retry */
        free(node_used);
        return 0;
    }
    free(node_used);
    return 1;
}

/* Counter-based RNG of the build-defined random-window policy: splitmix64 finaliser of (seed, replica, tick).
 * The CUDA kernel (rlgpuschedule_b200/csrc/env_step.cuh) uses the same function. */
static uint32_t rlgs_hash3(uint32_t seed, uint32_t replica, uint32_t tick) {
    uint64_t z = ((uint64_t)seed << 32) ^ ((uint64_t)replica * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)tick * 0xBF58476D1CE4E5B9ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

/* core/scheduling/schedule.py:40-60 + algorithm.py:189-202 (schedule_fifo) + :28-32 (ms_yarn_placement).
 * `pick` = position in the queue of the job to try (0 = the fifo head; the RL environment passes the
 * agent's choice inside the look-ahead window, cf. schedule_horus' window algorithm.py:204-240). */
static int schedule_once(osim *s, int delta, int pick) {
    int nfree = 0;
    for (int i = 0; i < s->N; ++i) nfree += node_is_free(s, i); /* schedule.py:36-38 */
    if (nfree < 1) return 0;
    int job = s->queue[pick];                                  /* jobs_manager.py:32-37 */
    ojob *j = &s->jobs[job];
    if (!(j->submit_time <= delta)) return 0;
    int ok = (j->gpus > (double)s->c->num_gpu_p_node) ? try_cross_node_alloc_ms(s, job)
                                                        : try_single_node_alloc_ms(s, job);
    if (ok < 0) return -1;
    if (!ok) return 0;
    memmove(s->queue + pick, s->queue + pick + 1, (size_t)(s->qlen - 1 - pick) * sizeof(int32_t)); s->qlen--; /* queues[0].pop(pick) */
    /* add_to_running -> start_job -> Node.execute_job -> Task.execute / Job.try_execute
       (schedule.py:164-167, jobs_manager.py:189-207, node.py:173-198, job.py:44-48,160-175) */
    if (s->net_model) {
        /* schedule.py:49-54 -> calculate_network_costs + Job.add_network_costs (job.py:196-197).  The reference's
           own code raises here (Job.is_distributed reads self.ps_count, job.py:199-200), so this is build-defined:
           distributed = more than one task; no PS tasks exist, so ps_nodes ^ wk_nodes = the worker nodes. */
        if (j->task_count > 1) {
            int cross = 0;
            for (int t = 0; t < j->task_count; ++t) { int seen = 0; for (int u = 0; u < t; ++u) seen |= j->task_node[u] == j->task_node[t]; cross += !seen; }
            double model_per_sec = s->net_model[job] / s->net_bandwidth;
            double nodes_induced_sec = (double)cross * s->net_latency;
            double iteration_round_trip = s->net_iters[job] * 2.0;
            j->duration += (model_per_sec + nodes_induced_sec) * iteration_round_trip;
        }
    }
    for (int t = 0; t < j->task_count; ++t) { int nd = j->task_node[t]; s->nodes[nd].placed_tasks--; s->nodes[nd].running_tasks++; }
    j->start_time = delta; j->migration_count += 1; j->running = 1;
    s->running[s->rlen++] = job;
    return 1;
}

static int cmp_double(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }

/* schedule.py:95-133 + jobs_manager.py:72-87 */
static void construct_info(osim *s, oracle_row *r, double *scratch) {
    int idle_nodes = 0, busy_nodes = 0, idle_gpus = 0, busy_gpus = 0;
    double avg_mem = 0, mu = 0; long long sum_cap = 0;
    for (int i = 0; i < s->N; ++i) {
        if (node_is_idle(s, i)) idle_nodes++; else busy_nodes++;
        for (int g = 0; g < s->G; ++g) {
            odevice *d = &s->devs[i * s->G + g];
            if (d->ntasks == 0) idle_gpus++;
            else {
                busy_gpus++; avg_mem += device_current_memory(s, d);
                for (int k = 0; k < d->ntasks; ++k) mu += s->jobs[d->owner_job[k]].util_avg;
            }
            sum_cap += s->c->gpu_memory_capacity_mib;
        }
    }
    avg_mem /= (double)sum_cap;
    double pend = 0, maxp = 0;
    for (int q = 0; q < s->qlen; ++q) {
        double p = s->jobs[s->queue[q]].pending_time;
        if (p > maxp) maxp = p;   /* max(j_pend, max_pending), max_pending starts as the int 0 */
        scratch[q] = p; pend += p;
    }
    /* the printed value stays the int 0 only for an empty queue: every queued job has
       pending_time >= 1.0 when the row is built (step() ran after its arrival) */
    r->max_is_int_zero = !(maxp > 0);
    r->idle_nodes = idle_nodes; r->busy_nodes = busy_nodes; r->busy_gpus = busy_gpus; r->idle_gpus = idle_gpus;
    r->avg_gpu_memory_allocated = avg_mem; r->util_mu_sum = mu;
    r->avg_pending = pend / ((double)s->qlen + 1e-9);
    if (s->qlen == 0) r->median_pending = NAN;
    else {
        qsort(scratch, (size_t)s->qlen, sizeof(double), cmp_double);
        int n = s->qlen;
        r->median_pending = (n & 1) ? scratch[n / 2] : (scratch[n / 2 - 1] + scratch[n / 2]) / 2.0; /* np.median */
    }
    r->max_pending = maxp;
    r->running = s->rlen; r->queued = s->qlen; r->finished = s->flen;
}

/*
 * Runs the reference's Scheduler.start() loop (core/scheduling/schedule.py:178-216).
 * Inputs are per job in the order of JobTraceReader.prepare_jobs()' sorted DataFrame
 * (core/jobs/job_generator.py:181-196): nt = normalized_time/10000 after the shift.
 * Outputs: finish_order[k] = index (into the inputs) of the k-th finished job (job.csv row order),
 * start/end per job, rows[0..n_ticks).  counters[0..3] = sum over ticks of queued / running jobs,
 * ticks, starts (for the events/bytes accounting of SURVEY.md §8d).
 * Returns 0, -1 = rows_cap too small, -2 = the reference would raise (q10 / zero-task job).
 */
/* thread-local network-cost inputs picked up by the next oracle_env_yarn call of this thread */
static __thread const double *tl_net_model = NULL, *tl_net_iters = NULL; static __thread double tl_net_bw = 0, tl_net_lat = 0;
static __thread double *tl_dur_out = NULL;
void oracle_set_netcost(const double *model_mb, const double *iterations, double bandwidth, double latency, double *dur_out) {
    tl_net_model = model_mb; tl_net_iters = iterations; tl_net_bw = bandwidth; tl_net_lat = latency; tl_dur_out = dur_out;
}

/* policy: 0 = fifo head (the reference), 1 = uniformly random job of the first min(k, queued) jobs drawn
 * from rlgs_hash3(seed, replica, tick), 2 = actions[tick] (-1 = no scheduling attempt this tick; values
 * >= min(k, queued) are no-ops too).  Policies 1 and 2 are the build-defined RL environment (the
 * reference's model/env.py is an empty stub): PARITY UNPINNED. */
int oracle_env_yarn(const oracle_cluster *c, int32_t n, const double *nt, const double *duration,
                    const double *used_gpus, const int32_t *gpc, const double *mem_max_mib,
                    const double *util_avg, int32_t policy, int32_t window_k, uint32_t seed, uint32_t replica,
                    const int32_t *actions, int64_t n_actions,
                    int32_t *finish_order, int32_t *start_tick, int32_t *end_tick, int32_t *n_finished,
                    oracle_row *rows, int64_t rows_cap, int64_t *n_ticks, int64_t *counters) {
    osim S; memset(&S, 0, sizeof S);
    S.c = c; S.N = c->num_switch * c->num_node_p_switch; S.G = c->num_gpu_p_node; S.n = n;
    S.nodes = (onode *)calloc((size_t)S.N, sizeof(onode));
    S.devs = (odevice *)calloc((size_t)S.N * S.G, sizeof(odevice));
    S.jobs = (ojob *)calloc((size_t)n + 1, sizeof(ojob));
    S.queue = (int32_t *)malloc(((size_t)n + 1) * 4); S.running = (int32_t *)malloc(((size_t)n + 1) * 4);
    S.fin = finish_order;
    double *scratch = (double *)malloc(((size_t)n + 1) * 8);
    int64_t total_tasks = 0; int rc = 0;
    for (int i = 0; i < n; ++i) {
        ojob *j = &S.jobs[i];
        j->nt = nt[i]; j->duration = duration[i]; j->gpus = used_gpus[i]; j->gpc = gpc[i];
        j->mem_mib = mem_max_mib[i]; j->util_avg = util_avg ? util_avg[i] : 0.0;
        j->submit_time = (int32_t)nt[i];                            /* job.py:93 int() */
        if (gpc[i] <= 0) { rc = -2; goto done; }                    /* ZeroDivisionError at job.py:96 */
        j->task_count = (int32_t)py_floordiv(used_gpus[i], (double)gpc[i]); /* job.py:96-99 */
        if (j->task_count <= 0) { rc = -2; goto done; }             /* StopIteration at node.py:118 */
        total_tasks += j->task_count;
        start_tick[i] = -1; end_tick[i] = -1;
    }
    S.net_model = tl_net_model; S.net_iters = tl_net_iters; S.net_bandwidth = tl_net_bw; S.net_latency = tl_net_lat;
    S.task_pool = (int32_t *)malloc(((size_t)total_tasks + 1) * 4);
    { int64_t off = 0; for (int i = 0; i < n; ++i) { S.jobs[i].task_node = S.task_pool + off; for (int t = 0; t < S.jobs[i].task_count; ++t) S.task_pool[off + t] = -1; off += S.jobs[i].task_count; } }

    int64_t delta = 0, sumQ = 0, sumR = 0, starts = 0;
    int remaining = n, running_jobs = 0;
    while (remaining + running_jobs > 0 && !(policy == 2 && delta >= n_actions)) {   /* schedule.py:185 (queue ignored, q2); an action tape ends the episode prefix */
        /* gen_jobs (jobs_manager.py:228-241, job_generator.py:198-207): rows with nt <= delta */
        int k0 = S.cursor;
        while (S.cursor < n && S.jobs[S.cursor].nt <= (double)delta) S.cursor++;
        int k = S.cursor - k0;
        if (k > 0) { /* insert(): queue.insert(i, job_i) for i = 0..k-1 => batch at the FRONT (q1) */
            memmove(S.queue + k, S.queue, (size_t)S.qlen * 4);
            for (int i = 0; i < k; ++i) S.queue[i] = k0 + i;
            S.qlen += k;
        }
        if (S.qlen > 0) {
            int pick = 0, win = S.qlen < window_k ? S.qlen : window_k;
            if (policy == 1) pick = (int)(rlgs_hash3(seed, replica, (uint32_t)delta) % (uint32_t)win);
            else if (policy == 2) pick = (delta < n_actions) ? actions[delta] : -1;
            if (pick >= 0 && pick < win) { int r = schedule_once(&S, (int)delta, pick); if (r < 0) { rc = -2; goto done; } starts += r; }
        }
        remaining = n - S.cursor;                                    /* schedule.py:191 */
        delta += 1;
        /* jobs_manager.step(): add_pending_time + Job.step for running jobs (jobs_manager.py:143-148) */
        for (int q = 0; q < S.qlen; ++q) S.jobs[S.queue[q]].pending_time += 1;
        for (int r = 0; r < S.rlen; ++r) S.jobs[S.running[r]].time_processed += 1;
        /* release_finished_jobs (schedule.py:141-162) over prepare_finish_tasks (jobs_manager.py:243-250) */
        int w = 0;
        for (int r = 0; r < S.rlen; ++r) {
            int job = S.running[r]; ojob *j = &S.jobs[job];
            if ((double)j->time_processed < j->duration) { S.running[w++] = job; continue; }
            for (int t = 0; t < j->task_count; ++t) {
                int nd = j->task_node[t];
                S.nodes[nd].running_tasks--;
                node_release_allocated_resources(&S, nd, job, t);
                j->tasks_finished++;
            }
            j->running = 0; j->finished = 1; j->end_time = (int32_t)delta;
            S.fin[S.flen++] = job;
        }
        S.rlen = w; running_jobs = S.rlen;
        if (delta > rows_cap) { rc = -1; goto done; }
        construct_info(&S, &rows[delta - 1], scratch);
        sumQ += S.qlen; sumR += S.rlen;
    }
    if (tl_dur_out) for (int i = 0; i < n; ++i) tl_dur_out[i] = S.jobs[i].duration;
    for (int i = 0; i < n; ++i) if (S.jobs[i].finished) { start_tick[i] = S.jobs[i].start_time; end_tick[i] = S.jobs[i].end_time; }
                               else if (S.jobs[i].running) start_tick[i] = S.jobs[i].start_time;
    *n_finished = S.flen; *n_ticks = delta;
    if (counters) { counters[0] = sumQ; counters[1] = sumR; counters[2] = delta; counters[3] = starts; }
done:
    free(S.nodes); free(S.devs); free(S.jobs); free(S.queue); free(S.running); free(scratch); free(S.task_pool);
    return rc;
}

int oracle_fifo_yarn(const oracle_cluster *c, int32_t n, const double *nt, const double *duration,
                     const double *used_gpus, const int32_t *gpc, const double *mem_max_mib,
                     const double *util_avg,
                     int32_t *finish_order, int32_t *start_tick, int32_t *end_tick, int32_t *n_finished,
                     oracle_row *rows, int64_t rows_cap, int64_t *n_ticks, int64_t *counters) {
    return oracle_env_yarn(c, n, nt, duration, used_gpus, gpc, mem_max_mib, util_avg, 0, 1, 0, 0, NULL, 0,
                           finish_order, start_tick, end_tick, n_finished, rows, rows_cap, n_ticks, counters);
}

/* ---------------------------------------------------------------------------------------------
 * Multi-threaded batch driver for the CPU baseline (bench.py cpu_baseline / --impl reference):
 * `n_runs` independent replica-runs of the same trace spread over `n_threads` pthreads, each with
 * its own output buffers.  Returns the total number of events (3 per finished job) or < 0.
 * ------------------------------------------------------------------------------------------- */
#include <pthread.h>

typedef struct {
    const oracle_cluster *c; int32_t n;
    const double *nt, *duration, *used_gpus, *mem; const int32_t *gpc; const double *util;
    int64_t rows_cap; int n_runs, n_threads, tid; int64_t events; int rc;
} batch_arg;

static void *batch_worker(void *p) {
    batch_arg *a = (batch_arg *)p;
    int32_t *fin = (int32_t *)malloc(((size_t)a->n + 1) * 4), *st = (int32_t *)malloc(((size_t)a->n + 1) * 4), *en = (int32_t *)malloc(((size_t)a->n + 1) * 4);
    oracle_row *rows = (oracle_row *)malloc(sizeof(oracle_row) * (size_t)a->rows_cap);
    a->events = 0; a->rc = 0;
    for (int r = a->tid; r < a->n_runs; r += a->n_threads) {
        int32_t nf = 0; int64_t nticks = 0;
        int rc = oracle_fifo_yarn(a->c, a->n, a->nt, a->duration, a->used_gpus, a->gpc, a->mem, a->util, fin, st, en, &nf, rows, a->rows_cap, &nticks, NULL);
        if (rc) { a->rc = rc; break; }
        a->events += 3 * (int64_t)nf;
    }
    free(fin); free(st); free(en); free(rows);
    return NULL;
}

int64_t oracle_fifo_yarn_batch(const oracle_cluster *c, int32_t n, const double *nt, const double *duration,
                               const double *used_gpus, const int32_t *gpc, const double *mem_max_mib, const double *util_avg,
                               int64_t rows_cap, int n_runs, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    batch_arg *args = (batch_arg *)malloc(sizeof(batch_arg) * (size_t)n_threads);
    for (int t = 0; t < n_threads; ++t) {
        batch_arg a = {c, n, nt, duration, used_gpus, mem_max_mib, gpc, util_avg, rows_cap, n_runs, n_threads, t, 0, 0};
        args[t] = a;
        pthread_create(&th[t], NULL, batch_worker, &args[t]);
    }
    int64_t ev = 0; int bad = 0;
    for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); ev += args[t].events; if (args[t].rc) bad = args[t].rc; }
    free(th); free(args);
    return bad ? bad : ev;
}

/* =============================================================================================
 * Legacy event-driven schedules: sjf family and dlas family.   PINNED GIVEN SHIMS (see the file header).
 *
 * The reference keeps these only as dead code (run_sim.py:162-287 smallest_first_sim_jobs, :299-431
 * shortest_first_sim_jobs, :664-947 dlas_sim_jobs); they read globals that are never defined (JOBS, CLUSTER, LOG,
 * scheduler).  oracle/ref_legacy_runner.py defines those at run time and calls the functions as they are;
 * this restatement equals the files those runs wrote (tests/test_oracle_golden.py, 26 fixtures), including
 * behaviour that only execution revealed: the 'end_jobs' list that run_sim.py:706-710 writes into the head
 * start event survives a queue-jump event and completes jobs early (see oracle_dlas).
 *
 * Build-defined inputs: submit_time = ceil(normalized_time) ticks,
 * duration = max(1, ceil(minutes*0.5)) ticks, num_gpu = ceil(used_gpus); job_events = arrivals
 * grouped by submit_time; move_to_runnable = status PENDING, last_check_time = event time, all
 * counters 0, start_time = "maxsize".
 * Row per event = the legacy cluster.csv columns (log.py:137-258): time, idle_node, full_node,
 * busy_gpu, pending_job, running_job, completed_job (busy_node / idle_gpu derive from them).
 * For a placement scheme the reference leaves the node/gpu columns at 0 (the code computing them is
 * commented out, log.py:171-196); here they carry the values that commented code would produce, and the
 * CSV formatters print 0 like the reference unless asked for them.
 * ========================================================================================== */

typedef struct {
    int32_t time, idle_nodes, full_nodes, busy_gpus, pending, running, completed, pad;
} legacy_row;

typedef struct {
    int32_t submit, duration, num_gpu;
    int32_t status;               /* 0 = not yet runnable, 1 PENDING, 2 RUNNING, 3 END */
    int32_t start_time, end_time; /* start_time -1 = sys.maxsize */
    int32_t total_executed, executed, pending_time, last_pending, remaining;
    int32_t q_id, preempt, resume, promote, last_check;
} ljob;

enum { L_NONE = 0, L_PENDING = 1, L_RUNNING = 2, L_END = 3 };

static void list_remove(int32_t *a, int32_t *len, int32_t v) { /* python list.remove(v) */
    for (int i = 0; i < *len; ++i) if (a[i] == v) { memmove(a + i, a + i + 1, (size_t)(*len - i - 1) * 4); (*len)--; return; }
}

/* sort key of the smallest / shortest family: 0 = num_gpu (sjf, run_sim.py:237), 1 = remaining_time (shortest, :385),
 * 2 = remaining_gputime = remaining_time * num_gpu (shortest-gpu, :382-383) */
static int64_t sort_key(const ljob *j, int mode) {
    return mode == 0 ? j->num_gpu : (mode == 1 ? (int64_t)j->remaining : (int64_t)j->remaining * j->num_gpu);
}

static void stable_sort_by_key(int32_t *a, int n, const ljob *jobs, int32_t *tmp, int mode) { /* list.sort(key=...): stable */
    for (int w = 1; w < n; w *= 2) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = (sort_key(&jobs[a[j]], mode) < sort_key(&jobs[a[i]], mode)) ? a[j++] : a[i++];
            while (i < mid) tmp[k++] = a[i++];
            while (j < hi) tmp[k++] = a[j++];
        }
        memcpy(a, tmp, (size_t)n * 4);
    }
}

static void legacy_node_stats(const osim *s, legacy_row *r) {
    int idle = 0, full = 0, busy_g = 0;
    for (int i = 0; i < s->N; ++i) {
        int fd = node_free_devices(s, i);
        busy_g += s->G - fd;
        if (fd == s->G) idle++; else if (fd == 0) full++;
    }
    r->idle_nodes = idle; r->full_nodes = full; r->busy_gpus = busy_g;
}

static void legacy_setup(osim *S, const oracle_cluster *c, int32_t n, const double *nt, const double *duration,
                         const double *used_gpus, const int32_t *gpc, const double *mem_max_mib, ljob *L) {
    memset(S, 0, sizeof *S);
    S->c = c; S->N = c->num_switch * c->num_node_p_switch; S->G = c->num_gpu_p_node; S->n = n;
    S->nodes = (onode *)calloc((size_t)S->N, sizeof(onode));
    S->devs = (odevice *)calloc((size_t)S->N * S->G, sizeof(odevice));
    S->jobs = (ojob *)calloc((size_t)n + 1, sizeof(ojob));
    int64_t total_tasks = 0;
    for (int i = 0; i < n; ++i) {
        ojob *j = &S->jobs[i];
        j->nt = nt[i]; j->duration = duration[i]; j->gpus = used_gpus[i]; j->gpc = gpc[i]; j->mem_mib = mem_max_mib[i];
        j->task_count = (int32_t)py_floordiv(used_gpus[i], (double)gpc[i]);
        total_tasks += j->task_count;
        L[i].submit = (int32_t)ceil(nt[i]);
        double d = ceil(duration[i]); L[i].duration = d < 1 ? 1 : (int32_t)d;
        L[i].num_gpu = (int32_t)ceil(used_gpus[i]);
        L[i].status = L_NONE; L[i].start_time = -1; L[i].end_time = -1;
    }
    S->task_pool = (int32_t *)malloc(((size_t)total_tasks + 1) * 4);
    int64_t off = 0;
    for (int i = 0; i < n; ++i) { S->jobs[i].task_node = S->task_pool + off; for (int t = 0; t < S->jobs[i].task_count; ++t) S->task_pool[off + t] = -1; off += S->jobs[i].task_count; }
}

static void legacy_teardown(osim *S) { free(S->nodes); free(S->devs); free(S->jobs); free(S->task_pool); }

static void legacy_empty_infra(osim *S) { /* CLUSTER.empty_infra(), infra/cluster.py:88-95 */
    memset(S->nodes, 0, sizeof(onode) * (size_t)S->N);
    memset(S->devs, 0, sizeof(odevice) * (size_t)S->N * S->G);
}

/* scheduler.try_get_job_res(CLUSTER, JOBS, job) with --scheme yarn: the live yarn fit (build-defined) */
static int legacy_try_get_job_res(osim *S, int job) {
    ojob *j = &S->jobs[job];
    for (int t = 0; t < j->task_count; ++t) j->task_node[t] = -1;
    int ok = (j->gpus > (double)S->c->num_gpu_p_node) ? try_cross_node_alloc_ms(S, job) : try_single_node_alloc_ms(S, job);
    return ok > 0;
}

/* Outputs per job (trace order): start, end (-1 = never), pending_time, preempt, resume; finish_order;
 * one row per event.  counters[0] = sum over events of runnable jobs (the per-event sweep size),
 * counters[1] = events (arrivals + finishes + preemptions + resumes/starts).
 * Returns 0, -1 = rows_cap too small. */
int oracle_sjf_yarn(const oracle_cluster *c, int32_t n, const double *nt, const double *duration, const double *used_gpus,
                    const int32_t *gpc, const double *mem_max_mib, int32_t sort_mode,
                    int32_t *finish_order, int32_t *start_tick, int32_t *end_tick, int32_t *pending_out, int32_t *preempt_out,
                    int32_t *resume_out, int32_t *n_finished, legacy_row *rows, int64_t rows_cap, int64_t *n_events, int64_t *counters) {
    osim S; ljob *L = (ljob *)calloc((size_t)n + 1, sizeof(ljob));
    legacy_setup(&S, c, n, nt, duration, used_gpus, gpc, mem_max_mib, L);
    int32_t *runnable = (int32_t *)malloc(((size_t)n + 1) * 4), *tmp = (int32_t *)malloc(((size_t)n + 1) * 4);
    int32_t *end_jobs = (int32_t *)malloc(((size_t)n + 1) * 4);
    int rlen = 0, n_end = 0, end_time_next = INT32_MAX, cursor = 0, flen = 0, rc = 0;
    int64_t nev = 0, sweeps = 0, ev_count = 0;
    while ((n - cursor) + rlen > 0) {                               /* run_sim.py:168 */
        if ((n - cursor) + n_end == 0) break;                         /* :169-171 "cluster is not large enough" */
        int start_time = cursor < n ? L[cursor].submit : INT32_MAX;   /* :174-181 */
        int end_time = n_end > 0 ? end_time_next : INT32_MAX;
        int event_time, has_end = 0, has_start = 0;
        if (end_time < start_time) { event_time = end_time; has_end = 1; }            /* :183-185 */
        else if (end_time > start_time) { event_time = start_time; has_start = 1; }   /* :186-189 */
        else { event_time = start_time; has_start = 1; has_end = 1; }                 /* :190-193 */
        if (has_end)                                                  /* :198-204 */
            for (int k = 0; k < n_end; ++k) {
                int e = end_jobs[k];
                L[e].status = L_END; L[e].end_time = event_time;      /* release_job_res + LOG.job_complete */
                list_remove(runnable, &rlen, e);
                finish_order[flen++] = e; ev_count++;
            }
        if (has_start)                                                /* :208-214 */
            while (cursor < n && L[cursor].submit == event_time) {
                ljob *j = &L[cursor];
                j->status = L_PENDING; j->last_check = event_time; j->remaining = j->duration;  /* move_to_runnable */
                runnable[rlen++] = cursor; cursor++; ev_count++;
            }
        for (int k = 0; k < rlen; ++k) {                              /* :216-230 */
            ljob *j = &L[runnable[k]];
            int d = event_time - j->last_check;
            if (j->status == L_RUNNING) { j->total_executed += d; j->remaining = j->duration - j->total_executed; }
            else if (j->status == L_PENDING) j->pending_time += d;
            j->last_check = event_time;
        }
        sweeps += rlen;
        stable_sort_by_key(runnable, rlen, L, tmp, sort_mode);        /* :237 / :380-385 */
        legacy_empty_infra(&S);                                       /* :241 */
        for (int k = 0; k < rlen; ++k) {                              /* :242-267 (status flips applied in place) */
            int job = runnable[k]; ljob *j = &L[job];
            if (legacy_try_get_job_res(&S, job)) {
                if (j->start_time == -1) j->start_time = event_time;
                if (j->status == L_PENDING) { j->status = L_RUNNING; j->resume += 1; ev_count++; }
            } else if (j->status == L_RUNNING) { j->status = L_PENDING; j->preempt += 1; ev_count++; }
        }
        end_time_next = INT32_MAX; n_end = 0;                         /* :270-284 */
        for (int k = 0; k < rlen; ++k) { ljob *j = &L[runnable[k]]; if (j->status == L_RUNNING && event_time + j->remaining < end_time_next) end_time_next = event_time + j->remaining; }
        for (int k = 0; k < rlen; ++k) { ljob *j = &L[runnable[k]]; if (j->status == L_RUNNING && event_time + j->remaining == end_time_next) end_jobs[n_end++] = runnable[k]; }
        if (nev >= rows_cap) { rc = -1; break; }                      /* :287 LOG.checkpoint */
        legacy_row *r = &rows[nev++]; memset(r, 0, sizeof *r);
        r->time = event_time; legacy_node_stats(&S, r);
        for (int k = 0; k < rlen; ++k) { if (L[runnable[k]].status == L_RUNNING) r->running++; else r->pending++; }
        r->completed = flen;
    }
    for (int i = 0; i < n; ++i) { start_tick[i] = L[i].start_time; end_tick[i] = L[i].end_time; pending_out[i] = L[i].pending_time; preempt_out[i] = L[i].preempt; resume_out[i] = L[i].resume; }
    *n_finished = flen; *n_events = nev;
    if (counters) { counters[0] = sweeps; counters[1] = ev_count; }
    free(runnable); free(tmp); free(end_jobs); free(L); legacy_teardown(&S);
    return rc;
}

/* dlas_sim_jobs(gputime, solve_starvation=0) with count-based admission (run_sim.py:664-947): gputime = 1 is
 * `--schedule dlas-gpu` (attained service in GPU-ticks), gputime = 0 is `--schedule dlas` (ticks, README.md:54).
 * queue_limit[num_queue-1] thresholds.  Extra output: promote (always 0 with solve_starvation=0),
 * counters[2] = demotions. */
int oracle_dlas(const oracle_cluster *c, int32_t n, const double *nt, const double *duration, const double *used_gpus,
                    int32_t gputime, int32_t num_queue, const int32_t *queue_limit,
                    int32_t *finish_order, int32_t *start_tick, int32_t *end_tick, int32_t *pending_out, int32_t *preempt_out,
                    int32_t *resume_out, int32_t *n_finished, legacy_row *rows, int64_t rows_cap, int64_t *n_events, int64_t *counters) {
    const int total_gpu = c->num_switch * c->num_node_p_switch * c->num_gpu_p_node;
    const int num_node = c->num_switch * c->num_node_p_switch;
    ljob *L = (ljob *)calloc((size_t)n + 1, sizeof(ljob));
    for (int i = 0; i < n; ++i) {
        L[i].submit = (int32_t)ceil(nt[i]);
        double d = ceil(duration[i]); L[i].duration = d < 1 ? 1 : (int32_t)d;
        L[i].num_gpu = (int32_t)ceil(used_gpus[i]); L[i].start_time = -1; L[i].end_time = -1;
    }
    int32_t *runnable = (int32_t *)malloc(((size_t)n + 1) * 4), *end_jobs = (int32_t *)malloc(((size_t)n + 1) * 4);
    int32_t *tmp = (int32_t *)malloc(((size_t)n + 1) * 4);
    int32_t *attached = (int32_t *)malloc(((size_t)n + 1) * 4); int n_att = -1;   /* 'end_jobs' key of the head start event, -1 = absent */
    int32_t **queues = (int32_t **)malloc(sizeof(int32_t *) * (size_t)num_queue); int32_t *qlen = (int32_t *)calloc((size_t)num_queue, 4);
    for (int q = 0; q < num_queue; ++q) queues[q] = (int32_t *)malloc(((size_t)n + 1) * 4);
    int rlen = 0, n_end = 0, end_time_next = INT32_MAX, cursor = 0, flen = 0, rc = 0, next_job_jump = INT32_MAX, free_gpu = total_gpu;
    int64_t nev = 0, sweeps = 0, ev_count = 0, demotions = 0;
    while ((n - cursor) + rlen > 0) {                                /* :679 */
        if ((n - cursor) + n_end == 0) break;                          /* :680-682 */
        int start_time = cursor < n ? L[cursor].submit : INT32_MAX;    /* :685-694 */
        int end_time = n_end > 0 ? end_time_next : INT32_MAX;
        int event_time, has_end = 0, has_start = 0;
        if (end_time < start_time) { event_time = end_time; has_end = 1; }             /* :699-701 */
        else if (end_time > start_time) { event_time = start_time; has_start = 1; }    /* :702-705 */
        else {                                                                         /* :706-710 */
            /* `event = start_event; event['end_jobs'] = end_events[0]['end_jobs']` writes the key into the dict that stays at
               JOBS.job_events[0]: when the jump test below replaces the event, the head start event keeps this list, and it is
               honoured when that start event is finally handled — even if the jobs in it were demoted / preempted meanwhile
               (they are completed early, whatever their status).  Pinned by tests/golden/dlasgpu_*. */
            event_time = start_time; has_start = 1; has_end = 1;
            memcpy(attached, end_jobs, (size_t)n_end * 4); n_att = n_end;
        }
        if (event_time > next_job_jump) { event_time = next_job_jump; has_end = has_start = 0; }  /* :715-717 */
        const int32_t *ending = end_jobs; int n_ending = has_end ? n_end : 0;
        if (has_start) { ending = attached; n_ending = n_att > 0 ? n_att : 0; }        /* 'end_jobs' in start_event */
        if (n_ending > 0)                                              /* :721-727 */
            for (int k = 0; k < n_ending; ++k) {
                int e = ending[k];
                free_gpu += L[e].num_gpu; if (free_gpu > total_gpu) free_gpu = total_gpu;   /* cluster.py:1462-1468 */
                L[e].status = L_END; L[e].end_time = event_time;
                list_remove(runnable, &rlen, e);
                list_remove(queues[L[e].q_id], &qlen[L[e].q_id], e);
                finish_order[flen++] = e; ev_count++;
            }
        if (has_start)                                                 /* :730-737 */
            while (cursor < n && L[cursor].submit == event_time) {
                ljob *j = &L[cursor];
                j->status = L_PENDING; j->last_check = event_time; j->q_id = 0;
                runnable[rlen++] = cursor; queues[0][qlen[0]++] = cursor; cursor++; ev_count++;
            }
        if (has_start) n_att = -1;                                     /* :737 JOBS.job_events.pop(0): the next start event is a fresh dict */
        for (int k = 0; k < rlen; ++k) {                               /* :740-792 */
            int job = runnable[k]; ljob *j = &L[job];
            int d = event_time - j->last_check;
            if (j->status == L_RUNNING) {
                j->total_executed += d; j->executed += d; j->last_check = event_time;
                int64_t j_gt = gputime ? (int64_t)j->executed * j->num_gpu : (int64_t)j->executed;   /* :748-752 */
                int cur = j->q_id;
                if (cur < num_queue - 1 && j_gt >= queue_limit[cur]) { /* :755-759 one level per event */
                    j->q_id = cur + 1;
                    queues[cur + 1][qlen[cur + 1]++] = job;
                    list_remove(queues[cur], &qlen[cur], job);
                    demotions++; ev_count++;
                }
            } else if (j->status == L_PENDING) {
                j->last_check = event_time; j->pending_time += d;
                if (j->executed > 0) j->last_pending += d;            /* :768-769 (solve_starvation = 0: no promotion) */
            }
        }
        sweeps += rlen;
        free_gpu = total_gpu;                                          /* :799 CLUSTER.empty_infra() */
        int n_run = 0, n_pre = 0;
        for (int q = 0; q < num_queue; ++q)                            /* :808-823 */
            for (int k = 0; k < qlen[q]; ++k) {
                int job = queues[q][k]; ljob *j = &L[job];
                if (free_gpu >= j->num_gpu) { if (j->status == L_PENDING) tmp[n_run++] = job; free_gpu -= j->num_gpu; }
                else if (j->status == L_RUNNING) end_jobs[n_pre++] = job;   /* end_jobs reused as the preempt list */
            }
        for (int k = 0; k < n_pre; ++k) { ljob *j = &L[end_jobs[k]]; j->status = L_PENDING; j->preempt += 1; ev_count++; }   /* :825-829 */
        for (int k = 0; k < n_run; ++k) { ljob *j = &L[tmp[k]]; j->status = L_RUNNING; j->resume += 1; ev_count++;        /* :830-834 */
                                          if (j->start_time == -1) j->start_time = event_time; }
        for (int q = 0; q < num_queue; ++q) {                          /* :838-848 pending jobs behind running ones */
            int w = 0, p = 0;
            for (int k = 0; k < qlen[q]; ++k) { int job = queues[q][k]; if (L[job].status == L_PENDING) tmp[p++] = job; else queues[q][w++] = job; }
            memcpy(queues[q] + w, tmp, (size_t)p * 4);
        }
        end_time_next = INT32_MAX; n_end = 0;                          /* :908-922 */
        for (int k = 0; k < rlen; ++k) { ljob *j = &L[runnable[k]]; if (j->status == L_RUNNING) { int et = event_time + j->duration - j->total_executed; if (et < end_time_next) end_time_next = et; } }
        for (int k = 0; k < rlen; ++k) { ljob *j = &L[runnable[k]]; if (j->status == L_RUNNING && event_time + j->duration - j->total_executed == end_time_next) end_jobs[n_end++] = runnable[k]; }
        next_job_jump = INT32_MAX;                                     /* :925-943 */
        for (int k = 0; k < rlen; ++k) {
            ljob *j = &L[runnable[k]];
            if (j->status == L_RUNNING && j->q_id < num_queue - 1) {
                /* int(math.ceil((queue_limit[q] - executed_time) / num_gpu) + event_time): as written, executed_time is
                   NOT multiplied by num_gpu here */
                int jt = gputime ? (int)ceil((double)(queue_limit[j->q_id] - j->executed) / (double)j->num_gpu) + event_time
                                 : queue_limit[j->q_id] - j->executed + event_time;                  /* :929-932 */
                if (jt < next_job_jump) next_job_jump = jt;
            }
        }
        if (nev >= rows_cap) { rc = -1; break; }                       /* :947 LOG.checkpoint, count branch log.py:225-238 */
        legacy_row *r = &rows[nev++]; memset(r, 0, sizeof *r);
        r->time = event_time; r->busy_gpus = total_gpu - free_gpu;
        int busy_node = (r->busy_gpus + c->num_gpu_p_node - 1) / c->num_gpu_p_node;
        r->full_nodes = busy_node; r->idle_nodes = num_node - busy_node;
        for (int k = 0; k < rlen; ++k) { if (L[runnable[k]].status == L_RUNNING) r->running++; else r->pending++; }
        r->completed = flen;
    }
    for (int i = 0; i < n; ++i) { start_tick[i] = L[i].start_time; end_tick[i] = L[i].end_time; pending_out[i] = L[i].pending_time; preempt_out[i] = L[i].preempt; resume_out[i] = L[i].resume; }
    *n_finished = flen; *n_events = nev;
    if (counters) { counters[0] = sweeps; counters[1] = ev_count; counters[2] = demotions; }
    for (int q = 0; q < num_queue; ++q) free(queues[q]);
    free(queues); free(qlen); free(runnable); free(end_jobs); free(tmp); free(attached); free(L);
    return rc;
}

/* =============================================================================================
 * Pack placements: `--schedule horus|gandiva` with horus_placement (any of --scheme horus|horus+|gandiva).
 *
 * Restates core/scheduling/algorithm.py:34-180 (horus_placement), :204-240 (schedule_horus), :189-202
 * (schedule_fifo, used by gandiva), core/scheduling/horus.py:6-56 (the two score functions; the score
 * function is keyed by the SCHEDULE name because schedule.py:47 passes self.schedule as `scheme`),
 * the pack=True branches of infra/node.py:146-221 and infra/device.py:19-77, the utilisation heap queue
 * (core/jobs/base_factory.py:1-12, job_queue_manager.py:129-154, stdlib heapq) and the interference
 * bookkeeping (infra/node.py:71-91, core/jobs/jobs_manager.py:175-187).
 *
 * Parity status: PINNED for traces whose jobs have gpu_utilization_max == gpu_utilization_avg: every
 * np.random.normal(loc, 0) of infra/device.py:30,52 then returns loc exactly and the reference is
 * deterministic (tests/golden/horus_*, gandiva_*).  With a spread the reference draws from an unseeded
 * numpy RNG; this file then uses the build-defined counter-based draw `pack_draw` below: UNPINNED.
 * ========================================================================================== */
#define RLGS_ORACLE_MAX_Q 8
typedef struct {
    osim S;                       /* must stay first: the helpers above take osim* */
    int gandiva;                  /* 0 = horus_score, 1 = gandiva_score */
    int yarn;                     /* 1 = --scheme yarn: ms_yarn_placement (algorithm.py:28-32) under the horus / gandiva schedule */
    /* horus+ (schedule_horus_plus, algorithm.py:242-290): K utilisation heaps filled by a k-means over the queued jobs */
    int plus, K;
    int32_t *pq[RLGS_ORACLE_MAX_Q]; int pqn[RLGS_ORACLE_MAX_Q]; double credits[RLGS_ORACLE_MAX_Q];
    const double *mem_avg;        /* memory_avg in MiB (Job.gpu_mem_avg) */
    uint32_t kseed, kcalls;       /* injected np.random.randint / np.random.choice: call counter (oracle/ref_runner.py _INJECT) */
    const double *util_max;
    int rng_on; uint32_t seed, replica;
    uint8_t **interf;             /* Task.interfered per job, per task */
    uint8_t *interf_pool;
    uint8_t **tbump;              /* per task: Task.duration == original_duration + 5 (jobs_manager.py:184-185) */
    uint8_t *tbump_pool;
    uint32_t *pj_bits;            /* node.placed_jobs membership, [job][W] */
    int W;
    int crashed;
} hsim;

static uint64_t pack_mix64(uint64_t z) { z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
/* Build-defined stand-in for np.random.normal: Irwin-Hall sum of twelve 16-bit uniforms keyed by
 * (seed, replica, tick, look-ahead position, task pass, device, slot); |z| <= 6, variance 1 - 2^-32. */
static double pack_draw(const hsim *H, uint32_t tick, uint32_t attempt, uint32_t pass, uint32_t dev, uint32_t slot) {
    uint64_t h = pack_mix64((((uint64_t)H->seed << 32) | H->replica) + 0x9E3779B97F4A7C15ull);
    h = pack_mix64(h ^ (((uint64_t)tick << 32) | ((uint64_t)attempt << 16) | pass));
    h = pack_mix64(h ^ (((uint64_t)dev << 8) | slot));
    uint32_t sum = 0;
    for (int k = 1; k <= 3; ++k) { uint64_t w = pack_mix64(h + (uint64_t)k * 0x9E3779B97F4A7C15ull); sum += (uint32_t)(w & 0xffff) + (uint32_t)((w >> 16) & 0xffff) + (uint32_t)((w >> 32) & 0xffff) + (uint32_t)(w >> 48); }
    return ((double)sum - 393210.0) / 65536.0;
}

/* stdlib heapq (Lib/heapq.py: _siftdown, _siftup, heappush, heappop) over int ids with a `<` callback */
typedef int (*lt_fn)(const void *ctx, int a, int b);
static void hq_siftdown(int32_t *h, int startpos, int pos, lt_fn lt, const void *ctx) {
    int newitem = h[pos];
    while (pos > startpos) { int pp = (pos - 1) >> 1, parent = h[pp]; if (lt(ctx, newitem, parent)) { h[pos] = parent; pos = pp; continue; } break; }
    h[pos] = newitem;
}
static void hq_siftup(int32_t *h, int len, int pos, lt_fn lt, const void *ctx) {
    int endpos = len, startpos = pos, newitem = h[pos], child = 2 * pos + 1;
    while (child < endpos) {
        int right = child + 1;
        if (right < endpos && !lt(ctx, h[child], h[right])) child = right;
        h[pos] = h[child]; pos = child; child = 2 * pos + 1;
    }
    h[pos] = newitem;
    hq_siftdown(h, startpos, pos, lt, ctx);
}
static void hq_push(int32_t *h, int *len, int item, lt_fn lt, const void *ctx) { h[(*len)++] = item; hq_siftdown(h, 0, *len - 1, lt, ctx); }
static int hq_pop(int32_t *h, int *len, lt_fn lt, const void *ctx) {
    int last = h[--(*len)];
    if (*len > 0) { int ret = h[0]; h[0] = last; hq_siftup(h, *len, 0, lt, ctx); return ret; }
    return last;
}
/* CompareAbleByUtilization.__lt__ (base_factory.py:8-12): `if self.avg: return self.avg < other.avg; return False` */
static int lt_job_util(const void *ctx, int a, int b) { const osim *s = (const osim *)ctx; double ua = s->jobs[a].util_avg; return ua != 0 ? ua < s->jobs[b].util_avg : 0; }
/* NodeDeviceInfo.__lt__ (algorithm.py:25-26): self.min_score > other.min_score */
static int lt_node_info(const void *ctx, int a, int b) { const double *sc = (const double *)ctx; return sc[a] > sc[b]; }

static void pj_set(hsim *H, int node, int job) { uint32_t *w = &H->pj_bits[(size_t)job * H->W + (node >> 5)], b = 1u << (node & 31); if (!(*w & b)) { *w |= b; H->S.nodes[node].placed_jobs++; } }
static int pj_pop(hsim *H, int node, int job) { uint32_t *w = &H->pj_bits[(size_t)job * H->W + (node >> 5)], b = 1u << (node & 31); if (!(*w & b)) return 0; *w &= ~b; H->S.nodes[node].placed_jobs--; return 1; }

/* Device.get_current_utilization (infra/device.py:48-54) */
static double pack_device_util(const hsim *H, const odevice *d, int dev, uint32_t tick, uint32_t attempt, uint32_t pass) {
    double util = 0;
    for (int k = 0; k < d->ntasks; ++k) {
        const ojob *t = &H->S.jobs[d->owner_job[k]];
        double x = t->util_avg;
        if (H->rng_on) { double sd = (H->util_max[d->owner_job[k]] - t->util_avg) / 2; if (sd != 0) x = t->util_avg + sd * pack_draw(H, tick, attempt, pass, (uint32_t)dev, (uint32_t)k); }
        util += (x < 100) ? x : 100;
        util = (util < 100) ? util : 100;
    }
    return util;
}
/* Device.add_task(pack=True) (infra/device.py:26-46): the interference slowdown is computed and dropped, only the flag stays */
static int pack_device_add_task(hsim *H, odevice *d, int job, int task) {
    if (!device_can_fit(&H->S, d, &H->S.jobs[job])) return 0;
    H->interf[job][task] = d->ntasks >= 2;
    if (d->ntasks < 2) H->tbump[job][task] = 0;                /* else-branch: task.duration = task.original_duration (device.py:39-41) */
    for (int k = 0; k < d->ntasks; ++k) if (d->owner_job[k] == job && d->owner_task[k] == task) return 1;  /* dict overwrite of a leaked entry */
    d->owner_job[d->ntasks] = job; d->owner_task[d->ntasks] = task; d->ntasks++;
    return 1;
}
/* Node.can_fit(pack=True) (infra/node.py:146-171) */
static int pack_node_can_fit(const hsim *H, int i, const ojob *j) {
    const osim *s = &H->S;
    if (node_cpu_free(s, i) - CPUS_PER_TASK < 0 || node_mem_free(s, i) - MEM_PER_TASK < 0) return 0;
    for (int g = 0; g < s->G; ++g) if (device_can_fit(s, &s->devs[i * s->G + g], j)) return 1;
    return 0;
}
/* Node.try_reserve_and_placed_task(pack=True) (infra/node.py:200-221): cpu/mem and the accepting devices stay
 * charged when fewer than task.gpu devices take the task */
static int pack_try_reserve(hsim *H, int i, int job, int task) {
    osim *s = &H->S; const ojob *j = &s->jobs[job];
    if (!pack_node_can_fit(H, i, j)) return 0;
    s->nodes[i].cpu_used += CPUS_PER_TASK; s->nodes[i].mem_used += MEM_PER_TASK;
    int should = j->gpc;
    for (int g = 0; g < s->G; ++g) { if (should <= 0) break; if (pack_device_add_task(H, &s->devs[i * s->G + g], job, task)) should--; }
    if (should == 0) s->nodes[i].placed_tasks++;
    return should == 0;
}
/* horus_score / gandiva_score (core/scheduling/horus.py:28-56 / :6-25); returns min_cost */
static double pack_score_node(const hsim *H, int i, const ojob *j, uint32_t tick, uint32_t attempt, uint32_t pass) {
    const osim *s = &H->S; double cap = s->c->gpu_memory_capacity_mib, min_cost = 999;
    for (int g = 0; g < s->G; ++g) {
        const odevice *d = &s->devs[i * s->G + g];
        if (!device_can_fit(s, d, j)) continue;
        double cost, cur_mem = device_current_memory(s, d), cur_util = pack_device_util(H, d, i * s->G + g, tick, attempt, pass);
        if (!H->gandiva) {
            double mem_cost = (cur_mem + j->mem_mib) / cap, x = cur_util + j->util_avg;
            double y = 0.0 * x + 4E-5; y = y * x + -0.00302; y = y * x + 1.16664;   /* np.polyval(NV_2080_COEF, x): Horner */
            cost = (mem_cost * 0.5) + (y * 0.5) + (double)d->ntasks;
        } else {
            double mem_cost = cur_mem + j->mem_mib / cap;
            cost = (mem_cost * 0.5) + (cur_util / 100) + (double)d->ntasks;
        }
        if (cost < min_cost) min_cost = cost;
    }
    return min_cost;
}
/* release at completion: Node.release_allocated_resources + JobsManager.reset_interference */
static void pack_release_finished(hsim *H, int nd, int job, int task) {
    osim *s = &H->S;
    node_release_allocated_resources(s, nd, job, task);
    for (int g = 0; g < s->G; ++g) {
        odevice *d = &s->devs[nd * s->G + g];
        if (d->ntasks > 1) continue;
        for (int k = 0; k < d->ntasks; ++k) {
            int oj = d->owner_job[k], ot = d->owner_task[k];
            if (H->interf[oj][ot] && s->jobs[oj].running) { H->interf[oj][ot] = 0; H->tbump[oj][ot] = 1; }   /* duration = original + max(int(diff / 2), 5), diff = 0 or 5 */
        }
    }
}

static double pack_bump(const hsim *H, int job) {   /* Job.get_duration() - Job.duration (job.py:206-210) */
    for (int t = 0; t < H->S.jobs[job].task_count; ++t) if (H->tbump[job][t]) return 5.0;
    return 0.0;
}

/* horus_placement (algorithm.py:34-180).  Returns 1 placed, 0 not placed, -2 the reference would raise. */
static int horus_placement(hsim *H, int job, uint32_t tick, uint32_t attempt) {
    osim *s = &H->S; ojob *j = &s->jobs[job];
    if (H->yarn) {   /* placement_algorithms['yarn'] under these schedules: the same fit as fifo + yarn */
        int ok = (j->gpus > (double)s->c->num_gpu_p_node) ? try_cross_node_alloc_ms(s, job) : try_single_node_alloc_ms(s, job);
        return ok < 0 ? -2 : ok;
    }
    const int N = s->N, T = j->task_count;
    int cap_entries = T * N + 1;
    double *score = (double *)malloc(sizeof(double) * (size_t)cap_entries);
    int32_t *enode = (int32_t *)malloc(4 * (size_t)cap_entries), *heap = (int32_t *)malloc(4 * (size_t)cap_entries);
    int ne = 0, hlen = 0;
    for (int t = 0; t < T; ++t)
        for (int i = 0; i < N; ++i) {
            if (!node_is_free(s, i)) continue;                 /* infrastructure.get_free_nodes() */
            if (!pack_node_can_fit(H, i, j)) continue;
            score[ne] = pack_score_node(H, i, j, tick, attempt, (uint32_t)t); enode[ne] = i;
            hq_push(heap, &hlen, ne, lt_node_info, score); ne++;
            if ((double)hlen > j->gpus) (void)hq_pop(heap, &hlen, lt_node_info, score);
        }
    /* sorted(nodes_stack, key=min_score): stable insertion sort of the heap array */
    for (int a = 1; a < hlen; ++a) { int v = heap[a], b = a - 1; while (b >= 0 && score[heap[b]] > score[v]) { heap[b + 1] = heap[b]; --b; } heap[b + 1] = v; }
    int32_t *map = (int32_t *)malloc(4 * (size_t)(hlen * T + 1)), *nmapped = (int32_t *)calloc((size_t)hlen + 1, 4), *nnodes = (int32_t *)calloc((size_t)hlen + 1, 4);
    const int R = s->c->num_switch, P = s->c->num_node_p_switch;
    int32_t *racks = (int32_t *)malloc(4 * (size_t)R);
    for (int e = 0; e < hlen; ++e) {
        int32_t *m = map + (size_t)e * T; int cnt = 0;
        int home = enode[heap[e]];
        for (int t = 0; t < T; ++t)
            if (pack_try_reserve(H, home, job, t)) { pj_set(H, home, job); m[cnt++] = home; }   /* successes are a prefix: tasks are identical */
        /* get_racks_by_dist (infra/infrastructure.py:135-147): stable sort of rack ids by |id - home rack| */
        int hr = home / P;
        for (int r = 0; r < R; ++r) racks[r] = r;
        for (int a = 1; a < R; ++a) { int v = racks[a], b = a - 1; while (b >= 0 && abs(racks[b] - hr) > abs(v - hr)) { racks[b + 1] = racks[b]; --b; } racks[b + 1] = v; }
        for (int ri = 0; ri < R; ++ri) {
            if (cnt >= T) break;
            for (int i = racks[ri] * P; i < (racks[ri] + 1) * P; ++i) {
                if (cnt >= T) break;
                for (int t = cnt; t < T; ++t) {              /* every task not yet in the mapping tries this node, in task order */
                    if (pack_try_reserve(H, i, job, t)) {
                        /* tasks are identical and a failed attempt only fills the node further, so a success cannot
                           follow a failure on one node: the mapped tasks stay the prefix 0..cnt-1 */
                        if (t != cnt) H->crashed = 1;
                        pj_set(H, i, job); m[cnt++] = i;
                    }
                    if (cnt >= T) break;
                }
            }
        }
        nmapped[e] = cnt;
        for (int t = 0; t < cnt; ++t) { int seen = 0; for (int u = 0; u < t; ++u) seen |= m[u] == m[t]; nnodes[e] += !seen; }
        /* clear previously reserved for backtracking (algorithm.py:122-133) */
        for (int t = 0; t < cnt; ++t) {
            if (t == 0) pj_pop(H, m[0], job);
            s->nodes[m[t]].placed_tasks--;
            node_release_allocated_resources(s, m[t], job, t);
        }
    }
    int best = -1;
    for (int e = 0; e < hlen; ++e) if (nmapped[e] >= T && (best < 0 || nnodes[e] < nnodes[best])) best = e;
    int rc = 0;
    if (best >= 0) {
        int32_t *m = map + (size_t)best * T; int cnt = 0;
        for (int t = 0; t < T; ++t)
            if (pack_try_reserve(H, m[t], job, t)) { j->task_node[t] = m[t]; pj_set(H, m[t], job); cnt++; }
        rc = (cnt == T) ? 1 : -2;                               /* assert cnt == len(next_job.tasks) (algorithm.py:177) */
    }
    free(score); free(enode); free(heap); free(map); free(nmapped); free(nnodes); free(racks);
    return rc;
}

static void pack_start_job(osim *s, int job, int delta) {   /* add_to_running -> start_job -> execute_job -> try_execute */
    ojob *j = &s->jobs[job];
    for (int t = 0; t < j->task_count; ++t) { int nd = j->task_node[t]; s->nodes[nd].placed_tasks--; s->nodes[nd].running_tasks++; }
    j->start_time = delta; j->migration_count += 1; j->running = 1;
    s->running[s->rlen++] = job;
}

/* ---- horus+: core/jobs/utils.py:36-67 (clusterize), job_queue_manager.py:103-127 (credits), jobs_manager.py:115-140 (insert) */
static uint32_t plus_draw(hsim *H, uint32_t call, uint32_t elem, uint32_t n) {   /* = _draw of oracle/ref_runner.py */
    uint64_t h = pack_mix64((((uint64_t)H->kseed << 32) | call) + 0x9E3779B97F4A7C15ull);
    h = pack_mix64(h ^ (uint64_t)elem);
    return (uint32_t)((h >> 11) % n);
}
static double plus_job_dist(const hsim *H, int x, int y) {   /* job_dist, utils.py:4-12: left-to-right sum of absolute differences */
    const ojob *a = &H->S.jobs[x], *b = &H->S.jobs[y];
    double score = (double)abs(a->task_count - b->task_count);
    score += fabs(a->util_avg - b->util_avg);
    score += (double)abs(a->gpc - b->gpc);
    score += fabs(a->gpus - b->gpus);
    score += fabs(H->util_max[x] - H->util_max[y]);
    score += fabs(H->mem_avg[x] - H->mem_avg[y]);
    score += fabs(a->mem_mib - b->mem_mib);
    return score;
}
static double plus_transform(const hsim *H, int x) {         /* transform_to_dist, utils.py:14-22 */
    const ojob *a = &H->S.jobs[x];
    double score = (double)a->task_count;
    score += a->util_avg; score += (double)a->gpc; score += a->gpus; score += H->util_max[x]; score += H->mem_avg[x]; score += a->mem_mib;
    return score;
}
static double np_pairwise_sum(const double *a, int n) {       /* numpy DOUBLE_pairwise_sum (np.mean = (0 + this) / n) */
    if (n < 8) { double res = 0.; for (int i = 0; i < n; ++i) res += a[i]; return res; }
    if (n <= 128) {
        double r[8]; int i;
        for (i = 0; i < 8; ++i) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8) for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2; n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}
/* clusterize(jobs, k): assign[i] = queue of jobs[i] */
static void plus_clusterize(hsim *H, const int32_t *jobs, int n, int32_t *assign, int32_t *old, double *tmp) {
    const int K = H->K;
    int cent[RLGS_ORACLE_MAX_Q];
    { uint32_t call = H->kcalls++; for (int c = 0; c < K; ++c) cent[c] = jobs[plus_draw(H, call, (uint32_t)c, (uint32_t)n)]; }   /* np.random.randint(len(jobs), size=k) */
    for (int i = 0; i < n; ++i) { assign[i] = -1; old[i] = -1; }
    int iter = 0;
    while (iter < 1000) {
        int same = 1; for (int i = 0; i < n; ++i) same &= assign[i] == old[i];
        if (same && iter != 0) break;
        memcpy(old, assign, 4 * (size_t)n);
        iter++;
        for (int i = 0; i < n; ++i) {                          /* np.argmin: first minimum */
            int best = 0; double bd = plus_job_dist(H, jobs[i], cent[0]);
            for (int c = 1; c < K; ++c) { double d = plus_job_dist(H, jobs[i], cent[c]); if (d < bd) { bd = d; best = c; } }
            assign[i] = best;
        }
        for (int c = 0; c < K; ++c) {
            int m = 0;
            for (int i = 0; i < n; ++i) if (assign[i] == c) tmp[m++] = plus_transform(H, jobs[i]);
            if (m > 0) {
                double mean = (0.0 + np_pairwise_sum(tmp, m)) / (double)m;
                long long score = (long long)mean;             /* .astype(int): truncation */
                int pick = -1; double cs = 99999999999.0;       /* get_closest, utils.py:24-34 */
                for (int i = 0; i < n; ++i) if (assign[i] == c) { double t = fabs(plus_transform(H, jobs[i]) - (double)score); if (t < cs) { cs = t; pick = jobs[i]; } }
                if (pick < 0) { H->crashed = 1; return; }      /* assert c is not None */
                cent[c] = pick;
            } else cent[c] = jobs[plus_draw(H, H->kcalls++, 0, (uint32_t)n)];   /* np.random.choice(len(jobs)) */
        }
    }
}
static void plus_update_credits(hsim *H, double *tmp) {       /* job_queue_manager.py:115-127 */
    for (int q = 0; q < H->K; ++q) {
        int n = H->pqn[q];
        if (n == 0) { H->credits[q] = 0; continue; }
        for (int i = 0; i < n; ++i) tmp[i] = H->S.jobs[H->pq[q][i]].pending_time;
        qsort(tmp, (size_t)n, sizeof(double), cmp_double);
        double med = (n & 1) ? tmp[n / 2] : (tmp[n / 2 - 1] + tmp[n / 2]) / 2.0;
        double mp = med > 0 ? med : 0;
        H->credits[q] = mp < 1 ? (double)n : mp * (double)n;
    }
}

/*
 * Scheduler.start() (schedule.py:178-216) with schedule = horus (gandiva = 0) or gandiva (= 1).
 * Same inputs / outputs as oracle_env_yarn plus util_max, the look-ahead `num_buffer` (flags --num_buffer,
 * run_sim.py:76) and dur_out[i] = Job.get_duration() at the end (original + 5 when a task was de-interfered).
 * rng_on = 0 pins every draw to its mean (= the reference on zero-spread traces).
 */
int oracle_pack(const oracle_cluster *c, int32_t n, const double *nt, const double *duration,
                const double *used_gpus, const int32_t *gpc, const double *mem_max_mib,
                const double *util_avg, const double *util_max, int32_t gandiva, int32_t yarn, int32_t num_buffer,
                int32_t rng_on, uint32_t seed, uint32_t replica,
                int32_t *finish_order, int32_t *start_tick, int32_t *end_tick, int32_t *n_finished, double *dur_out,
                int32_t *jct_out, int32_t *starts_out,
                int32_t num_queue_plus, const double *mem_avg_mib, uint32_t inject_seed,
                oracle_row *rows, int64_t rows_cap, int64_t *n_ticks, int64_t *counters) {
    hsim HS; memset(&HS, 0, sizeof HS);
    hsim *H = &HS; osim *S = &H->S;
    S->c = c; S->N = c->num_switch * c->num_node_p_switch; S->G = c->num_gpu_p_node; S->n = n;
    S->nodes = (onode *)calloc((size_t)S->N, sizeof(onode));
    S->devs = (odevice *)calloc((size_t)S->N * S->G, sizeof(odevice));
    S->jobs = (ojob *)calloc((size_t)n + 1, sizeof(ojob));
    S->queue = (int32_t *)malloc(((size_t)n + 1) * 4); S->running = (int32_t *)malloc(((size_t)n + 1) * 4);
    S->fin = finish_order;
    H->gandiva = gandiva; H->yarn = yarn; H->util_max = util_max; H->rng_on = rng_on; H->seed = seed; H->replica = replica;
    H->W = (S->N + 31) / 32;
    H->plus = num_queue_plus > 0; H->K = num_queue_plus > RLGS_ORACLE_MAX_Q ? RLGS_ORACLE_MAX_Q : (num_queue_plus > 0 ? num_queue_plus : 0);
    H->mem_avg = mem_avg_mib; H->kseed = inject_seed;
    for (int q = 0; q < H->K; ++q) H->pq[q] = (int32_t *)malloc(((size_t)n + 1) * 4);
    int32_t *kjobs = (int32_t *)malloc(((size_t)n + 1) * 4), *kassign = (int32_t *)malloc(((size_t)n + 1) * 4), *kold = (int32_t *)malloc(((size_t)n + 1) * 4);
    int32_t *lookq = (int32_t *)malloc(4 * (size_t)(num_buffer > 0 ? num_buffer : 1));
    H->pj_bits = (uint32_t *)calloc((size_t)(n + 1) * H->W, 4);
    H->tbump = (uint8_t **)calloc((size_t)n + 1, sizeof(uint8_t *));
    H->interf = (uint8_t **)calloc((size_t)n + 1, sizeof(uint8_t *));
    double *scratch = (double *)malloc(((size_t)n + 1) * 8);
    int64_t total_tasks = 0; int rc = 0;
    for (int i = 0; i < n; ++i) {
        ojob *j = &S->jobs[i];
        j->nt = nt[i]; j->duration = duration[i]; j->gpus = used_gpus[i]; j->gpc = gpc[i];
        j->mem_mib = mem_max_mib[i]; j->util_avg = util_avg[i];
        j->submit_time = (int32_t)nt[i];
        if (gpc[i] <= 0) { rc = -2; goto done; }
        j->task_count = (int32_t)py_floordiv(used_gpus[i], (double)gpc[i]);
        if (j->task_count <= 0) { rc = -2; goto done; }
        total_tasks += j->task_count;
        start_tick[i] = -1; end_tick[i] = -1;
    }
    S->task_pool = (int32_t *)malloc(((size_t)total_tasks + 1) * 4);
    H->interf_pool = (uint8_t *)calloc((size_t)total_tasks + 1, 1);
    H->tbump_pool = (uint8_t *)calloc((size_t)total_tasks + 1, 1);
    { int64_t off = 0; for (int i = 0; i < n; ++i) { S->jobs[i].task_node = S->task_pool + off; H->interf[i] = H->interf_pool + off; H->tbump[i] = H->tbump_pool + off; for (int t = 0; t < S->jobs[i].task_count; ++t) S->task_pool[off + t] = -1; off += S->jobs[i].task_count; } }
    int32_t *look = (int32_t *)malloc(4 * (size_t)(num_buffer > 0 ? num_buffer : 1));
    uint8_t *fin_mark = (uint8_t *)malloc((size_t)n + 1);

    int64_t delta = 0, sumQ = 0, sumR = 0, starts = 0;
    int remaining = n, running_jobs = 0;
    while (remaining + running_jobs > 0) {
        int k0 = S->cursor;
        while (S->cursor < n && S->jobs[S->cursor].nt <= (double)delta) S->cursor++;
        if (H->plus) {
            /* jobs_manager.insert (:115-140) runs every tick, also with no arrival: every queue is emptied (heappop order, queue
               after queue), the new jobs are appended, the whole list is clustered again and pushed back */
            int m = 0;
            for (int q = 0; q < H->K; ++q) { while (H->pqn[q] > 0) kjobs[m++] = hq_pop(H->pq[q], &H->pqn[q], lt_job_util, S); }
            for (int i = k0; i < S->cursor; ++i) kjobs[m++] = i;
            if (m > 0) {
                plus_clusterize(H, kjobs, m, kassign, kold, scratch);
                if (H->crashed) { rc = -2; goto done2; }
                for (int i = 0; i < m; ++i) { int q = kassign[i]; H->credits[q] += 1; hq_push(H->pq[q], &H->pqn[q], kjobs[i], lt_job_util, S); }
            }
            S->qlen = m;
        } else
        if (!gandiva) for (int i = k0; i < S->cursor; ++i) hq_push(S->queue, &S->qlen, i, lt_job_util, S);   /* heappush per job (job_queue_manager.py:147-152) */
        else if (S->cursor > k0) {                                                                              /* list.insert(i, job): batch at the front (q1) */
            int k = S->cursor - k0;
            memmove(S->queue + k, S->queue, (size_t)S->qlen * 4);
            for (int i = 0; i < k; ++i) S->queue[i] = k0 + i;
            S->qlen += k;
        }
        if (S->qlen > 0) {                                             /* _schedule (schedule.py:40-60) */
            int nfree = 0;
            for (int i = 0; i < S->N; ++i) nfree += node_is_free(S, i);
            if (nfree >= 1) {
                if (H->plus) {                                         /* schedule_horus_plus (algorithm.py:242-290) */
                    int min_k = num_buffer < S->qlen ? num_buffer : S->qlen; if (min_k < 0) min_k = 0;
                    for (int a = 0; a < min_k; ++a) {
                        plus_update_credits(H, scratch);
                        int qi = 0; for (int q = 1; q < H->K; ++q) if (H->credits[q] > H->credits[qi]) qi = q;     /* np.argmax: first maximum */
                        if (H->pqn[qi] == 0) { rc = -2; goto done2; }   /* pop() returns None: AttributeError at j.is_waiting() */
                        look[a] = hq_pop(H->pq[qi], &H->pqn[qi], lt_job_util, S); lookq[a] = qi;
                    }
                    int pos = -1;
                    for (int a = 0; a < min_k && pos < 0; ++a) {
                        int r = horus_placement(H, look[a], (uint32_t)delta, (uint32_t)a);
                        if (r < 0 || H->crashed) { rc = -2; goto done2; }
                        if (r) pos = a;
                    }
                    for (int a = 0; a < min_k; ++a) if (a != pos) { H->credits[lookq[a]] += 1; hq_push(H->pq[lookq[a]], &H->pqn[lookq[a]], look[a], lt_job_util, S); }
                    if (pos >= 0) { S->qlen--; pack_start_job(S, look[pos], (int)delta); starts++; }
                } else
                if (!gandiva) {                                        /* schedule_horus (algorithm.py:204-240) */
                    int min_k = num_buffer < S->qlen ? num_buffer : S->qlen; if (min_k < 0) min_k = 0;
                    for (int a = 0; a < min_k; ++a) look[a] = hq_pop(S->queue, &S->qlen, lt_job_util, S);
                    int pos = -1;
                    for (int a = 0; a < min_k && pos < 0; ++a) {
                        int r = horus_placement(H, look[a], (uint32_t)delta, (uint32_t)a);
                        if (r < 0 || H->crashed) { rc = -2; goto done2; }
                        if (r) pos = a;
                    }
                    for (int a = 0; a < min_k; ++a) if (a != pos) hq_push(S->queue, &S->qlen, look[a], lt_job_util, S);
                    if (pos >= 0) { pack_start_job(S, look[pos], (int)delta); starts++; }
                } else {                                               /* schedule_fifo (algorithm.py:189-202) */
                    int job = S->queue[0];
                    if (S->jobs[job].submit_time <= delta) {
                        int r = horus_placement(H, job, (uint32_t)delta, 0);
                        if (r < 0 || H->crashed) { rc = -2; goto done2; }
                        if (r) { memmove(S->queue, S->queue + 1, (size_t)(S->qlen - 1) * 4); S->qlen--; pack_start_job(S, job, (int)delta); starts++; }
                    }
                }
            }
        }
        remaining = n - S->cursor;
        delta += 1;
        if (H->plus) {                                                 /* the statistics of construct_info read S->queue: keep it as the concatenation */
            int m = 0;
            for (int q = 0; q < H->K; ++q) for (int i = 0; i < H->pqn[q]; ++i) S->queue[m++] = H->pq[q][i];
            S->qlen = m;
        }
        for (int q = 0; q < S->qlen; ++q) S->jobs[S->queue[q]].pending_time += 1;
        for (int r = 0; r < S->rlen; ++r) S->jobs[S->running[r]].time_processed += 1;
        if (H->plus) plus_update_credits(H, scratch);                  /* jobs_manager.step (:143-148) */
        /* release_finished_jobs: the list of finishing jobs is fixed first (prepare_finish_tasks, jobs_manager.py:243-250) */
        for (int r = 0; r < S->rlen; ++r) { int job = S->running[r]; fin_mark[r] = !((double)S->jobs[job].time_processed < S->jobs[job].duration + pack_bump(H, job)); }
        int w = 0;
        for (int r = 0; r < S->rlen; ++r) {
            int job = S->running[r]; ojob *j = &S->jobs[job];
            if (!fin_mark[r]) { S->running[w++] = job; continue; }
            for (int t = 0; t < j->task_count; ++t) {
                int nd = j->task_node[t];
                S->nodes[nd].running_tasks--;
                pack_release_finished(H, nd, job, t);
                j->tasks_finished++;
            }
            j->running = 0; j->finished = 1; j->end_time = (int32_t)delta;
            S->fin[S->flen++] = job;
        }
        S->rlen = w; running_jobs = S->rlen;                           /* schedule.py:195: counted BEFORE the post-tick plugin */
        if (gandiva && S->qlen > 0) {                                  /* time_slice_check (algorithm.py:420-440), quanta = 100 */
            int ntodo = 0;
            for (int r = 0; r < S->rlen; ++r) { int tp = S->jobs[S->running[r]].time_processed; if (tp > 1 && tp % 100 == 0) { fin_mark[r] = 1; ntodo++; } else fin_mark[r] = 0; }
            if (ntodo) {
                int w2 = 0, nr = S->rlen;
                int32_t *todo = (int32_t *)malloc(4 * (size_t)ntodo); int nt2 = 0;
                for (int r = 0; r < nr; ++r) { if (fin_mark[r]) todo[nt2++] = S->running[r]; else S->running[w2++] = S->running[r]; }
                S->rlen = w2;                                          /* every preempted job leaves running_jobs (jobs_manager.py:151) ... */
                /* ... one after the other: a job preempted later is still "running" while an earlier one releases */
                for (int k = 0; k < nt2; ++k) S->jobs[todo[k]].running = 1;
                for (int k = 0; k < nt2; ++k) {
                    int job = todo[k]; ojob *j = &S->jobs[job];
                    j->running = 0;                                    /* popped from running_jobs first */
                    for (int t = 0; t < j->task_count; ++t) {
                        int nd = j->task_node[t];
                        if (!H->yarn) pj_pop(H, nd, job);              /* placed_jobs.pop(job_id) once per node */
                        else { int seen = 0; for (int u = 0; u < t; ++u) seen |= j->task_node[u] == nd; if (!seen) S->nodes[nd].placed_jobs--; }
                        S->nodes[nd].running_tasks--;
                        pack_release_finished(H, nd, job, t);          /* release(reserved=True) + reset_interference of the others */
                    }
                    if (H->yarn) for (int t = 0; t < j->task_count; ++t) j->task_node[t] = -1;   /* the yarn fit tracks its own assignment dict (algorithm.py:312) */
                    j->pending_time = 0;                               /* Job.preempted (job.py:177-181) */
                    memmove(S->queue + 1, S->queue, (size_t)S->qlen * 4); S->queue[0] = job; S->qlen++;   /* insert([job]): front */
                }
                free(todo);
            }
        }
        if (delta > rows_cap) { rc = -1; goto done2; }
        construct_info(S, &rows[delta - 1], scratch);
        sumQ += S->qlen; sumR += S->rlen;
    }
    for (int i = 0; i < n; ++i) {
        if (dur_out) dur_out[i] = S->jobs[i].duration + pack_bump(H, i);
        if (jct_out) jct_out[i] = S->jobs[i].time_processed;
        if (starts_out) starts_out[i] = S->jobs[i].migration_count;
        if (S->jobs[i].finished) { start_tick[i] = S->jobs[i].start_time; end_tick[i] = S->jobs[i].end_time; }
        else if (S->jobs[i].running) start_tick[i] = S->jobs[i].start_time;
    }
    *n_finished = S->flen; *n_ticks = delta;
    if (counters) { counters[0] = sumQ; counters[1] = sumR; counters[2] = delta; counters[3] = starts; }
done2:
    free(look); free(fin_mark);
done:
    for (int q = 0; q < H->K; ++q) free(H->pq[q]);
    free(kjobs); free(kassign); free(kold); free(lookq);
    free(S->nodes); free(S->devs); free(S->jobs); free(S->queue); free(S->running); free(scratch); free(S->task_pool);
    free(H->pj_bits); free(H->tbump); free(H->tbump_pool); free(H->interf); free(H->interf_pool);
    return rc;
}
