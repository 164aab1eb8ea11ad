/*
 * rlgs.h — C ABI of the B200-native cluster-simulator hot path (librlgs.so).
 *
 * This is the drop-in boundary for the reference's per-tick / per-event advance.  Every entry
 * point names the reference interface it replaces (paths are relative to the reference repo,
 * matthewygf/RLGPUSchedule); INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *  - plain C types only; every function returns an int32 status (RLGS_OK or a negative RLGS_ERR_*);
 *    rlgs_last_error() gives a thread-local message; no exception crosses the boundary.
 *  - the opaque rlgs_sim handle owns all device memory and its pinned host staging; host buffers
 *    passed in are caller-owned and never retained after the call returns.
 *  - a handle is not thread-safe; distinct handles may be used from distinct host threads.
 *  - a "replica" is one independent simulation (one trace on one simulated cluster); replicas never
 *    communicate.  On the device a group of 8, 16 or 32 lanes of a warp advances one replica (opts.lanes_per_replica).
 *  - the library needs a CUDA device: there is no CPU fallback (rlgs_create fails with RLGS_ERR_CUDA).
 */
#ifndef RLGS_H
#define RLGS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLGS_VERSION 200 /* 0.2.0: rlgs_opts gained rows_format / lanes_per_replica (appended), RLGS_ERR_SLOTS / RLGS_ERR_WIRE, rlgs_row16 */

enum {
    RLGS_OK = 0,
    RLGS_ERR_BAD_ARG = -1,
    RLGS_ERR_CUDA = -2,
    RLGS_ERR_OOM = -3,
    RLGS_ERR_UNSUPPORTED = -4, /* policy / option the device path does not implement */
    RLGS_ERR_CAPACITY = -5,    /* a fixed-size device table overflowed and could not be grown */
    RLGS_ERR_STATE = -6,       /* call order (e.g. run before load_trace, env_step before env_reset) */
    RLGS_ERR_SLOTS = -7,       /* more jobs ran concurrently than opts.slot_cap on-chip slots: recreate with a larger slot_cap */
    RLGS_ERR_WIRE = -8         /* a value does not fit the 16-byte wire row (rlgs_row16): rerun with RLGS_ROWFMT_WIDE */
};

/* --schedule (run_sim.py:38-49); live fifo = core/scheduling/algorithm.py:189-202,
 * sjf = run_sim.py:162-287 (dead code, restated), dlas-gpu = run_sim.py:664-947 (dead code, restated) */
enum { RLGS_SCHED_FIFO = 0, RLGS_SCHED_SJF = 1, RLGS_SCHED_DLAS_GPU = 2,
       RLGS_SCHED_DLAS = 3, /* dlas_sim_jobs(gputime=False): thresholds on attained time instead of GPU-time */
       RLGS_SCHED_SHORTEST = 4,     /* shortest_first_sim_jobs, run_sim.py:299-431: shortest remaining time first */
       RLGS_SCHED_SHORTEST_GPU = 5, /* ... shortest remaining GPU-time first */
       RLGS_SCHED_HORUS = 6,        /* schedule_horus, core/scheduling/algorithm.py:204-240: utilisation-ordered heap queue
                                       (core/jobs/base_factory.py:1-12) + look-ahead window of opts.num_buffer jobs */
       RLGS_SCHED_GANDIVA = 7,      /* algorithm.py:292-298: schedule_fifo on a plain list + gandiva_score (horus.py:6-25) +
                                       time_slice_check (algorithm.py:420-440): preempt every 100 processed ticks */
       RLGS_SCHED_HORUS_PLUS = 8    /* schedule_horus_plus, algorithm.py:242-290: opts.num_queue utilisation heaps refilled every tick by
                                       a k-means of the queued jobs (core/jobs/utils.py:36-67), next job from the queue with the most
                                       credit (job_queue_manager.py:103-127); the k-means draws are counter-based (opts.pack_seed) */ };
/* --scheme (run_sim.py:27-37); yarn = core/scheduling/algorithm.py:28-32,301-417,
 * count = resource counting only (infra/cluster.py free_gpu accounting used by run_sim.py:808-823) */
enum { RLGS_PLACE_YARN = 0, RLGS_PLACE_COUNT = 1,
       RLGS_PLACE_HORUS = 2 /* horus_placement, core/scheduling/algorithm.py:34-180 with horus_score (horus.py:28-56):
                               packs up to 4 tasks per device; what --scheme horus|horus+|gandiva select under --schedule horus|gandiva */ };
/* rows_mode: NONE = no per-tick rows; FULL = one row per tick kept in a device-resident store and
 * copied to the handle's pinned host store inside rlgs_run (overlapped with compute, one stream per
 * replica group); DEVICE = rows stay in HBM until rlgs_read_rows / rlgs_rows_view asks for them. */
enum { RLGS_ROWS_NONE = 0, RLGS_ROWS_FULL = 1, RLGS_ROWS_DEVICE = 2 };

/* rows_format: how a per-tick row is stored on the device and moved to the host (fifo tick loop; the other schedules always
 * use RLGS_ROWFMT_WIDE).  WIDE = rlgs_row, 64 bytes, self-contained.  WIRE16 = rlgs_row16, 16 bytes: the per-tick state that is
 * not an integral of the start / finish event stream; rlgs_read_rows expands it to rlgs_row on the host (two prefix sums over
 * the per-job tables the run produced, see rlgs_row16).  WIRE12 = rlgs_row12, 12 bytes: the counts leave too (they are cumulative
 * counts of the same tables).  With opts.fetch_jobs = 2 a 60k-job run crosses PCIe as 1.24 MB instead of 4.76 MB.
 * EVENT16 = rlgs_row16e, 16 bytes: the 12-byte row plus the tick's start event (at most one job starts per tick,
 * schedule.py:188-190).  The row stream then IS the event log: without network costs a job ends dur_ticks after it started and
 * jobs finish in (end tick, start tick) order, so rlgs_read_jobs / rlgs_read_rows rebuild the per-job tables from the rows of the
 * replica when they were not copied (opts.fetch_jobs = 0), and a run crosses PCIe as 16 bytes per tick, all of it while the
 * simulation is still running.
 * EVENT4 = rlgs_row4e, 4 bytes: the event log alone — idle_nodes, "the queue head started at this tick" and a check field.
 * Under fifo the job that starts is always the queue head and the queue is a function of the trace and of the earlier start
 * events, so a replay of the queue on the host (one push per arrival, one pop per start event, O(ticks + jobs)) names every
 * started job and yields the pending-time statistics of every tick (they are arrival ticks at fixed positions of the queue);
 * the tables and the remaining row fields then follow as for EVENT16.  Fifo without network costs only. */
enum { RLGS_ROWFMT_WIDE = 0, RLGS_ROWFMT_WIRE16 = 1, RLGS_ROWFMT_WIRE12 = 2, RLGS_ROWFMT_EVENT16 = 3, RLGS_ROWFMT_EVENT4 = 4 };

/* Cluster spec: flags --num_switch .. --mem_p_node (run_sim.py:50-82) or cluster_spec.csv
 * (infra/infrastructure.py:78-105).  Replaces Infrastructure._init_nodes (infrastructure.py:45-69). */
typedef struct {
    int32_t num_switch;
    int32_t num_node_p_switch;
    int32_t num_gpu_p_node; /* 1..32 */
    int32_t num_cpu_p_node;
    int32_t mem_p_node;
    int32_t reserved;
} rlgs_cluster_spec;

#define RLGS_MAX_QUEUES 8

typedef struct {
    int32_t device;        /* CUDA ordinal */
    int32_t n_replicas;    /* >= 1 */
    int32_t schedule;      /* RLGS_SCHED_* */
    int32_t placement;     /* RLGS_PLACE_* */
    int32_t rows_mode;     /* RLGS_ROWS_*: keep one cluster.csv row per tick / event */
    int32_t slot_cap;      /* running-job slots per replica held on chip; 0 = auto */
    int32_t n_streams;     /* replica groups, each on its own CUDA stream (kernel + result copies); 0 = auto */
    int32_t ticks_per_launch; /* 0 = run to completion in one launch; >0 = bounded launches (state is saved
                                 to / restored from HBM between launches) */
    int32_t num_queue;     /* dlas-gpu: number of MLFQ queues (README.md:57-62), 1..RLGS_MAX_QUEUES; horus+: number of job queues */
    int32_t enable_network_costs; /* --enable_network_costs (run_sim.py:54); network_service.py:3-39 */
    int32_t fetch_jobs;    /* 1 = copy the per-job tables (start, end, finish_order) to the host inside rlgs_run as well; 2 = end and
                              finish_order only: under fifo without network costs a finished job started at end - dur_ticks, which is
                              what rlgs_read_jobs / the row expansion then use (the start plane stays on the device, fetched if ever needed) */
    int32_t num_buffer;    /* horus: look-ahead window, --num_buffer (run_sim.py:76); 0 = the reference default 5 */
    int32_t queue_limit[RLGS_MAX_QUEUES]; /* dlas-gpu demotion thresholds in GPU-ticks */
    double bandwidth;          /* --bandwidth MB/s (run_sim.py:59) */
    double internode_latency;  /* --internode_latency s (run_sim.py:65) */
    int64_t max_ticks;         /* safety stop; 0 = none */
    int64_t rows_cap;          /* initial capacity (ticks) of the row store; 0 = auto; grown when a replica fills it */
    int32_t pack_rng;          /* horus: 0 = every utilisation draw of infra/device.py:52 returns its mean (the reference on
                                  traces with gpu_utilization_max == gpu_utilization_avg); 1 = build-defined counter-based draw */
    uint32_t pack_seed;        /* seed of that draw */
    int32_t rows_format;       /* RLGS_ROWFMT_* (fifo tick loop only) */
    int32_t lanes_per_replica; /* fifo tick loop: lanes of a warp that advance one replica: 8, 16 or 32 (a warp carries 4, 2 or 1
                                  replicas); 0 = chosen from n_replicas so that the GPU is filled */
} rlgs_opts;

/*
 * One job of a trace, already in queue-arrival order (the row order of JobTraceReader.prepare_jobs,
 * core/jobs/job_generator.py:181-196).  32-byte records, the unit the device streams from HBM.
 * Derived on the host exactly as the reference's Job constructor derives them (core/jobs/job.py:75-110,
 * jobs_manager.py:233-238) — see rlgpuschedule_b200/ingest.py.
 */
typedef struct {
    int32_t arrival_tick;  /* first integer tick d with normalized_time <= d (job_generator.py:203) */
    int32_t dur_ticks;     /* max(1, ceil(duration)): ticks from start to finish (jobs_manager.py:243-250) */
    uint16_t gpus;         /* ceil(used_gpus): the value `free_devices >= job.gpus` compares against */
    uint16_t tasks;        /* int(used_gpus // gpu_per_container) (job.py:96-99) */
    uint16_t gpus_per_task;/* gpu_per_container */
    uint16_t least_nodes_fits; /* bit15 = an empty device accepts the task (infra/device.py:67-77);
                                  bits0-14 = ceil(used_gpus / num_gpu_p_node) (algorithm.py:310) */
    int64_t mem_term;      /* tasks * gpus_per_task * min(cap, memory_max MiB): what the job adds to the
                              avg_gpu_memory_allocated numerator, in units of 2^-mem_shift MiB */
    uint16_t util_mu_q;    /* gpu_utilization_avg * 512 (statistics of the unseeded RNG column only) */
    uint16_t util_sd_q;    /* (gpu_utilization_max - gpu_utilization_avg)/2 * 512 */
    int32_t index;         /* position of this job in the trace (0..n-1) */
} rlgs_job;

/* Optional per-job inputs of the network-cost model (core/network/network_service.py:34-37). */
typedef struct {
    const double *duration;   /* minutes * scale_factor, float64 */
    const double *model_mb;   /* model size in MB (model/model_factory.py:19-55), 0 if unknown */
    const double *iterations; /* training iterations, 0 if unknown */
} rlgs_netcost_inputs;

/* One cluster.csv row as integer sufficient statistics (core/scheduling/schedule.py:95-133,
 * log_manager.py:118-135).  The float columns are finished on the host (rlgpuschedule_b200/log_manager.py).
 * fifo: one row per tick, row i has delta = i + 1.
 * sjf / dlas-gpu: one row per event with the legacy cluster.csv columns (log.py:137-258):
 *   idle_nodes = idle_node, busy_gpus = busy_gpu, running / queued / finished = running_job / pending_job /
 *   completed_job, median_lo = event time, median_hi = full_node; the remaining fields are 0. */
typedef struct {
    int32_t idle_nodes;   /* num_idle_nodes; num_busy_nodes = N - idle_nodes */
    int32_t busy_gpus;    /* num_busy_gpus; num_idle_gpus = D - busy_gpus */
    int32_t running, queued, finished;
    int32_t median_lo, median_hi; /* pending time of the two middle queued jobs (np.median = their mean) */
    int32_t max_pending;
    int64_t sum_pending;  /* avg_pending_time = sum_pending / (queued + 1e-9) */
    int64_t mem_sum;      /* sum over busy devices of mem_term: avg_gpu_memory_allocated numerator */
    int64_t util_mu_sum;  /* sum over busy devices of util_mu_q */
    int64_t util_var_sum; /* sum over busy devices of util_sd_q^2 */
} rlgs_row;

/* 16-byte wire row of the fifo tick loop (RLGS_ROWFMT_WIRE16), row i has delta = i + 1.  Bit fields, least significant first:
 *   w[0]: idle_nodes:12 | finished:20          w[1]: queued:20 | max_pending[11:0]:12
 *   w[2]: max_pending[23:12]:12 | median_lo[19:0]:20     w[3]: median_lo[23:20]:4 | median_hi:24 | 0:4
 * Limits: nodes <= 4095, jobs < 2^20, ticks < 2^24 (else rlgs_run returns RLGS_ERR_WIRE / rlgs_create RLGS_ERR_UNSUPPORTED).
 * The remaining rlgs_row fields are sums of per-job constants over the running (or queued) set, i.e. integrals of the
 * event stream the run also returns (start_tick, end_tick, finish_order):
 *   arrived(i) = #jobs with arrival_tick <= i;  running = arrived - queued - finished;  started = running + finished
 *   X(i) = sum of x over the first started(i) jobs in start order - sum of x over the first finished(i) jobs in finish order
 *          for x = devices, mem_term, util_mu_q * devices, util_sd_q^2 * devices      -> busy_gpus, mem_sum, util_mu_sum, util_var_sum
 *   sum_pending = queued * (i + 1) - (sum of arrival_tick over the first arrived(i) jobs - same over the first started(i) in start order)
 * rlgs_read_rows does this expansion (int64 prefix sums, exact). */
typedef struct { uint32_t w[4]; } rlgs_row16;

/* 12-byte wire row (RLGS_ROWFMT_WIRE12): the per-tick quantities that are not functions of the event tables.
 *   w[0]: max_pending:24 | idle_nodes[7:0]:8      w[1]: median_lo:24 | idle_nodes[11:8]:4 | 0:4      w[2]: median_hi:24 | 0:8
 * finished(i) = #jobs with end_tick <= i + 1, started(i) = #jobs with start_tick <= i, queued = arrived(i) - started(i); the
 * rest as for rlgs_row16.  Same limits. */
typedef struct { uint32_t w[3]; } rlgs_row12;

/* 16-byte event row (RLGS_ROWFMT_EVENT16): w[0..2] as rlgs_row12, w[3] = 1 + trace index of the job that started at this
 * tick (row i: start_tick = i), 0 = none.  Derived on the host when the tables were not copied (fifo without network costs):
 * end_tick = start_tick + dur_ticks, finish_order = the started jobs sorted by (end_tick, start_tick) — the order in which
 * release_finished_jobs walks running_jobs (schedule.py:141-162).  The device keeps its own tables (rlgs_opts.fetch_jobs = 1 copies
 * them; tests compare both). */
typedef struct { uint32_t w[4]; } rlgs_row16e;

/* 4-byte event row (RLGS_ROWFMT_EVENT4), row i has delta = i + 1:
 *   w: idle_nodes:12 | started:1 | queued[18:0]:19
 * started = 1 when the scheduling attempt of tick i started the queue head (schedule.py:188-190).  The host replays the queue
 * exactly as the tick loop keeps it (jobs_manager.py:228-241: the jobs arriving at a tick go to the FRONT of the queue in trace
 * order; the attempt takes the front): the front at a row with started = 1 is the job whose start_tick is i.  After the tick's
 * pushes and pop the queue holds Q jobs, front first; with delta = i + 1
 *   max_pending = delta - arrival tick of the back of the queue (the tick at which the queue last became non-empty)
 *   median_lo / median_hi = delta - arrival tick of the jobs at positions (Q - 1) / 2 and Q / 2 from the front
 * (the queue is sorted by arrival tick, newest first, so these are np.median's two middle elements, jobs_manager.py:87).
 * queued[18:0] must equal the low bits of the replayed Q at every row (else RLGS_ERR_STATE).  Tables and sums as for rlgs_row16e. */
typedef struct { uint32_t w; } rlgs_row4e;

typedef struct {
    int64_t n_ticks;       /* rows produced (fifo: ticks; sjf/dlas: events) */
    int64_t makespan;      /* last simulated time */
    int64_t sum_jct;       /* sum over finished jobs of end - submit-arrival tick */
    int64_t sum_queued;    /* sum over ticks of queued jobs  (job-updates accounting) */
    int64_t sum_running;   /* sum over ticks of running jobs */
    int64_t events;        /* arrivals + starts + finishes + preemptions + queue jumps */
    int32_t n_jobs, n_arrived, n_started, n_finished;
    int32_t max_queued, max_running;
    int32_t status;        /* RLGS_OK or RLGS_ERR_* raised on the device for this replica */
    int32_t done;
} rlgs_summary;

typedef struct rlgs_sim rlgs_sim;

int32_t rlgs_version(void);
const char *rlgs_last_error(void);

/* Replaces Infrastructure(FLAGS) + JobQueueManager/JobsManager/Scheduler construction (run_sim.py:1716-1735). */
int32_t rlgs_create(const rlgs_cluster_spec *spec, const rlgs_opts *opts, rlgs_sim **out);
void rlgs_destroy(rlgs_sim *sim);

/* Replaces JobTraceReader ingestion into JobsManager (jobs_manager.py:16-18,228-241): copies `n` job
 * records to the device once and attaches them to replicas [first_replica, first_replica+n_replicas).
 * `net` may be NULL unless opts.enable_network_costs is set. */
int32_t rlgs_load_trace(rlgs_sim *sim, int32_t first_replica, int32_t n_replicas, const rlgs_job *jobs, int32_t n,
                        const rlgs_netcost_inputs *net);

/* Per-job inputs of the pack placement (RLGS_PLACE_HORUS), arrays of n entries in trace order; attach them with
 * rlgs_load_pack_inputs after rlgs_load_trace of the same replica range.  Memory amounts are integers in units of
 * 2^-mem_shift MiB (the shift rlgs_job.mem_term uses). */
typedef struct {
    const double *util_avg;     /* gpu_utilization_avg (Task.gpu_utilization_avg, core/jobs/job.py:30) */
    const double *util_sd;      /* (gpu_utilization_max - gpu_utilization_avg) / 2 (infra/device.py:52) */
    const int64_t *task_mem;    /* memory_max of one task, not clamped (Task.gpu_memory_max) */
    const int32_t *heap_cap;    /* floor(used_gpus): size of horus_placement's node heap (algorithm.py:64) */
    int32_t mem_shift;
    int32_t gpu_mem_cap_mib;    /* --gpu_memory_capacity * 1024 (infra/infrastructure.py:36) */
    /* RLGS_SCHED_HORUS_PLUS only (k-means features, core/jobs/utils.py:4-22); NULL otherwise */
    const double *util_max;     /* gpu_utilization_max */
    const double *mem_avg_mib;  /* memory_avg in MiB */
    const double *used_gpus;    /* used_gpus (float) */
} rlgs_pack_inputs;
int32_t rlgs_load_pack_inputs(rlgs_sim *sim, int32_t first_replica, int32_t n_replicas, const rlgs_pack_inputs *in, int32_t n);

/* Replaces Scheduler.start() (core/scheduling/schedule.py:178-216): runs every replica to completion.
 * Blocking.  Resets replica state first, so it can be called repeatedly (bench steps). */
int32_t rlgs_run(rlgs_sim *sim);
/* Device time of the kernels of the last rlgs_run, from CUDA events on the launch stream. */
int32_t rlgs_last_run_ms(rlgs_sim *sim, float *kernel_ms, int32_t *n_launches);
/* enable != 0: launch on the caller's cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream; NULL is the
 * legacy default stream); enable == 0: back to the handle's own stream. */
int32_t rlgs_set_stream(rlgs_sim *sim, void *cuda_stream, int32_t enable);

int32_t rlgs_get_summary(rlgs_sim *sim, int32_t replica, rlgs_summary *out);
/* Replaces LogManager.jcts' walk over finished_jobs (log_manager.py:137-155): per job (trace order)
 * start/end tick (-1 = never) and preempt count; finish_order[k] = trace index of the k-th finished job.
 * Any pointer may be NULL.  Arrays must hold n_jobs entries. */
int32_t rlgs_read_jobs(rlgs_sim *sim, int32_t replica, int32_t *finish_order, int32_t *start_tick, int32_t *end_tick,
                       int32_t *preempt, int32_t *first_node);
/* Replaces the per-tick LogManager.step_cluster rows (log_manager.py:118-135): copies rows
 * [first, first+count) of `replica` (rows_mode FULL). */
int32_t rlgs_read_rows(rlgs_sim *sim, int32_t replica, int64_t first, int64_t count, rlgs_row *out);
/* The same rows in the handle's wire format (RLGS_ROWFMT_WIRE16 -> rlgs_row16, RLGS_ROWFMT_WIRE12 -> rlgs_row12), without the expansion. */
int32_t rlgs_read_rows16(rlgs_sim *sim, int32_t replica, int64_t first, int64_t count, rlgs_row16 *out);
int32_t rlgs_read_rows12(rlgs_sim *sim, int32_t replica, int64_t first, int64_t count, rlgs_row12 *out);
int32_t rlgs_read_rows16e(rlgs_sim *sim, int32_t replica, int64_t first, int64_t count, rlgs_row16e *out);
int32_t rlgs_read_rows4e(rlgs_sim *sim, int32_t replica, int64_t first, int64_t count, rlgs_row4e *out);
/* Zero-copy variant: rows [chunk*RLGS_ROWS_PER_CHUNK, ...) of `replica` inside the handle's pinned host
 * mirror (the store is chunk-major so that a whole chunk of every replica moves in one contiguous copy);
 * valid until the next rlgs_run / destroy. */
#define RLGS_ROWS_PER_CHUNK 4096
int32_t rlgs_rows_view(rlgs_sim *sim, int32_t replica, int32_t chunk, const rlgs_row **rows, int64_t *count);   /* RLGS_ROWFMT_WIDE */
int32_t rlgs_rows16_view(rlgs_sim *sim, int32_t replica, int32_t chunk, const rlgs_row16 **rows, int64_t *count); /* RLGS_ROWFMT_WIRE16 */
int32_t rlgs_rows12_view(rlgs_sim *sim, int32_t replica, int32_t chunk, const rlgs_row12 **rows, int64_t *count); /* RLGS_ROWFMT_WIRE12 */
int32_t rlgs_rows16e_view(rlgs_sim *sim, int32_t replica, int32_t chunk, const rlgs_row16e **rows, int64_t *count); /* RLGS_ROWFMT_EVENT16 */
int32_t rlgs_rows4e_view(rlgs_sim *sim, int32_t replica, int32_t chunk, const rlgs_row4e **rows, int64_t *count);   /* RLGS_ROWFMT_EVENT4 */
/* Host-only companion of RLGS_ROWFMT_EVENT4 (no handle, no CUDA call): expands the event log of ONE replica that the caller kept
 * itself (rlgs_rows4e_view / rlgs_read_rows4e), by the replay rule documented at rlgs_row4e.  jobs[n_jobs] = the trace the
 * replica ran, rows[n_rows] = its rows.  Outputs, each may be NULL: start_tick / end_tick [n_jobs] (-1 = never), finish_order
 * [n_jobs] (first *n_finished entries, rest -1), pending[3 * n_rows] = max_pending, median_lo, median_hi of every row.
 * RLGS_ERR_STATE when the rows do not replay against the trace (check field, start from an empty queue). */
int32_t rlgs_replay_rows4e(const rlgs_job *jobs, int32_t n_jobs, const rlgs_row4e *rows, int64_t n_rows, int32_t *start_tick,
                           int32_t *end_tick, int32_t *finish_order, int32_t *n_finished, int32_t *pending);
/* Per-job int32 column `plane` (trace order): what LOG.job_complete logs for the preemptive schedules
 * (log.py:316-330). */
enum { RLGS_PLANE_START = 0, RLGS_PLANE_END = 1, RLGS_PLANE_FINISH_ORDER = 2,
       RLGS_PLANE_AUX = 3,      /* fifo: first placement-log entry; sjf/dlas-gpu: pending_time;
                                   horus: 1 = Job.get_duration() is original + 5 (a task was de-interfered), 0 = original */
       RLGS_PLANE_PREEMPT = 4,  /* horus / gandiva: Job.time_processed() of a finished job (the jct column) */
       RLGS_PLANE_RESUME = 5 }; /* horus / gandiva: Job.migration_count of a finished job (the preempt column) */
int32_t rlgs_read_job_plane(rlgs_sim *sim, int32_t replica, int32_t plane, int32_t *out);
/* Episode return per replica: -(sum of job completion times), the reward of the vectorised Environment. */
int32_t rlgs_returns(rlgs_sim *sim, int64_t *out_n_replicas);
/* Device pointer to the same int64[n_replicas] buffer (the NCCL all-gather send buffer). */
int32_t rlgs_returns_device_ptr(rlgs_sim *sim, void **dev_ptr);

/* Per-job duration after network costs (Job.add_network_costs, job.py:196-197); enable_network_costs only. */
int32_t rlgs_read_durations(rlgs_sim *sim, int32_t replica, double *out);

/*
 * Vectorised RL environment: replaces the reference's stub model/env.py:1-6 (Environment.step(action): pass);
 * semantics are build-defined (DESIGN.md "Environment").  One step = one scheduler tick of every replica of a
 * fifo/yarn handle; the action picks which of the first `window_k` queued jobs gets the tick's placement
 * attempt (-1 = none).  policy 0 = queue head, 1 = random window (counter-based RNG keyed by seed, replica,
 * tick; n_ticks may be > 1 to roll whole episodes on the device), 2 = actions[] (n_ticks = 1).
 * obs [n_replicas][obs_dim = 3 N + 5 window_k + 4]: free GPUs / free cpu / free mem per node, (gpus, tasks, dur_ticks, pending,
 * trace index) of the window jobs (index -1 = empty place), then queued, running, finished, tick.  A handle created with
 * RLGS_ROWS_DEVICE + RLGS_ROWFMT_WIDE + lanes_per_replica 32 also records one rlgs_row per stepped tick (rlgs_read_rows after
 * rlgs_env_sync): what the host-callable scheduling plugins use.  reward = -(queued + running) per tick.  All pointers
 * are DEVICE pointers; calls are asynchronous on the handle's stream until rlgs_env_sync.  rlgs_env_step before rlgs_env_reset
 * (or after a new rlgs_load_trace) returns RLGS_ERR_STATE.
 */
int32_t rlgs_env_obs_dim(rlgs_sim *sim, int32_t window_k, int32_t *dim);
int32_t rlgs_env_reset(rlgs_sim *sim);
int32_t rlgs_env_step(rlgs_sim *sim, const int32_t *actions, float *obs, float *reward, uint8_t *done,
                      int32_t policy, int32_t window_k, uint32_t seed, int32_t n_ticks);
/* Writes the observation of the current state without advancing it (reward 0): the observation reset() hands out. */
int32_t rlgs_env_observe(rlgs_sim *sim, float *obs, float *reward, uint8_t *done, int32_t window_k);
int32_t rlgs_env_sync(rlgs_sim *sim);

#ifdef __cplusplus
}
#endif
#endif /* RLGS_H */
