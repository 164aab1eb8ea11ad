// pack_horus.cuh — the schedules horus, horus+ and gandiva of the live reference on the device, one warp per replica, over the
// packing placement (horus_placement) or the yarn fit.  pack_horus_kernel<GANDIVA, YARN, PLUS> is instantiated six times.
//
// Restates (reference paths):
//   Scheduler.start tick loop            core/scheduling/schedule.py:178-216
//   schedule_horus (look-ahead window)   core/scheduling/algorithm.py:204-240
//   horus_placement                      core/scheduling/algorithm.py:34-180   (heap of scored nodes, trial placement per
//                                        heap entry with rack walk + backtracking, fewest-nodes plan wins)
//   horus_score                          core/scheduling/horus.py:28-56        (float64, np.polyval Horner form)
//   Node / Device pack=True branches     infra/node.py:146-221, infra/device.py:19-77 (<= 4 tasks per device, 500 MiB margin,
//                                        cpu/mem and accepting devices stay charged when a task finds too few devices)
//   utilisation-ordered job queue        core/jobs/base_factory.py:1-12 + job_queue_manager.py:129-154 (stdlib heapq)
//   interference bookkeeping             infra/node.py:71-91, core/jobs/jobs_manager.py:175-187 (+5 ticks once a flagged
//                                        task is left alone on a device)
//   gandiva                              schedule_fifo algorithm.py:189-202 on a plain list, gandiva_score horus.py:6-25,
//                                        time_slice_check algorithm.py:420-444 -> JobsManager.preempt jobs_manager.py:150-173
//   horus+                               schedule_horus_plus algorithm.py:242-290, clusterize core/jobs/utils.py:36-67,
//                                        credits job_queue_manager.py:103-127, re-clustering insert jobs_manager.py:115-140
//   --scheme yarn under these schedules  ms_yarn_placement algorithm.py:28-32 (yarn_place.cuh)
//
// Everything order-dependent in the reference (heapq sift order, stable sorts, dict insertion order, the order of
// release calls) is kept, because ties are the common case (all idle nodes score the same).  Scalar phases (heap
// sifts, calendar walks) run on lane 0 between __syncwarp()s; device scans, node scoring, the rack pre-filter and the
// per-node capacity are lane-parallel.  State lives in global memory (L1/L2 resident per replica) so that a bounded
// launch can stop and resume at any tick; node scores and the node heap are shared-memory scratch.
//
// Memory amounts are integers in units of 2^-shift MiB (exact: ingest picks the shift), scores are float64 with the
// reference's operation order (the library is built with --fmad=false).
#pragma once
#include "yarn_place.cuh"

#define PACK_DEV_SLOTS 4      // infra/device.py:72
#define PACK_CAL_W 128
#define PACK_MAX_TASKS 32
#define PACK_MAX_HEAP 255

struct PackJob {              // per job of a trace, shared by the replicas that replay it
    double util_avg;          // gpu_utilization_avg
    double util_sd;           // (gpu_utilization_max - gpu_utilization_avg) / 2   (device.py:52)
    int64_t mem;              // memory_max of one task, units of 2^-shift MiB
    int32_t heap_cap;         // floor(used_gpus): `len(nodes_stack) > gpu_demand` (algorithm.py:64)
    int32_t task_off;         // first entry of the job in tnode[]
};

#define PACK_MAX_Q 8
struct PlusFeat {             // horus+ k-means features of a job (core/jobs/utils.py:4-22), float64 like the reference
    double f[7];              // len(tasks), gpu_utilization_avg, gpu_per_worker, gpus, gpu_utilization_max, gpu_mem_avg, gpu_mem_max (MiB)
    double tdist;             // transform_to_dist: their left-to-right sum
};

struct PackDesc {
    const rlgs_job *trace;
    const PackJob *pj;
    const PlusFeat *feat;     // horus+ only
    int32_t *planes[6];       // start (of the last run), end, finish_order, 1 = get_duration() is duration + 5, ticks processed (jct), number of starts
    int32_t *units;           // [N] tasks charged to the node (cpu = 12u, mem = 60u)
    int32_t *ntk;             // [N] len(placed_tasks) << 16 | len(running_tasks)
    int32_t *npj;             // [N] len(placed_jobs)
    int32_t *dn;              // [D] len(device.running_tasks)
    int64_t *dm;              // [D] sum of min(cap, memory_max) of the tasks on the device
    int2 *ent;                // [D][4] (job, task) in insertion order
    uint32_t *pjbits;         // [J][W] node.placed_jobs membership of each job
    double *qkey; int32_t *qjob;      // [J] the job queue: a heapq array ordered by utilisation
    int32_t *lprev, *lnext;   // [J] queued jobs in arrival order (pending-time statistics)
    int32_t *pend;            // [J] tick at which the running job finishes
    int32_t *cnext;           // [J] calendar chain
    int32_t *chead;           // [PACK_CAL_W]
    int16_t *tnode;           // [sum of tasks] node of each placed task (tasks_running_on)
    int32_t *fin;             // [J] jobs finishing (or losing their time slice) at this tick
    uint32_t *imask;          // [J] Task.interfered, bit per task
    uint32_t *bmask;          // [J] Task.duration == original + 5, bit per task
    int32_t *jflag;           // [J] bit0: the job has leaked device entries
    int32_t *pproc;           // [J] ticks processed in earlier runs (gandiva time slices)
    int32_t *nstart;          // [J] Job.migration_count: number of starts
    int32_t *qtick;           // [J] gandiva: tick the job entered the queue (arrival or preemption): pending_time = d - qtick
    int32_t *cbk;             // [J] calendar bucket the running job is chained in
    int32_t *snext, *shead, *sat;     // gandiva: calendar of time-slice ticks ([J], [PACK_CAL_W], [J])
    // --scheme yarn under these schedules: the node view of yarn_place.cuh and the placement of every running job
    uint32_t *ybusy, *ykey, *yever;   // [N], [N], [ceil(N/32)]
    int2 *plog;               // [sum of tasks] (node | tasks << 16, device mask) entries of job j at plog[task_off ..]
    int32_t *pcnt;            // [J] entries of the job's placement
    // horus+: queue q is the heap (qkey, qjob) + q * J; the k-means works on kjobs / kassign / kold / ktmp
    int32_t *kjobs, *kassign, *kold;  // [J]
    double *ktmp;             // [J]
    int64_t cap_units, margin_units;  // gpu memory capacity and the 500 MiB margin in units
    double cap_mib, unit_mib; // capacity in MiB, 2^-shift
    int32_t J, W;
};

struct PackState {
    int32_t d, cursor, F, Q, R;
    int32_t n_free_nodes, idle_nodes, busy_gpus, start_seq;
    int32_t lhead, ltail, mlo, mrank;
    int32_t done, status, max_q, max_r, r_pre, preempts;   // r_pre: running jobs before the post-tick plugin (schedule.py:195)
    int32_t pqn[PACK_MAX_Q];  // horus+: length of each queue
    uint32_t kcalls, pad1;    // horus+: np.random.randint / choice calls so far
    int64_t mem_sum, util_mu_sum, util_var_sum, sum_arr, sum_jct, sumQ, sumR, events;
#ifdef PACK_PROFILE
    int64_t prof[16];   // cycles: 0 arrivals, 1 queue pops, 2 score, 3 heap, 4 sort, 5 trials, 6 real place, 7 re-push, 8 start, 9 finish, 10 row, 11 attempts
#endif
};
#ifdef PACK_PROFILE
#define PACK_T0 long long _t0 = clock64()
#define PACK_T(i) do { long long _t1 = clock64(); st.prof[i] += _t1 - _t0; _t0 = _t1; } while (0)
#else
#define PACK_T0
#define PACK_T(i)
#endif

struct PackParams {
    int32_t num_buffer;       // --num_buffer (run_sim.py:76): look-ahead window
    int32_t rng_on;           // 0: every utilisation draw returns its mean
    uint32_t seed;
    int32_t nodes_per_rack, racks;
    int32_t tick_budget;
    int32_t gandiva;          // 1: --schedule gandiva = fifo queue + gandiva_score + time slicing (algorithm.py:292-298,420-444)
    int32_t plus_k;           // horus+: number of queues (--num_queue), 0 otherwise
    uint32_t plus_seed;       // horus+: seed of the k-means draws
    int64_t max_ticks;
};
#define PACK_QUANTA 100       // time_slice_check, algorithm.py:427

// ---- build-defined stand-in for np.random.normal (the reference's RNG is unseeded): Irwin-Hall sum of twelve
// 16-bit uniforms keyed by (seed, replica, tick, look-ahead position, task pass, device, slot)
__device__ __forceinline__ uint64_t pack_mix64(uint64_t z) { z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
__device__ __forceinline__ double pack_draw(uint32_t seed, uint32_t replica, uint32_t tick, uint32_t attempt, uint32_t pass, uint32_t dev, uint32_t slot) {
    uint64_t h = pack_mix64((((uint64_t)seed << 32) | replica) + 0x9E3779B97F4A7C15ull);
    h = pack_mix64(h ^ (((uint64_t)tick << 32) | ((uint64_t)attempt << 16) | pass));
    h = pack_mix64(h ^ (((uint64_t)dev << 8) | slot));
    uint32_t sum = 0;
#pragma unroll
    for (int k = 1; k <= 3; ++k) { uint64_t w = pack_mix64(h + (uint64_t)k * 0x9E3779B97F4A7C15ull); sum += (uint32_t)(w & 0xffff) + (uint32_t)((w >> 16) & 0xffff) + (uint32_t)((w >> 32) & 0xffff) + (uint32_t)(w >> 48); }
    return ((double)sum - 393210.0) / 65536.0;
}

// CompareAbleByUtilization.__lt__ (base_factory.py:8-12)
__device__ __forceinline__ bool pack_lt_util(double a, double b) { return a != 0.0 ? a < b : false; }

// heapq.heappush / heappop on (qkey, qjob): lane 0 walks the heap, the others wait at the __syncwarp
__device__ __forceinline__ void pack_q_siftdown(const PackDesc &D, int pos, double key, int job) {
    while (pos > 0) {
        int pp = (pos - 1) >> 1; double pk = D.qkey[pp];
        if (!pack_lt_util(key, pk)) break;
        D.qkey[pos] = pk; D.qjob[pos] = D.qjob[pp]; pos = pp;
    }
    D.qkey[pos] = key; D.qjob[pos] = job;
}
// the same two operations on queue `qi` of horus+ (arrays offset by qi * J)
__device__ __forceinline__ void pack_qk_push(const PackDesc &D, int lane, int qi, int &n, double key, int job) {
    PackDesc Q = D; Q.qkey = D.qkey + (size_t)qi * D.J; Q.qjob = D.qjob + (size_t)qi * D.J;
    if (lane == 0) pack_q_siftdown(Q, n, key, job);
    n += 1;
    __syncwarp();
}
__device__ __forceinline__ void pack_q_pop(const PackDesc &D, int lane, int &q, double &key, int &job);
__device__ __forceinline__ void pack_qk_pop(const PackDesc &D, int lane, int qi, int &n, double &key, int &job) {
    PackDesc Q = D; Q.qkey = D.qkey + (size_t)qi * D.J; Q.qjob = D.qjob + (size_t)qi * D.J;
    pack_q_pop(Q, lane, n, key, job);
}
__device__ __forceinline__ void pack_q_push(const PackDesc &D, int lane, int &q, double key, int job) {
    if (lane == 0) pack_q_siftdown(D, q, key, job);
    q += 1;
    __syncwarp();
}
__device__ __forceinline__ void pack_q_pop(const PackDesc &D, int lane, int &q, double &key, int &job) {
    q -= 1;
    if (lane == 0) {
        double lk = D.qkey[q]; int lj = D.qjob[q];
        if (q == 0) { key = lk; job = lj; }
        else {
            key = D.qkey[0]; job = D.qjob[0];
            int pos = 0, child = 1;
            while (child < q) {                           // _siftup: bubble the smaller child up to a leaf ...
                int right = child + 1;
                double ck = D.qkey[child];
                if (right < q) { double rk = D.qkey[right]; if (!pack_lt_util(ck, rk)) { child = right; ck = rk; } }
                D.qkey[pos] = ck; D.qjob[pos] = D.qjob[child];
                pos = child; child = 2 * pos + 1;
            }
            pack_q_siftdown(D, pos, lk, lj);              // ... then sift the displaced last item down from there
        }
    }
    __syncwarp();
    key = __shfl_sync(RLGS_FULL, key, 0); job = __shfl_sync(RLGS_FULL, job, 0);
}

// nodes_stack of horus_placement: NodeDeviceInfo.__lt__ is `self.min_score > other.min_score` (algorithm.py:25-26)
__device__ __forceinline__ void pack_h_siftdown(double *hscore, int32_t *hnode, int pos, double sc, int node) {
    while (pos > 0) {
        int pp = (pos - 1) >> 1; double ps = hscore[pp];
        if (!(sc > ps)) break;
        hscore[pos] = ps; hnode[pos] = hnode[pp]; pos = pp;
    }
    hscore[pos] = sc; hnode[pos] = node;
}
__device__ __forceinline__ void pack_h_pop(double *hscore, int32_t *hnode, int len) {   // len = length after the pop
    double ls = hscore[len]; int ln = hnode[len];
    if (len == 0) return;
    int pos = 0, child = 1;
    while (child < len) {
        int right = child + 1;
        double cs = hscore[child];
        if (right < len) { double rs = hscore[right]; if (!(cs > rs)) { child = right; cs = rs; } }
        hscore[pos] = cs; hnode[pos] = hnode[child];
        pos = child; child = 2 * pos + 1;
    }
    pack_h_siftdown(hscore, hnode, pos, ls, ln);
}

__device__ __forceinline__ bool pack_dev_fits(const PackDesc &D, int dev, int64_t m) {   // Device.can_fit (device.py:67-77)
    int n = D.dn[dev];
    int64_t cur = min(D.dm[dev], D.cap_units);
    return n < PACK_DEV_SLOTS && D.cap_units - (cur + m) > D.margin_units;
}

struct PackCtx {              // registers shared by the placement helpers
    int lane;
    int job, T, gpc;
    int64_t m;                // memory_max of one task of the job being placed
    int mu_q, sd_q;           // quantised utilisation statistics of the job (cluster.csv's RNG column)
    uint32_t interf;          // Task.interfered of the job's tasks, bit per task
    uint32_t bump;            // Task.duration == original + 5, bit per task (reset by add_task on a device with < 2 tasks)
    double *score;            // shared memory [N]: min_cost of each node, < 0 = node cannot take the task
    double *hscore; int32_t *hnode;   // shared memory [PACK_MAX_HEAP + 1]: horus_placement's nodes_stack
};
__host__ __device__ inline size_t pack_smem_bytes(int n_nodes) { return 8 * (size_t)((n_nodes + 1) & ~1) + 12 * (size_t)(PACK_MAX_HEAP + 1) + 8 * 8 * 8; }   // + horus+ centroid features [8][8]

__device__ __forceinline__ void pack_idle_delta(PackState &st, bool was_idle, bool now_idle) { st.idle_nodes += (int)now_idle - (int)was_idle; }
__device__ __forceinline__ bool pack_node_idle(const PackDesc &D, int i) { return D.ntk[i] == 0 && D.npj[i] == 0; }   // Node.is_idle (node.py:93-97)

__device__ __forceinline__ void pack_pj_set(const PackDesc &D, PackState &st, int node, int job) {
    uint32_t *w = &D.pjbits[(size_t)job * D.W + (node >> 5)]; const uint32_t b = 1u << (node & 31), v = *w;
    if (!(v & b)) { const bool was = pack_node_idle(D, node); const int np = D.npj[node]; __syncwarp(); *w = v | b; D.npj[node] = np + 1; pack_idle_delta(st, was, false); }
    __syncwarp();
}
__device__ __forceinline__ void pack_pj_pop(const PackDesc &D, PackState &st, int node, int job) {
    uint32_t *w = &D.pjbits[(size_t)job * D.W + (node >> 5)]; const uint32_t b = 1u << (node & 31), v = *w;
    if (v & b) { const int np = D.npj[node] - 1; const int tk = D.ntk[node]; __syncwarp(); *w = v & ~b; D.npj[node] = np; pack_idle_delta(st, false, np == 0 && tk == 0); }
    __syncwarp();
}

// Node.try_reserve_and_placed_task(pack=True) (node.py:200-221).  0 = node cannot take the task (no side effect),
// 1 = placed, 2 = too few devices accepted it: cpu/mem and those devices stay charged.
__device__ __forceinline__ int pack_try_reserve(const PackDesc &D, const ClusterConst &c, PackState &st, PackCtx &x, int node, int t) {
    const int u = D.units[node];
    if (u >= c.base_units) return 0;                                   // cpu_free - 12 < 0 or mem_free - 60 < 0
    const int dev = node * c.G + x.lane;
    const bool fit = x.lane < c.G && pack_dev_fits(D, dev, x.m);
    const unsigned fb = __ballot_sync(RLGS_FULL, fit);
    if (!fb) return 0;
    st.n_free_nodes += (int)node_is_free(u + 1, c) - (int)node_is_free(u, c);
    const int have = __popc(fb);
    const unsigned taken = have <= x.gpc ? fb : lowest_bits(fb, x.gpc);
    int n_before = 0; int64_t dmem = 0; int newly_busy = 0, added = 0;
    if ((taken >> x.lane) & 1) {
        n_before = D.dn[dev];
        bool present = false;                                          // dict overwrite of an entry leaked earlier
        for (int k = 0; k < n_before; ++k) { int2 e = D.ent[dev * PACK_DEV_SLOTS + k]; present |= (e.x == x.job && e.y == t); }
        if (!present) {
            const int64_t before = D.dm[dev], after = before + min(x.m, D.cap_units);
            D.ent[dev * PACK_DEV_SLOTS + n_before] = make_int2(x.job, t);
            D.dn[dev] = n_before + 1; D.dm[dev] = after;
            dmem = min(after, D.cap_units) - min(before, D.cap_units);
            newly_busy = n_before == 0; added = 1;
        }
    }
    __syncwarp();
    const int hi = 31 - __clz(taken);                                  // Task.interfered is rewritten by every add_task: the last device wins
    const int nb_hi = __shfl_sync(RLGS_FULL, n_before, hi);
    x.interf = (x.interf & ~(1u << t)) | ((nb_hi >= 2 ? 1u : 0u) << t);
    if (x.bump && __any_sync(RLGS_FULL, ((taken >> x.lane) & 1) && n_before < 2)) x.bump &= ~(1u << t);   // task.duration = original (device.py:39-41)
#pragma unroll
    for (int o = 16; o; o >>= 1) { dmem += __shfl_xor_sync(RLGS_FULL, dmem, o); newly_busy += __shfl_xor_sync(RLGS_FULL, newly_busy, o); added += __shfl_xor_sync(RLGS_FULL, added, o); }
    st.mem_sum += dmem; st.busy_gpus += newly_busy;
    st.util_mu_sum += (int64_t)added * x.mu_q; st.util_var_sum += (int64_t)added * x.sd_q * x.sd_q;
    D.units[node] = u + 1;
    if (have < x.gpc) { D.jflag[x.job] = 1; __syncwarp(); return 2; }   // remember that the job has leaked entries
    const bool was = pack_node_idle(D, node);
    const int tk = D.ntk[node];
    __syncwarp();
    D.ntk[node] = tk + (1 << 16);
    pack_idle_delta(st, was, false);
    __syncwarp();
    return 1;
}

// Node.release_allocated_resources (node.py:71-91) for one task; `running` = the task sits in node.running_tasks
// (completion) instead of node.placed_tasks (backtracking).  On completion the devices of the node that are left with
// at most one task hand their flagged task to JobsManager.reset_interference (jobs_manager.py:175-187).
__device__ __forceinline__ void pack_release(const PackDesc &D, const ClusterConst &c, PackState &st, int lane, int node, int job, int t, bool running) {
    const int u = D.units[node];
    st.n_free_nodes += (int)node_is_free(u - 1, c) - (int)node_is_free(u, c);
    const int tk = D.ntk[node] - (running ? 1 : (1 << 16));
    const int np = D.npj[node];
    const PackJob pj = D.pj[job];
    const rlgs_job rec = D.trace[job];
    const int dev = node * c.G + lane;
    int64_t dmem = 0; int now_idle = 0, removed = 0, n_after = -1;
    if (lane < c.G) {
        int n = D.dn[dev], at = -1;
        for (int k = 0; k < n; ++k) { int2 e = D.ent[dev * PACK_DEV_SLOTS + k]; if (e.x == job && e.y == t) at = k; }
        if (at >= 0) {
            for (int k = at; k + 1 < n; ++k) D.ent[dev * PACK_DEV_SLOTS + k] = D.ent[dev * PACK_DEV_SLOTS + k + 1];
            const int64_t before = D.dm[dev], after = before - min(pj.mem, D.cap_units);
            D.dn[dev] = n - 1; D.dm[dev] = after;
            dmem = min(after, D.cap_units) - min(before, D.cap_units);
            now_idle = n == 1; removed = 1; n -= 1;
        }
        n_after = n;
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o; o >>= 1) { dmem += __shfl_xor_sync(RLGS_FULL, dmem, o); now_idle += __shfl_xor_sync(RLGS_FULL, now_idle, o); removed += __shfl_xor_sync(RLGS_FULL, removed, o); }
    st.mem_sum += dmem; st.busy_gpus -= now_idle;
    st.util_mu_sum -= (int64_t)removed * rec.util_mu_q; st.util_var_sum -= (int64_t)removed * rec.util_sd_q * rec.util_sd_q;
    D.units[node] = u - 1; D.ntk[node] = tk;
    pack_idle_delta(st, false, tk == 0 && np == 0);
    __syncwarp();
    if (!running) return;
    // reduce_interference_set: flagged tasks alone on a device of this node, whose job is in running_jobs
    int oj = -1, ot = 0;
    if (n_after == 1) {
        int2 e = D.ent[dev * PACK_DEV_SLOTS];
        if ((D.imask[e.x] >> e.y) & 1) if (D.pend[e.x] > 0) { oj = e.x; ot = e.y; }      // pend > 0 <=> the job is in running_jobs
    }
    unsigned cand = __ballot_sync(RLGS_FULL, oj >= 0);
    while (cand) {
        const int src = __ffs(cand) - 1; cand &= cand - 1;
        const int j2 = __shfl_sync(RLGS_FULL, oj, src), t2 = __shfl_sync(RLGS_FULL, ot, src);
        const uint32_t mk = D.imask[j2];
        if ((mk >> t2) & 1) {
            const uint32_t bm = D.bmask[j2]; const int pe = D.pend[j2];
            __syncwarp();
            D.imask[j2] = mk & ~(1u << t2);
            D.bmask[j2] = bm | (1u << t2);                              // duration = original + max(int(diff / 2), 5), diff = 0 or 5
            if (bm == 0) D.pend[j2] = pe + 5;                           // Job.get_duration() is the max over the tasks
        }
        __syncwarp();
    }
}

// horus_score of every node for one task of the job (horus.py:28-56); score < 0 marks a node that is not free or
// cannot take the task (get_free_nodes + Node.can_fit(pack=True), algorithm.py:52-56)
template <bool GANDIVA>
__device__ __forceinline__ void pack_score_nodes(const PackDesc &D, const ClusterConst &c, const PackParams &P, const PackCtx &x,
                                                 uint32_t replica, uint32_t tick, uint32_t attempt, uint32_t pass, double job_util) {
    for (int base = 0; base < c.N; base += 32) {
        const int i = base + x.lane;
        double best = -1.0;
        if (i < c.N) {
            const int u = D.units[i];
            if (node_is_free(u, c) && u < c.base_units) {
                double min_cost = 999.0; bool any = false;
                for (int g = 0; g < c.G; ++g) {
                    const int dev = i * c.G + g;
                    const int n = D.dn[dev];
                    const int64_t cur = min(D.dm[dev], D.cap_units);
                    if (!(n < PACK_DEV_SLOTS && D.cap_units - (cur + x.m) > D.margin_units)) continue;
                    any = true;
                    double util = 0.0;                                  // Device.get_current_utilization (device.py:48-54)
                    for (int k = 0; k < n; ++k) {
                        const int oj = D.ent[dev * PACK_DEV_SLOTS + k].x;
                        const PackJob o = D.pj[oj];
                        double v = o.util_avg;
                        if (P.rng_on && o.util_sd != 0.0) v = o.util_avg + o.util_sd * pack_draw(P.seed, replica, tick, attempt, pass, (uint32_t)dev, (uint32_t)k);
                        util += (v < 100.0) ? v : 100.0;
                        util = (util < 100.0) ? util : 100.0;
                    }
                    double cost;
                    if (!GANDIVA) {
                        const double mem_cost = ((double)(cur + x.m) * D.unit_mib) / D.cap_mib;
                        const double xx = util + job_util;
                        double y = 0.0 * xx + 4E-5; y = y * xx + -0.00302; y = y * xx + 1.16664;   // np.polyval(NV_2080_COEF, .)
                        cost = (mem_cost * 0.5) + (y * 0.5) + (double)n;
                    } else {                                            // gandiva_score (horus.py:6-25): current memory in MiB + the task's share
                        const double mem_cost = (double)cur * D.unit_mib + ((double)x.m * D.unit_mib) / D.cap_mib;
                        cost = (mem_cost * 0.5) + (util / 100) + (double)n;
                    }
                    if (cost < min_cost) min_cost = cost;
                }
                if (any) best = min_cost;
            }
            x.score[i] = best;
        }
    }
    __syncwarp();
}

// Tasks of the job that node i accepts one after another (try_reserve_and_placed_task in a loop), and whether the attempt
// after the last accepted task would leak: the node still passes Node.can_fit but fewer than task.gpu devices accept.
__device__ __forceinline__ int pack_node_capacity(const PackDesc &D, const ClusterConst &c, const PackCtx &x, int i, bool &leak) {
    leak = false;
    int r = c.base_units - D.units[i];
    if (r <= 0) return 0;
    int s[32], tot = 0;
    for (int g = 0; g < c.G; ++g) {
        const int dev = i * c.G + g, n = D.dn[dev];
        const int64_t room = D.cap_units - D.margin_units - min(D.dm[dev], D.cap_units);   // a task fits while cur + m < cap - margin
        int k = 0;
        if (n < PACK_DEV_SLOTS && x.m < room) k = x.m > 0 ? (int)min((room - 1) / x.m, (int64_t)(PACK_DEV_SLOTS - n)) : PACK_DEV_SLOTS - n;
        s[g] = k; tot += k;
    }
    if (x.gpc == 1) return min(r, tot);                                // one device per task: never too few devices
    int cap = 0;
    while (r > 0) {
        int have = 0;
        for (int g = 0; g < c.G; ++g) have += s[g] > 0;
        if (have < x.gpc) { leak = have >= 1; break; }
        int need = x.gpc;
        for (int g = 0; g < c.G && need > 0; ++g) if (s[g] > 0) { s[g] -= 1; need -= 1; }
        cap += 1; r -= 1;
    }
    return cap;
}

// horus_placement (algorithm.py:34-180).  1 = placed (tnode / planes[4] written), 0 = not placed, < 0 = error status.
template <bool GANDIVA>
__device__ __forceinline__ int pack_place(const PackDesc &D, const ClusterConst &c, const PackParams &P, PackState &st, PackCtx &x,
                                          uint32_t replica, uint32_t attempt) {
    const PackJob pj = D.pj[x.job];
    const int cap = pj.heap_cap;
    PACK_T0;
#ifdef PACK_PROFILE
    st.prof[11] += 1;
#endif
    // ---- score the nodes (pass 0); nothing can take a task, or a heap without slots: no trial, no side effect
    pack_score_nodes<GANDIVA>(D, c, P, x, replica, (uint32_t)st.d, attempt, 0u, pj.util_avg);
    PACK_T(2);
    int n_fit = 0;
    for (int i = x.lane; i < c.N; i += 32) n_fit += x.score[i] >= 0.0;
#pragma unroll
    for (int o = 16; o; o >>= 1) n_fit += __shfl_xor_sync(RLGS_FULL, n_fit, o);
    if (n_fit == 0 || cap <= 0) return 0;
    // ---- is the placement bound to fail without leaking?  Then every trial walks every rack, maps the same tasks on every
    // node with capacity and undoes them; all that stays is placed_jobs[job] on those nodes, popped from the home of each
    // trial in turn (algorithm.py:122-127): afterwards only the home of the LAST trial lacks the key.  Excluded: a job with
    // leaked entries (re-adding a task over its own leaked entry takes no slot, and undoing the trial removes that entry) and
    // a job with +5 tasks (the skipped add_task calls would have reset them).
    uint32_t capbits[4] = {0u, 0u, 0u, 0u};                            // bit b of word w: node (32 * (32 w + b) + lane) has capacity
    bool doomed;
    {
        int tot = 0; bool leak_any = false;
        for (int base = 0, step = 0; base < c.N; base += 32, ++step) {
            const int i = base + x.lane;
            bool lk = false;
            const int cp = i < c.N ? pack_node_capacity(D, c, x, i, lk) : 0;
            tot += cp; leak_any |= lk;
            if (cp > 0) capbits[step >> 5] |= 1u << (step & 31);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(RLGS_FULL, tot, o);
        leak_any = __any_sync(RLGS_FULL, leak_any);
        doomed = tot < x.T && !leak_any && D.jflag[x.job] == 0 && x.bump == 0;
    }
    int home_last = -1;
    if (doomed && !P.rng_on && c.N <= 512) {
        // Only the home of the last trial matters = the last element of the sorted heap.  Without draws every node is pushed T
        // times with one score; the heap keeps the k = min(cap, T * n_fit) lowest entries, so its largest score is the score of the
        // node whose copies cover rank k.  If no other node shares that score the last element is that node, whatever the order
        // of the sifts; otherwise the heap is simulated below.
        const int k = min(cap, x.T * n_fit);
        int found = -1;
        for (int i = x.lane; i < c.N; i += 32) {
            const double si = x.score[i];
            if (si < 0.0) continue;
            int lt = 0, eq = 0;
            for (int j = 0; j < c.N; ++j) { const double sj = x.score[j]; lt += (sj >= 0.0) & (sj < si); eq += sj == si; }
            if (x.T * lt < k && k <= x.T * (lt + eq) && eq == 1) found = i;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) found = max(found, __shfl_xor_sync(RLGS_FULL, found, o));
        home_last = found;
#ifdef PACK_PROFILE
        st.prof[4] += found >= 0;
#endif
    }
    // ---- keep the `cap` best nodes in a heap (one pass per task: the reference re-scores per task)
    int hlen = 0;
    if (home_last < 0) {
        if (x.T == 1 && cap == 1) {
            // a heap of one slot keeps, of the nodes with the lowest score, the one pushed last (a push with an equal or lower
            // score displaces the resident): arg-min with ties to the higher node id, no sifting needed
            double bs = 1e300; int bn = -1;
            for (int i = x.lane; i < c.N; i += 32) { const double sc = x.score[i]; if (sc >= 0.0 && sc <= bs) { bs = sc; bn = i; } }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double os = __shfl_xor_sync(RLGS_FULL, bs, o); const int on = __shfl_xor_sync(RLGS_FULL, bn, o);
                if (on >= 0 && (bn < 0 || os < bs || (os == bs && on > bn))) { bs = os; bn = on; }
            }
            if (bn >= 0) { hlen = 1; if (x.lane == 0) { x.hscore[0] = bs; x.hnode[0] = bn; } }
        } else
        for (int pass = 0; pass < x.T; ++pass) {
            if (pass > 0 && P.rng_on) pack_score_nodes<GANDIVA>(D, c, P, x, replica, (uint32_t)st.d, attempt, (uint32_t)pass, pj.util_avg);
            for (int base = 0; base < c.N; base += 32) {
                const int i = base + x.lane;
                const double sc = i < c.N ? x.score[i] : -1.0;
                unsigned fb = __ballot_sync(RLGS_FULL, sc >= 0.0);
                while (fb) {
                    const int src = __ffs(fb) - 1; fb &= fb - 1;
                    const double s1 = __shfl_sync(RLGS_FULL, sc, src);
                    if (x.lane == 0) pack_h_siftdown(x.hscore, x.hnode, hlen, s1, base + src);     // heappush (only lane 0 touches the heap arrays)
                    hlen += 1;
                    if (hlen > cap) { hlen -= 1; if (x.lane == 0) pack_h_pop(x.hscore, x.hnode, hlen); }   // heappop: drop the worst
                }
            }
        }
#ifdef PACK_PROFILE
        { long long _t1 = clock64(); if (doomed) { st.prof[12] += _t1 - _t0; st.prof[13] += 1; } else { st.prof[14] += 1; } }
#endif
        PACK_T(3);
        if (hlen == 0) return 0;
        // ---- sorted(nodes_stack, key=min_score): stable insertion sort of the heap array
        if (x.lane == 0) for (int a = 1; a < hlen; ++a) {
            const double vs = x.hscore[a]; const int vn = x.hnode[a];
            int b = a - 1;
            while (b >= 0 && x.hscore[b] > vs) { x.hscore[b + 1] = x.hscore[b]; x.hnode[b + 1] = x.hnode[b]; --b; }
            x.hscore[b + 1] = vs; x.hnode[b + 1] = vn;
        }
        __syncwarp();
        if (doomed) home_last = x.hnode[hlen - 1];
    }
    if (doomed) {                                                      // same end state without the trials
        for (int base = 0, step = 0; base < c.N; base += 32, ++step) {
            unsigned sb = __ballot_sync(RLGS_FULL, (capbits[step >> 5] >> (step & 31)) & 1u);
            while (sb) { const int node = base + __ffs(sb) - 1; sb &= sb - 1; pack_pj_set(D, st, node, x.job); }
        }
        pack_pj_pop(D, st, home_last, x.job);
        PACK_T(5);
        return 0;
    }
    PACK_T(4);
    // ---- one trial placement per heap entry; lane t keeps the node of task t of the current / best plan
    int best_nn = RLGS_NEVER, best_map = -1, cur_map = -1;
    // one task on one device and a single candidate: the trial takes the first fitting device of that node (its score says
    // there is one), is undone, and the real placement repeats it; placed_jobs is set, popped and set again.  Skip the rehearsal.
    const bool rehearsed = !(x.T == 1 && x.gpc == 1 && hlen == 1);
    if (!rehearsed) { best_nn = 1; best_map = x.hnode[0]; }
    for (int e = 0; rehearsed && e < hlen; ++e) {
        const int home = x.hnode[e];
        int cnt = 0;
        for (int t = 0; t < x.T; ++t) {
            const int r = pack_try_reserve(D, c, st, x, home, t);
            if (r == 1) { if (t != cnt) return RLGS_ERR_STATE; pack_pj_set(D, st, home, x.job); if (x.lane == cnt) cur_map = home; cnt++; }
            else if (r == 0) break;                                             // later tasks fail the same way, without side effects
        }
        // racks by distance from the home rack, ties towards the lower id (infrastructure.py:135-147)
        const int hr = home / P.nodes_per_rack;
        for (int dist = 0; dist < P.racks && cnt < x.T; ++dist) {
            for (int side = 0; side < (dist == 0 ? 1 : 2) && cnt < x.T; ++side) {
                const int rk = side == 0 ? hr - dist : hr + dist;
                if (rk < 0 || rk >= P.racks) continue;
                const int n0 = rk * P.nodes_per_rack, n1 = n0 + P.nodes_per_rack;
                for (int base = n0; base < n1 && cnt < x.T; base += 32) {
                    const int i = base + x.lane;
                    bool can = false;                                            // pre-filter: Node.can_fit(pack=True)
                    if (i < n1 && D.units[i] < c.base_units)
                        for (int g = 0; g < c.G; ++g) can |= pack_dev_fits(D, i * c.G + g, x.m);
                    unsigned cb = __ballot_sync(RLGS_FULL, can);
                    while (cb && cnt < x.T) {
                        const int node = base + __ffs(cb) - 1; cb &= cb - 1;
                        for (int t = cnt; t < x.T; ++t) {                        // every unmapped task tries this node
                            const int r = pack_try_reserve(D, c, st, x, node, t);
                            if (r == 1) { if (t != cnt) return RLGS_ERR_STATE; pack_pj_set(D, st, node, x.job); if (x.lane == cnt) cur_map = node; cnt++; }
                            else if (r == 0) break;
                            if (cnt >= x.T) break;
                        }
                    }
                }
            }
        }
        // distinct nodes of the plan (tasks fill node after node) and backtracking (algorithm.py:122-133)
        const int prev = __shfl_up_sync(RLGS_FULL, cur_map, 1);
        const int nn = __popc(__ballot_sync(RLGS_FULL, x.lane < cnt && (x.lane == 0 || prev != cur_map)));
        for (int t = 0; t < cnt; ++t) {
            const int node = __shfl_sync(RLGS_FULL, cur_map, t);
            if (t == 0) pack_pj_pop(D, st, node, x.job);
            pack_release(D, c, st, x.lane, node, x.job, t, false);
        }
        if (cnt >= x.T && nn < best_nn) { best_nn = nn; best_map = cur_map; }   // stable sort by len(nodes): first minimum
    }
    PACK_T(5);
    if (best_nn == RLGS_NEVER) return 0;
    // ---- place for real (algorithm.py:163-178)
    for (int t = 0; t < x.T; ++t) {
        const int node = __shfl_sync(RLGS_FULL, best_map, t);
        if (pack_try_reserve(D, c, st, x, node, t) != 1) return RLGS_ERR_UNSUPPORTED;   // the reference's `assert cnt == len(tasks)` fails
        pack_pj_set(D, st, node, x.job);
        D.tnode[pj.task_off + t] = (int16_t)node;
    }
    D.imask[x.job] = x.interf;
    __syncwarp();
    PACK_T(6);
    return 1;
}

// ---- --scheme yarn (ms_yarn_placement, algorithm.py:28-32) under the horus / gandiva schedules: the fit of yarn_place.cuh on a
// node view in global memory; devices are never shared, so there is no interference and no trial placement
__device__ __forceinline__ NodeView pack_yarn_view(const PackDesc &D) { NodeView nv; nv.units = D.units; nv.busy = D.ybusy; nv.ever = D.yever; nv.key = D.ykey; return nv; }
__device__ __forceinline__ int pack_yarn_attempt(const PackDesc &D, const ClusterConst &c, PackState &st, int lane, int job) {
    JobRec jr; { const int4 *q = reinterpret_cast<const int4 *>(D.trace + job); jr.a = q[0]; jr.b = q[1]; }
    const int off = D.pj[job].task_off;
    int sticky_idle = 0;                                               // yarn_place's own idle count assumes keys are never popped
    const PlaceResult pr = yarn_place(pack_yarn_view(D), c, jr, lane, D.plog, off, st.n_free_nodes, sticky_idle);
    if (!pr.ok) return 0;
    __syncwarp();
    int went_busy = 0;
    for (int base = 0; base < pr.nnodes; base += 32) {                  // one entry per node: tasks start running, placed_jobs gets the key
        const int e = base + lane;
        bool was = false;
        if (e < pr.nnodes) {
            const int2 en = D.plog[off + e];
            const int node = en.x & 0xffff, tasks = (en.x >> 16) & 0xffff;
            const int tk = D.ntk[node], np = D.npj[node];
            was = tk == 0 && np == 0;
            D.ntk[node] = tk + tasks; D.npj[node] = np + 1;
        }
        went_busy += __popc(__ballot_sync(RLGS_FULL, was));
    }
    st.idle_nodes -= went_busy;
    const int ndev = jr.tasks() * jr.gpc();
    const int64_t mu = jr.util() & 0xffff, sd = jr.util() >> 16;
    st.busy_gpus += ndev; st.mem_sum += jr.mem_term(); st.util_mu_sum += mu * ndev; st.util_var_sum += sd * sd * ndev;
    if (lane == 0) D.pcnt[job] = pr.nnodes;
    __syncwarp();
    return 1;
}
// completion (placed_jobs keeps the key, q3) or preemption (jobs_manager.py:150-173 pops it)
__device__ __forceinline__ void pack_yarn_release(const PackDesc &D, const ClusterConst &c, PackState &st, int lane, int job, bool preempt) {
    JobRec jr; { const int4 *q = reinterpret_cast<const int4 *>(D.trace + job); jr.a = q[0]; jr.b = q[1]; }
    const int off = D.pj[job].task_off, cnt = D.pcnt[job];
    int went_idle = 0;
    for (int base = 0; base < cnt; base += 32) {
        const int e = base + lane;
        const bool active = e < cnt;
        int2 en = make_int2(0, 0);
        if (active) en = D.plog[off + e];
        release_entry(pack_yarn_view(D), c, active, en, st.n_free_nodes);
        bool now = false;
        if (active) {
            const int node = en.x & 0xffff, tasks = (en.x >> 16) & 0xffff;
            const int tk = D.ntk[node] - tasks, np = D.npj[node] - (preempt ? 1 : 0);
            D.ntk[node] = tk; D.npj[node] = np;
            now = tk == 0 && np == 0;
        }
        went_idle += __popc(__ballot_sync(RLGS_FULL, now));
    }
    st.idle_nodes += went_idle;
    const int ndev = jr.tasks() * jr.gpc();
    const int64_t mu = jr.util() & 0xffff, sd = jr.util() >> 16;
    st.busy_gpus -= ndev; st.mem_sum -= jr.mem_term(); st.util_mu_sum -= mu * ndev; st.util_var_sum -= sd * sd * ndev;
    __syncwarp();
}

// ---- horus+ (schedule_horus_plus, algorithm.py:242-290): k-means of the queued jobs into K utilisation heaps, credit-based pick
__device__ __forceinline__ uint32_t plus_draw(uint32_t seed, uint32_t call, uint32_t elem, uint32_t n) {   // = _draw of oracle/ref_runner.py
    uint64_t h = pack_mix64((((uint64_t)seed << 32) | call) + 0x9E3779B97F4A7C15ull);
    h = pack_mix64(h ^ (uint64_t)elem);
    return (uint32_t)((h >> 11) % n);
}
__device__ __forceinline__ double plus_job_dist(const PlusFeat &a, const double *b) {   // job_dist (utils.py:4-12), left to right
    double s = fabs(a.f[0] - b[0]);
    s += fabs(a.f[1] - b[1]); s += fabs(a.f[2] - b[2]); s += fabs(a.f[3] - b[3]); s += fabs(a.f[4] - b[4]); s += fabs(a.f[5] - b[5]); s += fabs(a.f[6] - b[6]);
    return s;
}
__device__ __forceinline__ double plus_pairwise_leaf(const double *a, int n) {   // n <= 128
    if (n < 8) { double res = 0.; for (int i = 0; i < n; ++i) res += a[i]; return res; }
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) { r0 += a[i]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3]; r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7]; }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += a[i];
    return res;
}
// numpy's DOUBLE_pairwise_sum (np.mean = (0 + this) / n): blocks of <= 128 summed with 8 accumulators, halves (the left one a multiple
// of 8) combined as left + right.  The recursion is unrolled on an explicit stack: device threads have a small call stack.
__device__ __forceinline__ double plus_pairwise_sum(const double *a, int n) {
    const double *fa[24]; int fn[24], fs[24]; double fl[24];
    int sp = 0; double ret = 0.0;
    fa[0] = a; fn[0] = n; fs[0] = 0; sp = 1;
    while (sp > 0) {
        const int t = sp - 1;
        if (fs[t] == 0) {
            if (fn[t] <= 128) { ret = plus_pairwise_leaf(fa[t], fn[t]); sp -= 1; continue; }
            int n2 = fn[t] / 2; n2 -= n2 % 8;
            fs[t] = 1; fa[sp] = fa[t]; fn[sp] = n2; fs[sp] = 0; sp += 1;
        } else if (fs[t] == 1) {
            int n2 = fn[t] / 2; n2 -= n2 % 8;
            fl[t] = ret; fs[t] = 2; fa[sp] = fa[t] + n2; fn[sp] = fn[t] - n2; fs[sp] = 0; sp += 1;
        } else { ret = fl[t] + ret; sp -= 1; }
    }
    return ret;
}
// clusterize (utils.py:36-67) of kjobs[0..m): kassign[i] = queue.  cent_f = shared memory [K][8] scratch for the centroids' features.
__device__ __forceinline__ void plus_clusterize(const PackDesc &D, const PackParams &P, PackState &st, int lane, int m, double *cent_f) {
    const int K = P.plus_k;
    int my_cent = -1;                                                  // lane c keeps the job of centroid c
    { const uint32_t call = st.kcalls; st.kcalls += 1; if (lane < K) my_cent = D.kjobs[plus_draw(P.plus_seed, call, (uint32_t)lane, (uint32_t)m)]; }
    for (int i = lane; i < m; i += 32) { D.kassign[i] = -1; D.kold[i] = -1; }
    __syncwarp();
    for (int iter = 0; iter < 1000; ++iter) {
        bool diff = false;
        for (int i = lane; i < m; i += 32) diff |= D.kassign[i] != D.kold[i];
        if (!__any_sync(RLGS_FULL, diff) && iter != 0) break;
        for (int i = lane; i < m; i += 32) D.kold[i] = D.kassign[i];
        if (lane < K) { const PlusFeat cf = D.feat[my_cent]; for (int k = 0; k < 7; ++k) cent_f[lane * 8 + k] = cf.f[k]; }
        __syncwarp();
        for (int i = lane; i < m; i += 32) {                           // np.argmin over the centroids: first minimum
            const PlusFeat jf = D.feat[D.kjobs[i]];
            int best = 0; double bd = plus_job_dist(jf, cent_f);
            for (int c2 = 1; c2 < K; ++c2) { const double dd = plus_job_dist(jf, cent_f + c2 * 8); if (dd < bd) { bd = dd; best = c2; } }
            D.kassign[i] = best;
        }
        __syncwarp();
        for (int c2 = 0; c2 < K; ++c2) {
            int cnt = 0;                                               // members in list order -> ktmp (their transform_to_dist)
            for (int base = 0; base < m; base += 32) {
                const int i = base + lane;
                const bool mem = i < m && D.kassign[i] == c2;
                const unsigned mb = __ballot_sync(RLGS_FULL, mem);
                if (mem) D.ktmp[cnt + __popc(mb & ((1u << lane) - 1))] = D.feat[D.kjobs[i]].tdist;
                cnt += __popc(mb);
            }
            __syncwarp();
            int newc;
            if (cnt > 0) {
                double score = 0.0;
                if (lane == 0) { const double mean = (0.0 + plus_pairwise_sum(D.ktmp, cnt)) / (double)cnt; score = (double)(long long)mean; }   // .astype(int)
                score = __shfl_sync(RLGS_FULL, score, 0);
                double bd = 99999999999.0; int bi = 0x7fffffff;        // get_closest: first strict minimum in list order
                for (int i = lane; i < m; i += 32) if (D.kassign[i] == c2) { const double t = fabs(D.feat[D.kjobs[i]].tdist - score); if (t < bd) { bd = t; bi = i; } }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const double od = __shfl_xor_sync(RLGS_FULL, bd, o); const int oi = __shfl_xor_sync(RLGS_FULL, bi, o);
                    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
                }
                if (bi == 0x7fffffff) { st.status = RLGS_ERR_STATE; return; }
                newc = D.kjobs[bi];
            } else { newc = D.kjobs[plus_draw(P.plus_seed, st.kcalls, 0u, (uint32_t)m)]; st.kcalls += 1; }   // np.random.choice(len(jobs))
            if (lane == c2) my_cent = newc;
            __syncwarp();
        }
    }
}
// credit of queue qi at reference tick `ref` (job_queue_manager.py:115-127): median pending time x length
__device__ __forceinline__ double plus_credit(const PackDesc &D, int lane, int qi, int n, int ref) {
    if (n == 0) return 0.0;
    const int32_t *qj = D.qjob + (size_t)qi * D.J;
    const int r_lo = (n - 1) / 2, r_hi = n / 2;                        // order statistics of the arrival ticks
    int a_lo = 0, a_hi = 0;
    for (int i = lane; i < n; i += 32) {
        const int ai = D.trace[qj[i]].arrival_tick;
        int lt = 0, eq = 0;
        for (int j = 0; j < n; ++j) { const int aj = D.trace[qj[j]].arrival_tick; lt += aj < ai; eq += aj == ai; }
        if (lt <= r_lo && r_lo < lt + eq) a_lo = ai;
        if (lt <= r_hi && r_hi < lt + eq) a_hi = ai;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { a_lo = max(a_lo, __shfl_xor_sync(RLGS_FULL, a_lo, o)); a_hi = max(a_hi, __shfl_xor_sync(RLGS_FULL, a_hi, o)); }
    const double med = ((double)(ref - a_lo) + (double)(ref - a_hi)) / 2.0;    // np.median: mean of the two middle values
    const double mp = med > 0.0 ? med : 0.0;
    return mp < 1.0 ? (double)n : mp * (double)n;
}

// removes `job` from a singly linked calendar bucket (lane 0)
__device__ __forceinline__ void pack_chain_unlink(int32_t *head, int32_t *next, int bucket, int job) {
    int prev = -1, cur = head[bucket];
    while (cur >= 0 && cur != job) { prev = cur; cur = next[cur]; }
    if (cur < 0) return;
    if (prev < 0) head[bucket] = next[cur]; else next[prev] = next[cur];
}

template <bool GANDIVA, bool YARN, bool PLUS>
__global__ void __launch_bounds__(32, 8) pack_horus_kernel(const PackDesc *descs, PackState *states, PackParams P, ClusterConst c,
                                                        RowStore rs, int64_t *returns) {
    extern __shared__ __align__(16) unsigned char pack_smem[];
    double *sm_score = reinterpret_cast<double *>(pack_smem);
    double *sm_hscore = sm_score + ((c.N + 1) & ~1);
    int32_t *sm_hnode = reinterpret_cast<int32_t *>(sm_hscore + PACK_MAX_HEAP + 1);
    double *sm_cent = reinterpret_cast<double *>(sm_hnode + PACK_MAX_HEAP + 1);   // 8-byte aligned: (PACK_MAX_HEAP + 1) is even
    const int lane = lane_id();
    const PackDesc D = descs[blockIdx.x];
    PackState st = states[blockIdx.x];
    if (st.done) return;
    const bool rows_mode = rs.chunks != nullptr;
    const uint32_t replica = (uint32_t)(rs.replica + blockIdx.x);
    int tick_budget = P.tick_budget;
    if (rows_mode && (int64_t)st.d + tick_budget > (int64_t)rs.n_chunks * RLGS_ROW_CHUNK)
        tick_budget = (int)max((int64_t)0, (int64_t)rs.n_chunks * RLGS_ROW_CHUNK - st.d);
    const int J = D.J;
    if (st.d == 0 && st.cursor == 0) {                     // first launch of a run: empty cluster
        for (int i = lane; i < c.N; i += 32) { D.units[i] = 0; D.ntk[i] = 0; D.npj[i] = 0; }
        if (YARN) {
            const uint32_t empty_key = node_key(0, 0u, c);
            for (int i = lane; i < c.N; i += 32) { D.ybusy[i] = 0; D.ykey[i] = empty_key; }
            for (int i = lane; i < (c.N + 31) / 32; i += 32) D.yever[i] = 0;
        }
        for (int i = lane; i < c.D; i += 32) { D.dn[i] = 0; D.dm[i] = 0; }
        for (int i = lane; i < PACK_CAL_W; i += 32) { D.chead[i] = -1; D.shead[i] = -1; }
        for (size_t i = lane; i < (size_t)J * D.W; i += 32) D.pjbits[i] = 0;
        for (int i = lane; i < J; i += 32) { D.pend[i] = 0; D.imask[i] = 0; D.bmask[i] = 0; D.jflag[i] = 0; D.pproc[i] = 0; D.nstart[i] = 0; }
        __syncwarp();
    }
    const int d_stop = st.d + tick_budget;
    while (true) {
        if ((J - st.cursor) + st.r_pre == 0 && st.d > 0) { st.done = 1; break; }   // schedule.py:185 (the queue is not consulted, q2)
        if (st.d == d_stop) break;
        if (P.max_ticks > 0 && st.d >= P.max_ticks) { st.done = 1; st.status = RLGS_ERR_CAPACITY; break; }
        const int d = st.d;
        PACK_T0;

        // ---------------- arrivals (jobs_manager.py:228-241)
        if (!GANDIVA) {
            // horus: heappush in trace order (job_queue_manager.py:147-152); horus+: the queues are rebuilt below
            const int c0 = st.cursor;
            while (st.cursor < J) {
                const int job = st.cursor;
                const int arr = D.trace[job].arrival_tick;
                if (arr > d) break;
                if (PLUS) st.Q += 1; else pack_q_push(D, lane, st.Q, D.pj[job].util_avg, job);
                if (lane == 0) { D.lprev[job] = st.ltail; D.lnext[job] = -1; if (st.ltail >= 0) D.lnext[st.ltail] = job; }
                if (st.ltail < 0) st.lhead = job;
                st.ltail = job;
                __syncwarp();
                const int ql = st.Q;                              // queued jobs = heap length here (no look-ahead is out)
                if (ql == 1) { st.mlo = job; st.mrank = 0; }
                else if ((ql - 1) / 2 > st.mrank) { st.mlo = D.lnext[st.mlo]; st.mrank += 1; }
                st.sum_arr += arr;
                st.cursor += 1;
                if (st.Q > st.max_q) st.max_q = st.Q;
                __syncwarp();
            }
            if (PLUS && st.Q > 0) {
                // jobs_manager.insert (:115-140) runs every tick, also without arrivals: every queue is emptied in heappop order,
                // queue after queue, the new jobs are appended, the list is clustered again (clusterize) and pushed back in order
                int m = 0;
                for (int q = 0; q < P.plus_k; ++q) {
                    int n = st.pqn[q];
                    while (n > 0) { double key = 0.0; int job = 0; pack_qk_pop(D, lane, q, n, key, job); if (lane == 0) D.kjobs[m] = job; m += 1; }
                    st.pqn[q] = 0;
                }
                for (int i = c0 + lane; i < st.cursor; i += 32) D.kjobs[m + (i - c0)] = i;
                m += st.cursor - c0;
                __syncwarp();
                plus_clusterize(D, P, st, lane, m, sm_cent);
                if (st.status) { st.done = 1; break; }
                for (int i = 0; i < m; ++i) {
                    const int q = D.kassign[i], job = D.kjobs[i];
                    int n = st.pqn[q];
                    pack_qk_push(D, lane, q, n, D.pj[job].util_avg, job);
                    st.pqn[q] = n;
                }
            }
        } else {
            // gandiva: a plain list, queue.insert(i, job_i): the batch goes to the front in order (q1).  The queue is a stack
            // whose top (index Q - 1) is the front, so the batch is pushed last-to-first.
            int c1 = st.cursor;
            while (c1 < J && D.trace[c1].arrival_tick <= d) c1 += 1;
            const int k = c1 - st.cursor;
            for (int i = lane; i < k; i += 32) { const int job = c1 - 1 - i; D.qjob[st.Q + i] = job; D.qtick[job] = d; }
            st.sum_arr += (int64_t)k * d;
            st.Q += k; st.cursor = c1;
            if (st.Q > st.max_q) st.max_q = st.Q;
            __syncwarp();
        }

        PACK_T(0);
        // ---------------- _schedule (schedule.py:40-60) -> schedule_horus (algorithm.py:204-240) | schedule_fifo (:189-202)
        if (st.Q > 0 && st.n_free_nodes >= 1) {
            const int k = GANDIVA ? 1 : min(max(P.num_buffer, 0), st.Q);
            int my_job = -1, my_q = 0; double my_key = 0.0;
            if (PLUS) {
                // update_credits + np.argmax before every pop (algorithm.py:254-258); only the queue that lost a job changes
                double credits[PACK_MAX_Q];
                for (int q = 0; q < P.plus_k; ++q) credits[q] = plus_credit(D, lane, q, st.pqn[q], d);
                for (int a = 0; a < k; ++a) {
                    int qi = 0;
                    for (int q = 1; q < P.plus_k; ++q) if (credits[q] > credits[qi]) qi = q;
                    int n = st.pqn[qi];
                    if (n == 0) { st.status = RLGS_ERR_UNSUPPORTED; break; }       // pop() returns None in the reference: AttributeError
                    double key = 0.0; int job = 0;
                    pack_qk_pop(D, lane, qi, n, key, job);
                    st.pqn[qi] = n;
                    if (lane == a) { my_job = job; my_key = key; my_q = qi; }
                    credits[qi] = plus_credit(D, lane, qi, n, d);
                }
                if (st.status) { st.done = 1; break; }
            } else if (!GANDIVA) for (int a = 0; a < k; ++a) { double key = 0.0; int job = 0; pack_q_pop(D, lane, st.Q, key, job); if (lane == a) { my_job = job; my_key = key; } }
            else my_job = D.qjob[st.Q - 1];                   // the head of the list stays queued unless it is placed
            __syncwarp();
            PACK_T(1);
            int pos = -1, err = 0;
            for (int a = 0; a < k && pos < 0; ++a) {
                PackCtx x;
                x.lane = lane; x.job = __shfl_sync(RLGS_FULL, my_job, a);
                x.score = sm_score; x.hscore = sm_hscore; x.hnode = sm_hnode;
                const rlgs_job rec = D.trace[x.job];
                x.T = rec.tasks; x.gpc = rec.gpus_per_task; x.m = D.pj[x.job].mem; x.mu_q = rec.util_mu_q; x.sd_q = rec.util_sd_q;
                x.interf = D.imask[x.job]; x.bump = D.bmask[x.job];
                if (x.T > PACK_MAX_TASKS || D.pj[x.job].heap_cap > PACK_MAX_HEAP || D.pj[x.job].heap_cap < 0) { err = RLGS_ERR_UNSUPPORTED; break; }
                int r;
                if (YARN) r = pack_yarn_attempt(D, c, st, lane, x.job);
                else {
                    r = pack_place<GANDIVA>(D, c, P, st, x, replica, (uint32_t)a);
                    if (r < 0) { err = r; break; }
                    __syncwarp();
                    D.bmask[x.job] = x.bump;                  // trial add_task calls reset Task.duration even when the plan fails
                    __syncwarp();
                }
                if (r) pos = a;
            }
            if (err) { st.status = err; st.done = 1; break; }
#ifdef PACK_PROFILE
            _t0 = clock64();
#endif
            if (!GANDIVA) for (int a = 0; a < k; ++a) {     // jobs_manager.insert(look_ahead): heappush the rest in order
                const int job = __shfl_sync(RLGS_FULL, my_job, a); const double key = __shfl_sync(RLGS_FULL, my_key, a);
                const int q = __shfl_sync(RLGS_FULL, my_q, a);
                if (a == pos) continue;
                if (PLUS) { int n = st.pqn[q]; pack_qk_push(D, lane, q, n, key, job); st.pqn[q] = n; }   // back into the queue it came from
                else pack_q_push(D, lane, st.Q, key, job);
            }
            if (PLUS && pos >= 0) st.Q -= 1;
            __syncwarp();
            PACK_T(7);
            if (pos >= 0) {                                   // add_to_running -> start_job (schedule.py:164-167, jobs_manager.py:189-207)
                const int job = __shfl_sync(RLGS_FULL, my_job, pos);
                const rlgs_job rec = D.trace[job];
                const PackJob pj = D.pj[job];
                int tn = -1;
                if (!YARN && lane < rec.tasks) tn = D.tnode[pj.task_off + lane];
                if (!GANDIVA) {
                    const int p = D.lprev[job], n = D.lnext[job];
                    __syncwarp();
                    if (p >= 0) D.lnext[p] = n; else st.lhead = n;
                    if (n >= 0) D.lprev[n] = p; else st.ltail = p;
                    if (job == st.mlo) { if (n >= 0) st.mlo = n; else { st.mlo = p; st.mrank -= 1; } }
                    else if (job < st.mlo) st.mrank -= 1;
                    __syncwarp();
                    const int ql = st.Q;                      // queued jobs after the start
                    if (ql > 0) {
                        const int target = (ql - 1) / 2;
                        if (st.mrank < target) { st.mlo = D.lnext[st.mlo]; st.mrank += 1; }
                        else if (st.mrank > target) { st.mlo = D.lprev[st.mlo]; st.mrank -= 1; }
                    }
                    st.sum_arr -= rec.arrival_tick;
                } else {
                    st.sum_arr -= D.qtick[job];
                    st.Q -= 1;                                // queues[0].pop(0)
                }
                // the job runs until time_processed >= get_duration(): ticks still to do = duration (+5) - processed so far
                const int pb = D.pproc[job], ns = D.nstart[job];
                const int end = d + max(1, rec.dur_ticks + (D.bmask[job] ? 5 : 0) - pb);
                const int bk = end & (PACK_CAL_W - 1);
                const int hb = D.chead[bk];
                const int slice = d + PACK_QUANTA - pb % PACK_QUANTA;         // next tick with time_processed % 100 == 0
                const int sb = slice & (PACK_CAL_W - 1);
                const int hs = D.shead[sb];
                __syncwarp();
                D.planes[0][job] = d; D.nstart[job] = ns + 1;
                D.pend[job] = end; D.cbk[job] = bk;
                D.cnext[job] = hb; D.chead[bk] = job;
                if (GANDIVA) { D.sat[job] = slice; D.snext[job] = hs; D.shead[sb] = job; }
                // placed_tasks -> running_tasks on every node of the job (node.py:173-198); several tasks may share a node
                if (!YARN) for (int t = 0; t < rec.tasks; ++t) {
                    const int node = __shfl_sync(RLGS_FULL, tn, t);
                    const int v = D.ntk[node];
                    __syncwarp();
                    D.ntk[node] = v - (1 << 16) + 1;
                    __syncwarp();
                }
                st.R += 1; st.start_seq += 1;
                if (st.R > st.max_r) st.max_r = st.R;
                __syncwarp();
            }
        }

        PACK_T(8);
        // ---------------- delta_time += 1; step; release_finished_jobs in running_jobs (= start) order
        st.d = d + 1;
        {
            const int bk = st.d & (PACK_CAL_W - 1);
            int nf = 0;
            if (lane == 0) {                                   // jobs_to_finish is fixed before any release (jobs_manager.py:243-250)
                int prev = -1, cur = D.chead[bk];
                while (cur >= 0) {
                    const int nx = D.cnext[cur], pe = D.pend[cur];
                    if (pe == st.d || (pe & (PACK_CAL_W - 1)) != bk) {
                        if (prev < 0) D.chead[bk] = nx; else D.cnext[prev] = nx;
                        if (pe == st.d) { D.fin[nf] = cur; nf += 1; }
                        else { const int b2 = pe & (PACK_CAL_W - 1); D.cnext[cur] = D.chead[b2]; D.chead[b2] = cur; D.cbk[cur] = b2; }   // +5 moved it to another bucket
                    } else prev = cur;
                    cur = nx;
                }
            }
            __syncwarp();
            nf = __shfl_sync(RLGS_FULL, nf, 0);
            for (int done_n = 0; done_n < nf; ++done_n) {
                // next finisher in start order (one start per tick: start ticks are distinct)
                int best = -1, best_start = RLGS_NEVER, best_at = -1;
                for (int i = 0; i < nf; ++i) { const int jb = D.fin[i]; if (jb < 0) continue; const int s0 = D.planes[0][jb]; if (s0 < best_start) { best_start = s0; best = jb; best_at = i; } }
                __syncwarp();
                D.fin[best_at] = -1;
                const rlgs_job rec = D.trace[best];
                const int off = D.pj[best].task_off;
                if (YARN) pack_yarn_release(D, c, st, lane, best, false);
                else for (int t = 0; t < rec.tasks; ++t) pack_release(D, c, st, lane, (int)D.tnode[off + t], best, t, true);
                if (GANDIVA && lane == 0) pack_chain_unlink(D.shead, D.snext, D.sat[best] & (PACK_CAL_W - 1), best);
                __syncwarp();
                D.planes[1][best] = st.d;
                D.planes[2][st.F] = best;
                D.planes[3][best] = D.bmask[best] != 0;                         // Job.get_duration() - Job.duration = 5
                D.planes[4][best] = D.pproc[best] + (st.d - best_start);       // Job.time_processed()
                D.planes[5][best] = D.nstart[best];                            // Job.migration_count
                D.pend[best] = 0;                                              // running_jobs.pop
                st.F += 1; st.R -= 1;
                st.sum_jct += (int64_t)(st.d - rec.arrival_tick);
                __syncwarp();
            }
        }
        st.r_pre = st.R;                                       // schedule.py:195: counted before the post-tick plugin

        // ---------------- time_slice_check (algorithm.py:420-440): with a non-empty queue, every running job whose processed time
        // reached a multiple of 100 is preempted (jobs_manager.py:150-173) and re-queued at the front
        if (GANDIVA) {
            const int sb = st.d & (PACK_CAL_W - 1);
            const bool slicing = st.Q > 0;
            int nt = 0;
            if (lane == 0) {
                int prev = -1, cur = D.shead[sb];
                while (cur >= 0) {
                    const int nx = D.snext[cur];
                    if (D.sat[cur] == st.d) {
                        if (prev < 0) D.shead[sb] = nx; else D.snext[prev] = nx;
                        if (slicing) { D.fin[nt] = cur; nt += 1; }
                        else { const int s2 = st.d + PACK_QUANTA, b2 = s2 & (PACK_CAL_W - 1); D.sat[cur] = s2; D.snext[cur] = D.shead[b2]; D.shead[b2] = cur; }
                    } else prev = cur;
                    cur = nx;
                }
            }
            __syncwarp();
            nt = __shfl_sync(RLGS_FULL, nt, 0);
            for (int k2 = 0; k2 < nt; ++k2) {                  // every job of the todo list leaves running_jobs as its turn comes
                int best = -1, best_start = RLGS_NEVER, best_at = -1;
                for (int i = 0; i < nt; ++i) { const int jb = D.fin[i]; if (jb < 0) continue; const int s0 = D.planes[0][jb]; if (s0 < best_start) { best_start = s0; best = jb; best_at = i; } }
                __syncwarp();
                D.fin[best_at] = -1;
                const rlgs_job rec = D.trace[best];
                const int off = D.pj[best].task_off;
                const int pb = D.pproc[best];
                if (lane == 0) pack_chain_unlink(D.chead, D.cnext, D.cbk[best], best);
                __syncwarp();
                D.pend[best] = 0;                              // running_jobs.pop(job_id) comes first
                D.pproc[best] = pb + (st.d - best_start);
                D.planes[0][best] = -1;                        // the start column belongs to the run that finishes the job
                __syncwarp();
                if (YARN) pack_yarn_release(D, c, st, lane, best, true);
                else for (int t = 0; t < rec.tasks; ++t) {
                    const int node = (int)D.tnode[off + t];
                    pack_pj_pop(D, st, node, best);            // placed_jobs.pop(job_id), once per node
                    pack_release(D, c, st, lane, node, best, t, true);
                }
                const int qn = st.Q;
                __syncwarp();
                D.qjob[qn] = best; D.qtick[best] = st.d;       // Job.preempted(): pending_time = 0; insert([job]) at the front
                st.Q = qn + 1; st.sum_arr += st.d; st.R -= 1; st.preempts += 1;
                if (st.Q > st.max_q) st.max_q = st.Q;
                __syncwarp();
            }
        }

        PACK_T(9);
        // ---------------- stats row (schedule.py:95-133, 204-205)
        st.sumQ += st.Q; st.sumR += st.R;
        if (rows_mode) {
            int lo = 0, hi = 0, mx = 0;
            if (st.Q > 0) {
                if (!GANDIVA) {
                    const int a0 = D.trace[st.mlo].arrival_tick;
                    const int a1 = (st.Q & 1) ? a0 : D.trace[D.lnext[st.mlo]].arrival_tick;
                    lo = st.d - a1; hi = st.d - a0; mx = st.d - D.trace[st.lhead].arrival_tick;
                } else {                                       // the stack is ordered by entry tick: positions give the order statistics
                    lo = st.d - D.qtick[D.qjob[st.Q - 1 - (st.Q - 1) / 2]];
                    hi = st.d - D.qtick[D.qjob[st.Q - 1 - st.Q / 2]];
                    mx = st.d - D.qtick[D.qjob[0]];
                }
            }
            if (lane == 0) {
                rlgs_row *row = row_ptr(rs, blockIdx.x, st.d - 1);
                const int64_t sp = (int64_t)st.Q * st.d - st.sum_arr;
                int4 *o = reinterpret_cast<int4 *>(row);
                o[0] = make_int4(st.idle_nodes, st.busy_gpus, st.R, st.Q);
                o[1] = make_int4(st.F, lo, hi, mx);
                o[2] = make_int4((int)(uint32_t)sp, (int)(sp >> 32), (int)(uint32_t)st.mem_sum, (int)(st.mem_sum >> 32));
                o[3] = make_int4((int)(uint32_t)st.util_mu_sum, (int)(st.util_mu_sum >> 32), (int)(uint32_t)st.util_var_sum, (int)(st.util_var_sum >> 32));
            }
        }
        PACK_T(10);
    }
    st.events = (int64_t)st.cursor + st.start_seq + st.F + st.preempts;
    if (!st.done && st.status == RLGS_OK && st.d > 0 && (J - st.cursor) + st.r_pre == 0) st.done = 1;
    __syncwarp();
    if (lane == 0) {
        states[blockIdx.x] = st;
        if (st.done) returns[blockIdx.x] = -st.sum_jct;
    }
}
