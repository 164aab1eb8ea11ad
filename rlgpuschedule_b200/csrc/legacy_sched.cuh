// legacy_sched.cuh — the event-driven legacy schedules, one warp per replica.
//
//   sjf       restates smallest_first_sim_jobs   run_sim.py:162-287   (yarn placement = the live fit)
//   dlas-gpu  restates dlas_sim_jobs(gputime=1)  run_sim.py:664-947   (admission by GPU count)
//
// Both are dead code in the reference (undefined globals JOBS / CLUSTER / LOG / scheduler).  Parity is pinned
// given shims: oracle/ref_legacy_runner.py runs those loops unmodified with the missing globals supplied at
// run time, and the kernels reproduce the cluster.csv / job.csv of those runs byte for byte
// (tests/golden/{sjf,shortest,shortestgpu,dlasgpu,dlas}_*, tests/test_gpu_legacy.py).
//
// Per event every runnable job is touched once (the "per-event advance" of the north star):
// executed / pending time, demotion test, admission or re-placement, status flip, next-end and
// next-jump minima.  Runnable jobs are 32-byte entries streamed through the warp 32 at a time
// (one coalesced 1 KB read + write per chunk) in priority order:
//   dlas-gpu: per queue, RUNNING before PENDING (run_sim.py:838-848); the stable partition is a
//             warp-ballot compaction; greedy admission (run_sim.py:808-823) is a warp prefix scan
//             that restarts after each job that does not fit; demotions are appended to the next
//             queue in runnable (trace-index) order (run_sim.py:755-759).
//   sjf:      one array sorted by (num_gpu, trace index) — Python's stable list.sort(key=num_gpu)
//             on a list that only ever gets appended to (run_sim.py:237) — walked in order with
//             the yarn fit of yarn_place.cuh on a cluster emptied at every event (run_sim.py:241).
// Every job's last_check_time equals the previous event time (the sweep stamps all runnable jobs,
// run_sim.py:216-230 / :740-792), so one scalar replaces that field.
#pragma once
#include "yarn_place.cuh"

#define L_PENDING 1
#define L_RUNNING 2

struct Ent {
    int4 a;  // job, meta, duration, total_executed
    int4 b;  // executed (dlas) | attr0 (sjf), pending_time, preempt | resume<<16, last_pending (dlas) | attr1 (sjf)
    __device__ __forceinline__ int job() const { return a.x; }
    __device__ __forceinline__ int gpus() const { return a.y & 0xffff; }
    __device__ __forceinline__ int q() const { return (a.y >> 16) & 0xf; }
    __device__ __forceinline__ int status() const { return (a.y >> 20) & 3; }
    __device__ __forceinline__ bool started() const { return (a.y >> 22) & 1; }
    __device__ __forceinline__ void set_q(int q) { a.y = (a.y & ~(0xf << 16)) | (q << 16); }
    __device__ __forceinline__ void set_status(int s) { a.y = (a.y & ~(3 << 20)) | (s << 20); }
    __device__ __forceinline__ void set_started() { a.y |= 1 << 22; }
};

__device__ __forceinline__ Ent load_ent(const Ent *p) { Ent e; const int4 *q = reinterpret_cast<const int4 *>(p); e.a = q[0]; e.b = q[1]; return e; }
__device__ __forceinline__ void store_ent(Ent *p, const Ent &e) { int4 *q = reinterpret_cast<int4 *>(p); q[0] = e.a; q[1] = e.b; }

struct LegDesc {            // per-replica device pointers of the legacy kernels
    const rlgs_job *trace;
    Ent *buf[2];            // runnable entries (dlas: queues laid out back to back; sjf: buf[0] only)
    Ent *scratch_p;         // pending entries of the queue being rebuilt
    Ent *scratch_d[2];      // demotions into the current / next queue
    int32_t *end_list;      // jobs that ended at this event (sorted by trace index before they are logged)
    int32_t *planes[6];     // start, end, finish_order, pending_time, preempt, resume
    int2 *place_scratch;    // sjf: yarn_place's log output (ignored)
    int32_t J, cap;
};

struct LegState {
    int32_t t_prev, cursor, F, M;
    int32_t qlen[RLGS_MAX_QUEUES];
    int32_t cur, next_end, next_jump, n_rows;
    int32_t done, status, max_m, pad;
    int64_t events, sweep_jobs, sum_jct, demotions;
};

struct LegParams {
    int32_t nq;
    int32_t limit[RLGS_MAX_QUEUES];
    int32_t total_gpu, num_node, gpus_per_node;
    int32_t gputime;      // 1 = dlas-gpu (attained service in GPU-ticks), 0 = dlas (ticks)
    int32_t sort_mode;    // sjf kernel: 0 = num_gpu (sjf), 1 = remaining_time (shortest), 2 = remaining_gputime (shortest-gpu)
    int32_t event_budget;
    int64_t max_time;
};

// ---- small warp-cooperative sorts (lists are almost always <= 2 long) -------------------------
// ascending rank-sort of n distinct ints in global memory, using tmp[n]
__device__ __forceinline__ void sort_ints(int32_t *a, int32_t *tmp, int n, int lane) {
    if (n <= 1) return;
    for (int i = lane; i < n; i += 32) {
        int v = a[i], rank = 0;
        for (int j = 0; j < n; ++j) rank += a[j] < v;
        tmp[rank] = v;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) a[i] = tmp[i];
    __syncwarp();
}
// ascending rank-sort of n entries by job index: src -> dst
__device__ __forceinline__ void sort_ents(const Ent *src, Ent *dst, int n, int lane) {
    for (int i = lane; i < n; i += 32) {
        Ent e = load_ent(src + i);
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += src[j].a.x < e.a.x;
        store_ent(dst + rank, e);
    }
    __syncwarp();
}

// Greedy admission of one chunk in lane order (run_sim.py:808-823): a job runs iff free_gpu >= num_gpu
// at its turn; a job that does not fit is skipped and later smaller jobs may still fit.
__device__ __forceinline__ unsigned admit_chunk(int g, bool valid, int &free_gpu, int lane) {
    unsigned rem = __ballot_sync(RLGS_FULL, valid), adm = 0;
    while (rem && free_gpu > 0) {
        bool in = (rem >> lane) & 1;
        int incl = warp_incl_scan(in ? g : 0, lane);
        unsigned okb = __ballot_sync(RLGS_FULL, in && incl <= free_gpu);
        if (okb) {
            adm |= okb;
            free_gpu -= __shfl_sync(RLGS_FULL, incl, 31 - __clz(okb));
            rem &= ~okb;
        }
        if (!rem) break;
        rem &= rem - 1;  // the first remaining job does not fit: skip it
        int gm = ((rem >> lane) & 1) ? g : 0x7fffffff;
#pragma unroll
        for (int o = 16; o; o >>= 1) gm = min(gm, __shfl_xor_sync(RLGS_FULL, gm, o));
        if (gm > free_gpu) break;  // nothing left can fit
    }
    return adm;
}

// Jobs that never end (they never fit) still report their counters, like LOG would at shutdown.
__device__ __forceinline__ void flush_unfinished(const LegDesc &D, const Ent *buf, int M, int lane) {
    for (int i = lane; i < M; i += 32) {
        Ent e = load_ent(buf + i);
        D.planes[3][e.job()] = e.b.y;
        D.planes[4][e.job()] = e.b.z & 0xffff;
        D.planes[5][e.job()] = (e.b.z >> 16) & 0xffff;
    }
}

// bit 24 of Ent::a.y: the job is in the 'end_jobs' list that run_sim.py:706-710 attached to the head start event.
// `event = start_event; event['end_jobs'] = end_events[0]['end_jobs']` mutates the dict that stays at JOBS.job_events[0];
// when the queue-jump test (:715-717) then replaces the event, the list survives and is honoured when that start event is
// finally handled: its jobs complete there whatever their status (demoted / preempted meanwhile).  Pinned by the
// dlasgpu_* fixtures (the reference's code run unmodified).
#define L_ATTACHED (1 << 24)

struct DlasEvent {
    int time, d, q, nq;
    int end_ref;       // predicted end time that marks the natural end jobs: the event time, or next_end at a jump event
    int is_jump;       // queue-jump event: nothing ends, no arrivals
    int end_by_flag;   // start event with end_time > start_time: 'end_jobs' is whatever list is attached to it
    int flag_op;       // 0 keep, 1 clear (the start event is popped), 2 set from the natural end test (list attached, event replaced by a jump)
    int free_gpu, n_run, n_pend;
    int lane_end, lane_jump;   // per-lane minima, reduced across the warp once per event
    int lane_flips;            // per-lane count of resumes + preemptions, reduced once per event
    int wr, wp, nd, ne;
    int64_t events, demotions;
};

// Processes up to 32 entries of queue `ev.q` (one per lane).  mode 0 = resident entries (sweep them),
// 1 = entries demoted into this queue at this event (already swept), 2 = new arrivals.
__device__ __forceinline__ void dlas_chunk(const LegDesc &D, const LegParams &P, DlasEvent &ev, Ent e, bool valid, int mode,
                                           Ent *dst, Ent *dem_out, int t_prev, int lane) {
    bool ended = false, demote = false;
    if (valid && mode == 0) {
        // end_events hold the RUNNING jobs with the smallest predicted end (run_sim.py:908-922)
        const bool natural = e.status() == L_RUNNING && t_prev + e.a.z - e.a.w == ev.end_ref;
        if (ev.flag_op == 2) e.a.y = natural ? (e.a.y | L_ATTACHED) : (e.a.y & ~L_ATTACHED);
        ended = !ev.is_jump && (ev.end_by_flag ? (e.a.y & L_ATTACHED) != 0 : natural);
        if (ev.flag_op == 1) e.a.y &= ~L_ATTACHED;
        if (!ended) {                                                          // ended jobs left runnable_jobs before the sweep (:721-727)
            if (e.status() == L_RUNNING) {
                e.a.w += ev.d; e.b.x += ev.d;                                  // total_executed, executed (:741-744)
                if (ev.q < ev.nq - 1 && (P.gputime ? (int64_t)e.b.x * e.gpus() : (int64_t)e.b.x) >= P.limit[ev.q]) demote = true;  // :747-759
            } else {
                e.b.y += ev.d;                                                 // pending_time (:765-767)
                if (e.b.x > 0) e.b.w += ev.d;                                  // last_pending_time (:768-769)
            }
        }
    }
    // ---- ended jobs leave (run_sim.py:721-727)
    unsigned eb = __ballot_sync(RLGS_FULL, ended);
    if (ended) {
        int job = e.job();
        D.planes[1][job] = ev.time;
        D.planes[3][job] = e.b.y;
        D.planes[4][job] = e.b.z & 0xffff;
        D.planes[5][job] = (e.b.z >> 16) & 0xffff;
        D.end_list[ev.ne + __popc(eb & ((1u << lane) - 1))] = job;
    }
    ev.ne += __popc(eb);
    // ---- demoted jobs move to the next queue, still RUNNING until admission says otherwise
    unsigned db = __ballot_sync(RLGS_FULL, demote);
    if (demote) {
        e.set_q(ev.q + 1);
        store_ent(dem_out + ev.nd + __popc(db & ((1u << lane) - 1)), e);
    }
    ev.nd += __popc(db); ev.demotions += __popc(db); ev.events += __popc(db);
    // ---- admission in queue order (run_sim.py:808-823)
    bool stay = valid && !ended && !demote;
    unsigned adm = admit_chunk(e.gpus(), stay, ev.free_gpu, lane);
    bool run = (adm >> lane) & 1;
    bool resumed = stay && run && e.status() == L_PENDING, preempted = stay && !run && e.status() == L_RUNNING;
    if (resumed) {                                                             // :830-834
        e.set_status(L_RUNNING);
        e.b.z += 1 << 16;                                                       // resume += 1
        if (!e.started()) { e.set_started(); D.planes[0][e.job()] = ev.time; }
    } else if (preempted) {                                                    // :825-829
        e.set_status(L_PENDING);
        e.b.z = (e.b.z & ~0xffff) | ((e.b.z + 1) & 0xffff);                   // preempt += 1
    }
    ev.lane_flips += (int)resumed + (int)preempted;
    // ---- stable partition: RUNNING to dst, PENDING to the scratch (run_sim.py:838-848)
    unsigned rb = __ballot_sync(RLGS_FULL, stay && run), pb = __ballot_sync(RLGS_FULL, stay && !run);
    if (stay && run) store_ent(dst + ev.wr + __popc(rb & ((1u << lane) - 1)), e);
    if (stay && !run) store_ent(D.scratch_p + ev.wp + __popc(pb & ((1u << lane) - 1)), e);
    ev.wr += __popc(rb); ev.wp += __popc(pb);
    ev.n_run += __popc(rb); ev.n_pend += __popc(pb);
    // ---- next end / next queue jump over RUNNING jobs (run_sim.py:908-943): per-lane minima now, one warp reduction per event
    if (stay && run) {
        ev.lane_end = min(ev.lane_end, ev.time + e.a.z - e.a.w);
        if (e.q() < ev.nq - 1) {
            int num = P.limit[e.q()] - e.b.x, g = e.gpus();                    // as written: executed_time, not gpu-time
            int c;                                                             // math.ceil(num / g); plain `num` for time-based dlas (:932)
            if (!P.gputime) c = num;
            else if ((g & (g - 1)) == 0) { int sh = 31 - __clz(g); c = num >= 0 ? (num + g - 1) >> sh : -((-num) >> sh); }
            else c = num >= 0 ? (num + g - 1) / g : -((-num) / g);
            ev.lane_jump = min(ev.lane_jump, c + ev.time);
        }
    }
}

// MB = minimum resident blocks per SM the register allocation leaves room for: 20 (16 for sjf) = the registers the code wants
// (94 / 124: 20 / 16 warps per SM, the fastest choice up to 148 x 20 replicas), 32 = at most 64 registers, small spills, 32 warps
// per SM (measured on the 60k-job trace: +17 % events/s at 148 x 32 replicas, -11 % at 148 x 20).  rlgs_api.cu picks by the
// size of the launch.
template <int MB>
__global__ void __launch_bounds__(32, MB) dlas_gpu_kernel(const LegDesc *__restrict__ descs, LegState *__restrict__ states, LegParams P,
                                                      RowStore rs, int64_t *__restrict__ returns) {
    const int lane = lane_id();
    const LegDesc D = descs[blockIdx.x];
    LegState st = states[blockIdx.x];
    if (st.done || st.status != RLGS_OK) return;
    const int J = D.J;
    int budget = P.event_budget;
    const bool rows_on = rs.chunks != nullptr;
    int next_arr = st.cursor < J ? D.trace[st.cursor].arrival_tick : RLGS_NEVER;   // refreshed only when arrivals are consumed
    while (true) {
        if ((J - st.cursor) + st.M == 0) { st.done = 1; break; }                          // run_sim.py:679
        if ((J - st.cursor) == 0 && st.next_end == RLGS_NEVER) { st.done = 1; break; }    // :680-682 cluster too small
        if (budget-- <= 0) break;
        if (rows_on && st.n_rows >= (int64_t)rs.n_chunks * RLGS_ROW_CHUNK) break;
        int event_time = min(next_arr, st.next_end);                                      // :684-713
        bool has_start = next_arr <= st.next_end;
        const bool attach = next_arr == st.next_end && next_arr != RLGS_NEVER;            // :706-710 start_event['end_jobs'] = ...
        const bool is_jump = event_time > st.next_jump;
        if (is_jump) { event_time = st.next_jump; has_start = false; }                    // :715-717
        if (P.max_time > 0 && event_time > P.max_time) { st.done = 1; st.status = RLGS_ERR_CAPACITY; break; }
        // new arrivals at this event (run_sim.py:730-737)
        int k_arr = 0;
        if (has_start) {
            int pos = st.cursor;
            while (pos < J) {
                int i2 = pos + lane;
                bool a2 = i2 < J && D.trace[i2].arrival_tick == event_time;
                int c2 = __popc(__ballot_sync(RLGS_FULL, a2));
                k_arr += c2;
                if (c2 < 32) break;
                pos += 32;
            }
        }
        if (st.M + k_arr > D.cap) { st.status = RLGS_ERR_CAPACITY; st.done = 1; break; }
        DlasEvent ev;
        ev.time = event_time; ev.d = event_time - st.t_prev; ev.nq = P.nq;
        ev.is_jump = is_jump; ev.end_ref = (is_jump && attach) ? st.next_end : event_time;
        ev.end_by_flag = has_start && !attach;
        ev.flag_op = has_start ? 1 : ((is_jump && attach) ? 2 : 0);
        ev.free_gpu = P.total_gpu; ev.n_run = ev.n_pend = 0; ev.lane_end = ev.lane_jump = RLGS_NEVER; ev.lane_flips = 0;
        ev.ne = 0; ev.events = 0; ev.demotions = 0; ev.nd = 0;
        const Ent *src = D.buf[st.cur];
        Ent *dst = D.buf[st.cur ^ 1];
        int src_off = 0, dst_off = 0, n_dem_in = 0, dsel = 0;
        int new_qlen[RLGS_MAX_QUEUES];
        for (int q = 0; q < P.nq; ++q) {
            ev.q = q; ev.wr = 0; ev.wp = 0; ev.nd = 0;
            Ent *dq = dst + dst_off;
            Ent *dem_out = D.scratch_d[dsel ^ 1];
            const int len = st.qlen[q];
            for (int b = 0; b < len; b += 32) {
                bool valid = b + lane < len;
                Ent e; e.a = e.b = make_int4(0, 0, 0, 0);
                if (valid) e = load_ent(src + src_off + b + lane);
                dlas_chunk(D, P, ev, e, valid, 0, dq, dem_out, st.t_prev, lane);
            }
            // jobs demoted into this queue at this event: appended in runnable (trace) order
            if (n_dem_in > 0) {
                const Ent *din = D.scratch_d[dsel];
                for (int b = 0; b < n_dem_in; b += 32) {
                    bool valid = b + lane < n_dem_in;
                    Ent e; e.a = e.b = make_int4(0, 0, 0, 0);
                    if (valid) e = load_ent(din + b + lane);
                    dlas_chunk(D, P, ev, e, valid, 1, dq, dem_out, st.t_prev, lane);
                }
            }
            if (q == 0 && k_arr > 0) {
                for (int b = 0; b < k_arr; b += 32) {
                    int idx = st.cursor + b + lane;
                    bool valid = b + lane < k_arr;
                    Ent e; e.a = e.b = make_int4(0, 0, 0, 0);
                    if (valid) {
                        rlgs_job jr = D.trace[idx];
                        e.a = make_int4(idx, (int)jr.gpus | (L_PENDING << 20), jr.dur_ticks, 0);   // move_to_runnable, q_id = 0
                    }
                    dlas_chunk(D, P, ev, e, valid, 2, dq, dem_out, st.t_prev, lane);
                }
            }
            __syncwarp();
            // pending jobs go behind the running ones
            for (int b = 0; b < ev.wp; b += 32)
                if (b + lane < ev.wp) store_ent(dq + ev.wr + b + lane, load_ent(D.scratch_p + b + lane));
            new_qlen[q] = ev.wr + ev.wp;
            src_off += len; dst_off += new_qlen[q];
            // sort this queue's demotions by trace index for the next queue
            n_dem_in = ev.nd;
            __syncwarp();
            if (n_dem_in > 1) { sort_ents(dem_out, D.scratch_p, n_dem_in, lane); for (int b = lane; b < n_dem_in; b += 32) store_ent(dem_out + b, load_ent(D.scratch_p + b)); __syncwarp(); }
            dsel ^= 1;
        }
        // ended jobs are logged in runnable (trace) order (run_sim.py:721-727 walks end_jobs built from runnable_jobs)
        if (ev.ne > 0) {
            __syncwarp();
            sort_ints(D.end_list, reinterpret_cast<int32_t *>(D.scratch_p), ev.ne, lane);
            for (int b = lane; b < ev.ne; b += 32) {
                int job = D.end_list[b];
                D.planes[2][st.F + b] = job;
            }
            for (int b = 0; b < ev.ne; b += 32) {
                int64_t jct = 0;
                if (b + lane < ev.ne) jct = ev.time - D.trace[D.end_list[b + lane]].arrival_tick;
#pragma unroll
                for (int o = 16; o; o >>= 1) jct += __shfl_xor_sync(RLGS_FULL, jct, o);
                st.sum_jct += jct;
            }
            st.F += ev.ne;
        }
        int new_end = ev.lane_end, new_jump = ev.lane_jump, flips = ev.lane_flips;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            new_end = min(new_end, __shfl_xor_sync(RLGS_FULL, new_end, o));
            new_jump = min(new_jump, __shfl_xor_sync(RLGS_FULL, new_jump, o));
            flips += __shfl_xor_sync(RLGS_FULL, flips, o);
        }
        st.events += ev.events + flips + ev.ne + k_arr;
        st.demotions += ev.demotions;
        for (int q = 0; q < P.nq; ++q) st.qlen[q] = new_qlen[q];
        st.M = ev.n_run + ev.n_pend;
        st.sweep_jobs += st.M;   // runnable jobs swept at this event (after ends left and arrivals joined)
        if (st.M > st.max_m) st.max_m = st.M;
        if (k_arr > 0) { st.cursor += k_arr; next_arr = st.cursor < J ? D.trace[st.cursor].arrival_tick : RLGS_NEVER; }
        st.cur ^= 1;
        st.next_end = new_end; st.next_jump = new_jump;
        st.t_prev = event_time;
        if (rows_on && lane == 0) {                                                       // LOG.checkpoint, count branch (log.py:225-238)
            int busy = P.total_gpu - ev.free_gpu;
            int busy_node = (busy + P.gpus_per_node - 1) / P.gpus_per_node;
            int4 *o = reinterpret_cast<int4 *>(row_ptr(rs, blockIdx.x, st.n_rows));
            o[0] = make_int4(P.num_node - busy_node, busy, ev.n_run, ev.n_pend);
            o[1] = make_int4(st.F, event_time, busy_node, 0);
            o[2] = make_int4(0, 0, 0, 0); o[3] = make_int4(0, 0, 0, 0);
        }
        st.n_rows += 1;
        __syncwarp();
    }
    if (st.done) flush_unfinished(D, D.buf[st.cur], st.M, lane);
    if (lane == 0) {
        states[blockIdx.x] = st;
        if (st.done) returns[blockIdx.x] = -st.sum_jct;
    }
}

// ================================= sjf ==========================================================
// Entry words for sjf: a = {job, meta(gpus|status<<20|started<<22), duration, total_executed},
//                      b = {attr0 (gpus|tasks<<16), pending_time, preempt|resume<<16, attr1 (gpc|least<<16|fits<<31)}
struct SjfSmem { NodeView nv; };

__host__ __device__ inline size_t sjf_smem_bytes(int N) { return (3 * (size_t)N + (size_t)((N + 31) / 32)) * 4; }

template <int MB>   // see dlas_gpu_kernel (10k-job trace: +31 % events/s at 148 x 32 replicas)
__global__ void __launch_bounds__(32, MB) sjf_yarn_kernel(const LegDesc *__restrict__ descs, LegState *__restrict__ states, LegParams P,
                                                      ClusterConst c, RowStore rs, int64_t *__restrict__ returns) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = lane_id();
    const LegDesc D = descs[blockIdx.x];
    LegState st = states[blockIdx.x];
    if (st.done || st.status != RLGS_OK) return;
    NodeView nv;
    {
        int32_t *w = reinterpret_cast<int32_t *>(smem_raw);
        nv.units = w; w += c.N; nv.busy = reinterpret_cast<uint32_t *>(w); w += c.N; nv.ever = reinterpret_cast<uint32_t *>(w);
        w += (c.N + 31) / 32; nv.key = reinterpret_cast<uint32_t *>(w);
    }
    const int nw = 2 * c.N + (c.N + 31) / 32;
    const uint32_t empty_key = node_key(0, 0u, c);
    const int J = D.J;
    Ent *buf = D.buf[0];
    int budget = P.event_budget;
    const bool rows_on = rs.chunks != nullptr;
    while (true) {
        if ((J - st.cursor) + st.M == 0) { st.done = 1; break; }                          // run_sim.py:168
        if ((J - st.cursor) == 0 && st.next_end == RLGS_NEVER) { st.done = 1; break; }    // :169-171
        if (budget-- <= 0) break;
        if (rows_on && st.n_rows >= (int64_t)rs.n_chunks * RLGS_ROW_CHUNK) break;
        const int next_arr = st.cursor < J ? D.trace[st.cursor].arrival_tick : RLGS_NEVER;
        const int event_time = min(next_arr, st.next_end);                                // :173-193
        const bool has_start = next_arr <= st.next_end;
        if (P.max_time > 0 && event_time > P.max_time) { st.done = 1; st.status = RLGS_ERR_CAPACITY; break; }
        const int d = event_time - st.t_prev;
        int n_events = 0;
        // ---- arrivals: insert at their sorted position (stable sort by num_gpu, run_sim.py:208-214,237)
        if (has_start) {
            while (st.cursor < J) {
                rlgs_job jr = D.trace[st.cursor];
                if (jr.arrival_tick != event_time) break;
                if (st.M + 1 > D.cap) { st.status = RLGS_ERR_CAPACITY; st.done = 1; break; }
                // static key (sjf): position = number of entries with num_gpu <= this job's; dynamic keys
                // (shortest / shortest-gpu): append, the stable sort below finds the place (run_sim.py:343-352,380-385)
                int pos = 0;
                if (P.sort_mode == 0) {
                    for (int b = 0; b < st.M; b += 32) {
                        bool le = b + lane < st.M && (buf[b + lane].a.y & 0xffff) <= (int)jr.gpus;
                        pos += __popc(__ballot_sync(RLGS_FULL, le));
                    }
                } else pos = st.M;
                // shift [pos, M) up by one, from the back
                for (int hi = st.M; hi > pos; hi -= 32) {
                    int i = hi - 1 - lane;
                    Ent e; bool v = i >= pos;
                    if (v) e = load_ent(buf + i);
                    __syncwarp();
                    if (v) store_ent(buf + i + 1, e);
                    __syncwarp();
                }
                if (lane == 0) {
                    Ent e;
                    const int32_t *w = reinterpret_cast<const int32_t *>(&jr);
                    e.a = make_int4(st.cursor, (int)jr.gpus | (L_PENDING << 20) | (1 << 23), jr.dur_ticks, 0);   // bit 23: arrived at this event
                    e.b = make_int4(w[2], 0, 0, w[3]);
                    store_ent(buf + pos, e);
                }
                __syncwarp();
                st.M += 1; st.cursor += 1; n_events += 1;
            }
            if (st.status != RLGS_OK) break;
        }
        // ---- CLUSTER.empty_infra() (run_sim.py:241)
        for (int i = lane; i < nw; i += 32) nv.units[i] = 0;
        for (int i = lane; i < c.N; i += 32) nv.key[i] = empty_key;
        __syncwarp();
        int n_free_nodes = c.free_limit > 0 ? c.N : 0, idle_unused = c.N;   // yarn_place's sticky idle-node counter is a fifo statistic
        int w_out = 0, n_run = 0, n_pend = 0, new_end = RLGS_NEVER;
        const Ent *src = buf;
        const bool fused = P.sort_mode == 0;   // static key: end / sweep / re-place in one pass over the sorted array
        if (!fused) {
            // ---- pass A: ended jobs leave (logged in the previous event's order), sweep (run_sim.py:333-364)
            int w = 0;
            for (int b = 0; b < st.M; b += 32) {
                bool valid = b + lane < st.M;
                Ent e; e.a = e.b = make_int4(0, 0, 0, 0);
                if (valid) e = load_ent(buf + b + lane);
                bool ended = valid && e.status() == L_RUNNING && st.t_prev + e.a.z - e.a.w == event_time;
                unsigned eb = __ballot_sync(RLGS_FULL, ended);
                if (ended) {
                    int job = e.job();
                    D.planes[1][job] = event_time; D.planes[3][job] = e.b.y;
                    D.planes[4][job] = e.b.z & 0xffff; D.planes[5][job] = (e.b.z >> 16) & 0xffff;
                    D.planes[2][st.F + __popc(eb & ((1u << lane) - 1))] = job;
                }
                int64_t jct = ended ? event_time - D.trace[e.job()].arrival_tick : 0;
#pragma unroll
                for (int o = 16; o; o >>= 1) jct += __shfl_xor_sync(RLGS_FULL, jct, o);
                st.sum_jct += jct; st.F += __popc(eb); n_events += __popc(eb);
                bool stay = valid && !ended;
                if (stay) {
                    if (e.a.y & (1 << 23)) e.a.y &= ~(1 << 23);
                    else if (e.status() == L_RUNNING) e.a.w += d; else e.b.y += d;
                }
                unsigned sb = __ballot_sync(RLGS_FULL, stay);
                __syncwarp();
                if (stay) store_ent(buf + w + __popc(sb & ((1u << lane) - 1)), e);
                w += __popc(sb);
                __syncwarp();
            }
            st.M = w;
            // ---- pass B: stable sort by the dynamic key (Python list.sort: ties keep the previous order)
            for (int bi = 0; bi < st.M; bi += 32) {
                bool vi = bi + lane < st.M;
                Ent ei; ei.a = ei.b = make_int4(0, 0, 0, 0);
                if (vi) ei = load_ent(buf + bi + lane);
                const int64_t rem_i = ei.a.z - ei.a.w;
                const int64_t key_i = P.sort_mode == 1 ? rem_i : rem_i * ei.gpus();
                int rank = 0;
                for (int bj = 0; bj < st.M; bj += 32) {
                    bool vj = bj + lane < st.M;
                    int64_t key_j = 0x7fffffffffffffffll;
                    if (vj) { int4 a = buf[bj + lane].a; int64_t r = a.z - a.w; key_j = P.sort_mode == 1 ? r : r * (a.y & 0xffff); }
                    for (int k = 0; k < 32; ++k) {
                        int64_t kk = __shfl_sync(RLGS_FULL, key_j, k);
                        int pj = bj + k;
                        rank += (pj < st.M) && (kk < key_i || (kk == key_i && pj < bi + lane));
                    }
                }
                if (vi) store_ent(D.scratch_p + rank, ei);
            }
            __syncwarp();
            src = D.scratch_p;
        }
        // ---- pass in priority order: (fused: drop ended jobs, sweep,) re-place, flip, minima
        const int M0 = st.M;
        // Unit jobs (one GPU, one task, a device accepts it) lead the sjf order, and on the emptied cluster the first-fit of a run
        // of them is known in closed form: the r-th one lands on node r / c_node, device r % c_node, c_node = min(GPUs, task units)
        // per node.  While nothing else has been placed yet, a whole chunk of them is handled lane-parallel (one entry per lane
        // instead of one entry per warp step); the node arrays are materialised from the count before the first ordinary
        // placement and before the row statistics.  A chunk that starts with unit jobs and continues with others is split.
        const int c_node = min(c.G, c.base_units), unit_cap = c_node * c.N;
        bool pure = fused;          // every entry placed so far in this event was a unit job
        int unit_count = 0;         // unit jobs placed so far
        auto materialise = [&]() {
            for (int i = lane; i < c.N; i += 32) {
                const int placed = min(max(unit_count - i * c_node, 0), c_node);
                nv.units[i] = placed; nv.busy[i] = placed >= 32 ? 0xffffffffu : ((1u << placed) - 1u); nv.key[i] = node_key(placed, nv.busy[i], c);
            }
            int nf = 0;
            for (int b2 = 0; b2 < c.N; b2 += 32) {
                const int i = b2 + lane;
                nf += __popc(__ballot_sync(RLGS_FULL, i < c.N && node_is_free(min(max(unit_count - i * c_node, 0), c_node), c)));
            }
            n_free_nodes = nf;
            __syncwarp();
        };
        for (int b = 0; b < M0; b += 32) {
            const int cnt = min(32, M0 - b);
            Ent mine; mine.a = mine.b = make_int4(0, 0, 0, 0);
            if (lane < cnt) mine = load_ent(src + b + lane);
            int k0 = 0;                 // entries of this chunk already handled by the lane-parallel path
            if (pure) {
                const bool unit = lane < cnt && mine.b.x == (1 | (1 << 16)) && (mine.b.w & 0xffff) == 1 && (mine.b.w >> 31);
                const unsigned nb = ~__ballot_sync(RLGS_FULL, unit);
                const int up = nb ? __ffs(nb) - 1 : 32;     // leading unit jobs of the chunk
                if (up > 0) {
                    // ---- a run of unit jobs, one entry per lane (same effects as the ordinary loop below, in the same order)
                    const bool valid = lane < up;
                    Ent e = mine;
                    const int job = e.job();
                    const bool ended = valid && e.status() == L_RUNNING && st.t_prev + e.a.z - e.a.w == event_time;     // :198-204
                    const unsigned eb = __ballot_sync(RLGS_FULL, ended);
                    if (ended) {
                        D.planes[1][job] = event_time; D.planes[3][job] = e.b.y;
                        D.planes[4][job] = e.b.z & 0xffff; D.planes[5][job] = (e.b.z >> 16) & 0xffff;
                        D.planes[2][st.F + __popc(eb & ((1u << lane) - 1))] = job;
                    }
                    int64_t jct = ended ? event_time - D.trace[job].arrival_tick : 0;
#pragma unroll
                    for (int o = 16; o; o >>= 1) jct += __shfl_xor_sync(RLGS_FULL, jct, o);
                    st.sum_jct += jct; st.F += __popc(eb); n_events += __popc(eb);
                    const bool stay = valid && !ended;
                    if (stay) {
                        if (e.a.y & (1 << 23)) e.a.y &= ~(1 << 23);                                              // last_check_time == event_time
                        else if (e.status() == L_RUNNING) e.a.w += d; else e.b.y += d;                           // :216-230
                    }
                    const unsigned sb = __ballot_sync(RLGS_FULL, stay);
                    const bool ok = stay && unit_count + __popc(sb & ((1u << lane) - 1)) < unit_cap;               // :246 first fit in closed form
                    unit_count += __popc(__ballot_sync(RLGS_FULL, ok));
                    bool flipped = false;
                    if (ok) {
                        if (!e.started()) { e.set_started(); D.planes[0][job] = event_time; }                     // :251-252
                        if (e.status() == L_PENDING) { e.set_status(L_RUNNING); e.b.z += 1 << 16; flipped = true; }   // :265-267
                    } else if (stay && e.status() == L_RUNNING) {
                        e.set_status(L_PENDING); e.b.z = (e.b.z & ~0xffff) | ((e.b.z + 1) & 0xffff); flipped = true;  // :262-264
                    }
                    n_events += __popc(__ballot_sync(RLGS_FULL, flipped));
                    const bool running = stay && e.status() == L_RUNNING;
                    n_run += __popc(__ballot_sync(RLGS_FULL, running)); n_pend += __popc(__ballot_sync(RLGS_FULL, stay && !running));
                    int le = running ? event_time + e.a.z - e.a.w : RLGS_NEVER;                                   // :270-284
#pragma unroll
                    for (int o = 16; o; o >>= 1) le = min(le, __shfl_xor_sync(RLGS_FULL, le, o));
                    new_end = min(new_end, le);
                    if (stay) store_ent(buf + w_out + __popc(sb & ((1u << lane) - 1)), e);
                    w_out += __popc(sb);
                    __syncwarp();
                    if (up >= cnt) continue;
                    k0 = up;
                }
                materialise();      // another kind of job follows: continue on real node state
                pure = false;
            }
            bool keep = false;
            for (int k = k0; k < cnt; ++k) {
                Ent e; e.a = shfl_int4(mine.a, k); e.b = shfl_int4(mine.b, k);
                const int job = e.job();
                int status = e.status();
                if (fused) {
                    if (status == L_RUNNING && st.t_prev + e.a.z - e.a.w == event_time) {     // :198-204 end job
                        if (lane == 0) {
                            D.planes[1][job] = event_time; D.planes[3][job] = e.b.y;
                            D.planes[4][job] = e.b.z & 0xffff; D.planes[5][job] = (e.b.z >> 16) & 0xffff;
                            D.planes[2][st.F] = job;
                        }
                        st.sum_jct += event_time - D.trace[job].arrival_tick;
                        st.F += 1; n_events += 1;
                        continue;
                    }
                    if (e.a.y & (1 << 23)) e.a.y &= ~(1 << 23);                               // last_check_time == event_time: nothing to add
                    else if (status == L_RUNNING) e.a.w += d; else e.b.y += d;               // :216-230
                }
                JobRec jr; jr.a = make_int4(0, e.a.z, e.b.x, e.b.w); jr.b = make_int4(0, 0, 0, job);
                                PlaceResult pr = yarn_place(nv, c, jr, lane, D.place_scratch, 0, n_free_nodes, idle_unused);   // :246
                if (pr.ok) {
                    if (!e.started()) { e.set_started(); if (lane == 0) D.planes[0][job] = event_time; }    // :251-252
                    if (status == L_PENDING) { e.set_status(L_RUNNING); e.b.z += 1 << 16; n_events += 1; }   // :265-267
                } else if (status == L_RUNNING) {
                    e.set_status(L_PENDING); e.b.z = (e.b.z & ~0xffff) | ((e.b.z + 1) & 0xffff); n_events += 1;  // :262-264
                }
                if (e.status() == L_RUNNING) { n_run += 1; new_end = min(new_end, event_time + e.a.z - e.a.w); }  // :270-284
                else n_pend += 1;
                if (lane == k) { mine = e; keep = true; }
            }
            unsigned kb = __ballot_sync(RLGS_FULL, keep);
            if (keep) store_ent(buf + w_out + __popc(kb & ((1u << lane) - 1)), mine);
            w_out += __popc(kb);
            __syncwarp();
        }
        if (pure && unit_count > 0) materialise();   // the row statistics below read the node arrays
        st.sweep_jobs += w_out;  // runnable jobs swept at this event (after ends left and arrivals joined)
        st.M = w_out;
        if (st.M > st.max_m) st.max_m = st.M;
        st.events += n_events;
        st.next_end = new_end;
        st.t_prev = event_time;
        if (rows_on) {                                                                    // :287 LOG.checkpoint
            int idle = 0, full = 0, busy = 0;
            for (int b = 0; b < c.N; b += 32) {
                int i = b + lane;
                int pc = i < c.N ? __popc(nv.busy[i] & c.gmask) : -1;
                idle += __popc(__ballot_sync(RLGS_FULL, pc == 0));
                full += __popc(__ballot_sync(RLGS_FULL, pc == c.G));
                int s = pc > 0 ? pc : 0;
#pragma unroll
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(RLGS_FULL, s, o);
                busy += s;
            }
            if (lane == 0) {
                int4 *o = reinterpret_cast<int4 *>(row_ptr(rs, blockIdx.x, st.n_rows));
                o[0] = make_int4(idle, busy, n_run, n_pend);
                o[1] = make_int4(st.F, event_time, full, 0);
                o[2] = make_int4(0, 0, 0, 0); o[3] = make_int4(0, 0, 0, 0);
            }
        }
        st.n_rows += 1;
        __syncwarp();
    }
    if (st.done) flush_unfinished(D, buf, st.M, lane);
    if (lane == 0) {
        states[blockIdx.x] = st;
        if (st.done) returns[blockIdx.x] = -st.sum_jct;
    }
}
