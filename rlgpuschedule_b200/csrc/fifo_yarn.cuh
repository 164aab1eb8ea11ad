// fifo_yarn.cuh — the live tick loop of the reference (fifo schedule + yarn placement) as a
// persistent one-warp-per-replica kernel.
//
// Restates Scheduler.start() (core/scheduling/schedule.py:178-216) and what it drives each tick:
//   arrivals + front insertion   jobs_manager.py:228-241,115-140; job_queue_manager.py:147-154   (q1)
//   one scheduling attempt       schedule.py:40-60; algorithm.py:189-202                          (q5)
//   step (pending / processed)   jobs_manager.py:143-148,65-70; job.py:154-158,183-188
//   finish scan + release        jobs_manager.py:243-250; schedule.py:141-162; node.py:71-91
//   stats row                    schedule.py:95-133; jobs_manager.py:72-87
//
// Data movement per replica: job records (32 B) stream in from HBM in arrival order through a
// register ring (one coalesced 1 KB read per 32 jobs); the queue is a stack of the same records in
// global memory whose front entry lives in registers; node counters / busy masks and the
// running-job slots live in shared memory for the whole launch; one 64 B row per tick and
// start/end ticks per job stream out.  pending_time and time_processed are not stored: a queued
// job's pending time is d - arrival and a running job's processed time is d - start, so the
// per-tick "+= 1" sweeps of the reference reduce to the finish scan over the running slots.
#pragma once
#include "yarn_place.cuh"

struct SlotView {
    int4 *a;           // end tick (RLGS_NEVER = free) | start sequence | next slot in the calendar / free chain | job
    int4 *b;           // node | tasks<<16 (or 0xffff | nnodes<<16) | device mask (or first log entry) | util_mu_q|util_sd_q<<16 | -
    int64_t *memterm;
    int32_t *bkt;      // [RLGS_CAL_W] first slot of the jobs whose end tick == b (mod RLGS_CAL_W), -1 = empty
};

// Finish detection is a calendar: a started job is filed under its end tick modulo RLGS_CAL_W, so the
// per-tick finish scan of the reference (jobs_manager.py:243-250, every running job) touches only
// the bucket of the current tick.
#define RLGS_CAL_W 256

struct FifoSmem {
    NodeView nv;
    SlotView sv;
};

__host__ __device__ inline size_t fifo_smem_bytes(int N, int slot_cap) {
    size_t node_words = 3 * (size_t)N + (size_t)((N + 31) / 32);   // units, busy, key, ever bitmap
    node_words = (node_words + 3) & ~(size_t)3;  // 16-byte alignment of the int4 slot arrays
    return node_words * 4 + (size_t)slot_cap * (16 + 16 + 8) + RLGS_CAL_W * 4;
}

__device__ __forceinline__ FifoSmem fifo_carve(unsigned char *smem, int N, int slot_cap) {
    FifoSmem s;
    int32_t *w = reinterpret_cast<int32_t *>(smem);
    s.nv.units = w; w += N;
    s.nv.busy = reinterpret_cast<uint32_t *>(w); w += N;
    s.nv.ever = reinterpret_cast<uint32_t *>(w); w += (N + 31) / 32;
    s.nv.key = reinterpret_cast<uint32_t *>(w); w += N;
    while ((w - reinterpret_cast<int32_t *>(smem)) & 3) w += 1;
    s.sv.a = reinterpret_cast<int4 *>(w); w += 4 * slot_cap;
    s.sv.b = reinterpret_cast<int4 *>(w); w += 4 * slot_cap;
    s.sv.memterm = reinterpret_cast<int64_t *>(w); w += 2 * slot_cap;
    s.sv.bkt = w; w += RLGS_CAL_W;
    return s;
}

__device__ __forceinline__ JobRec load_rec(const rlgs_job *p) {
    JobRec r;
    const int4 *q = reinterpret_cast<const int4 *>(p);
    r.a = q[0]; r.b = q[1];
    return r;
}
__device__ __forceinline__ void store_rec(rlgs_job *p, const JobRec &r) {
    int4 *q = reinterpret_cast<int4 *>(p);
    q[0] = r.a; q[1] = r.b;
}
__device__ __forceinline__ JobRec shfl_rec(const JobRec &r, int src) {
    JobRec o; o.a = shfl_int4(r.a, src); o.b = shfl_int4(r.b, src);
    return o;
}

// Saves / restores the shared-memory state of a replica (chunked runs and env steps).  The node keys,
// the calendar buckets and the free-slot chain are derived data: they are rebuilt on restore.
__device__ __forceinline__ void fifo_state_io(const RepDesc &D, FifoSmem s, const ClusterConst &c, int slot_cap, RepState &st, int lane, bool save) {
    const int N = c.N;
    int nw = 2 * N + (N + 31) / 32;
    int32_t *sm = s.nv.units;  // units, busy, ever are contiguous
    for (int i = lane; i < nw; i += 32) {
        if (save) D.node_save[i] = sm[i]; else sm[i] = D.node_save[i];
    }
    const int hw = st.hw;
    for (int i = lane; i < hw; i += 32) {
        if (save) {
            int4 a = s.sv.a[i], b = s.sv.b[i];
            int64_t m = s.sv.memterm[i];
            D.slot_save[2 * i] = make_int4(a.x, a.w, b.x, b.y);
            D.slot_save[2 * i + 1] = make_int4(a.y, b.z, (int)(uint32_t)m, (int)(m >> 32));
        } else {
            int4 x = D.slot_save[2 * i], y = D.slot_save[2 * i + 1];
            s.sv.a[i] = make_int4(x.x, y.x, -1, x.y);
            s.sv.b[i] = make_int4(x.z, x.w, y.y, 0);
            s.sv.memterm[i] = (int64_t)(((uint64_t)(uint32_t)y.w << 32) | (uint32_t)y.z);
        }
    }
    __syncwarp();
    if (!save) {
        for (int i = lane; i < N; i += 32) s.nv.key[i] = node_key(s.nv.units[i], s.nv.busy[i], c);
        for (int i = lane; i < RLGS_CAL_W; i += 32) s.sv.bkt[i] = -1;
        __syncwarp();
        int free_head = -1;
        if (lane == 0) {
            for (int i = hw - 1; i >= 0; --i) {
                int e = s.sv.a[i].x;
                if (e == RLGS_NEVER) { s.sv.a[i].z = free_head; free_head = i; }
                else { int b = e & (RLGS_CAL_W - 1); s.sv.a[i].z = s.sv.bkt[b]; s.sv.bkt[b] = i; }
            }
        }
        st.free_hint = __shfl_sync(RLGS_FULL, free_head, 0);
        __syncwarp();
    }
}

// Releases the resources of the job in slot `sl` at tick d and records its completion.
__device__ __forceinline__ void fifo_finish_slot(const RepDesc &D, FifoSmem s, const ClusterConst &c, RepState &st, int sl, int lane) {
    const int4 sb = s.sv.b[sl];
    const int job = s.sv.a[sl].w;
    const uint32_t place = (uint32_t)sb.x, mask = (uint32_t)sb.y, util = (uint32_t)sb.z;
    const int64_t mterm = s.sv.memterm[sl];
    int ndev;
    if ((place & 0xffff) != 0xffff) {
        release_single(s.nv, c, (int)(place & 0xffff), (int)(place >> 16), mask, lane, st.n_free_nodes);
        ndev = __popc(mask);
    } else {
        int nn = (int)(place >> 16), off = (int)mask;
        ndev = 0;
        for (int b = 0; b < nn; b += 32) {
            bool act = b + lane < nn;
            int2 e = act ? D.place_log[off + b + lane] : make_int2(0, 0);
            release_entry(s.nv, c, act, e, st.n_free_nodes);
            int cnt = act ? __popc((uint32_t)e.y) : 0;
#pragma unroll
            for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(RLGS_FULL, cnt, o);
            ndev += cnt;
        }
    }
    __syncwarp();
    st.busy_gpus -= ndev;
    st.mem_sum -= mterm;   // mem_term already is the job's total over its devices
    int64_t mu = util & 0xffff, sd = util >> 16;
    st.util_mu_sum -= mu * ndev;
    st.util_var_sum -= sd * sd * ndev;
    if (lane == 0) {
        s.sv.a[sl] = make_int4(RLGS_NEVER, 0, st.free_hint, job);   // push on the free-slot chain
        D.end_tick[job] = st.d;
        D.finish_order[st.F] = job;
    }
    st.free_hint = sl;
    st.F += 1; st.R -= 1;
    st.head_blocked = 0;  // resources were freed: the queue head may fit now
    __syncwarp();
}

// Inputs / outputs of the vectorised RL environment (ENV instantiation of the kernel).  Build-defined
// semantics (the reference's model/env.py:1-6 is an empty stub): one step = one scheduler tick; the
// action picks which of the first `window_k` queued jobs gets this tick's placement attempt
// (cf. the k-job look-ahead window of schedule_horus, algorithm.py:204-240); -1 = no attempt.
struct EnvIO {
    const int32_t *actions;   // [replicas] action of this step (policy 2), device memory
    float *obs;               // [replicas][obs_dim]
    float *reward;            // [replicas]  -(queued + running) summed over the ticks of this launch
    uint8_t *done;            // [replicas]
    int32_t policy;           // 0 = queue head (fifo), 1 = random window (counter-based RNG), 2 = actions[]
    int32_t window_k;
    uint32_t seed;
    int32_t obs_dim;          // 3N + 4*window_k + 4
};

__device__ __forceinline__ uint32_t rlgs_hash3(uint32_t seed, uint32_t replica, uint32_t tick) {
    // splitmix64 finaliser of (seed, replica, tick); same function as oracle/cpu_sim.c
    uint64_t z = ((uint64_t)seed << 32) ^ ((uint64_t)replica * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)tick * 0xBF58476D1CE4E5B9ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

// Duration of a job that was just placed on `nodes_used` nodes (restates calculate_network_costs,
// core/network/network_service.py:3-39, + Job.add_network_costs job.py:196-197).  The reference's own
// path raises (Job.is_distributed reads a missing attribute, job.py:199-200), so the semantics are
// build-defined: distributed = more than one task; there are no PS tasks, so the symmetric difference
// of PS and worker nodes is the set of worker nodes.  Same float64 operation order as the Python.
__device__ __forceinline__ int netcost_dur_ticks(const RepDesc &D, const NetCost &net, int job, int tasks, int nodes_used, int lane) {
    const int J = D.J;
    double dur = D.net_in[job];
    if (tasks > 1) {
        double model_per_sec = D.net_in[J + job] / net.bandwidth;
        double nodes_induced_sec = (double)nodes_used * net.latency;
        double iteration_round_trip = D.net_in[2 * J + job] * 2.0;
        dur += (model_per_sec + nodes_induced_sec) * iteration_round_trip;
    }
    if (lane == 0) D.dur_out[job] = dur;
    double c = ceil(dur);
    return c < 1.0 ? 1 : (c > 1.0e9 ? 1000000000 : (int)c);
}

#ifndef RLGS_FIFO_MIN_BLOCKS
#define RLGS_FIFO_MIN_BLOCKS 18
#endif
template <bool ENV, bool ROWS, bool NET>
__global__ void __launch_bounds__(32, RLGS_FIFO_MIN_BLOCKS) fifo_yarn_kernel(const RepDesc *__restrict__ descs, RepState *__restrict__ states,
                                                       ClusterConst c, int slot_cap, int tick_budget, RowStore rs,
                                                       int64_t *__restrict__ returns, int64_t max_ticks, EnvIO env, NetCost net) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = lane_id();
    const RepDesc D = descs[blockIdx.x];
    RepState st = states[blockIdx.x];
    if (st.done || st.status != RLGS_OK) {
        if (ENV && lane == 0) { env.reward[blockIdx.x] = 0.f; env.done[blockIdx.x] = 1; }
        return;
    }
    FifoSmem s = fifo_carve(smem_raw, c.N, slot_cap);
    float reward_acc = 0.f;
    rlgs_row *row_cur = nullptr;   // next row of this replica inside the current chunk
    constexpr bool rows_mode = ROWS;   // compile-time: the row path disappears from the rows-off instantiations
    // device-resident chunk-major row store: a launch stops when the allocated chunks are full
    if (rows_mode && (int64_t)st.d + tick_budget > (int64_t)rs.n_chunks * RLGS_ROW_CHUNK)
        tick_budget = (int)max((int64_t)0, (int64_t)rs.n_chunks * RLGS_ROW_CHUNK - st.d);
    if (st.d == 0) {  // first launch of a run: empty cluster, no running jobs
        int nw = 2 * c.N + (c.N + 31) / 32;
        for (int i = lane; i < nw; i += 32) s.nv.units[i] = 0;
        const uint32_t empty_key = node_key(0, 0u, c);
        for (int i = lane; i < c.N; i += 32) s.nv.key[i] = empty_key;
        for (int i = lane; i < RLGS_CAL_W; i += 32) s.sv.bkt[i] = -1;
        __syncwarp();
    } else {
        fifo_state_io(D, s, c, slot_cap, st, lane, false);
    }
    const int J = D.J;

    // register ring over the trace: lane l holds job ring_base + l
    int ring_base = st.cursor & ~31;
    JobRec ring;
    {
        int idx = ring_base + lane;
        if (idx < J) ring = load_rec(D.trace + idx); else { ring.a = make_int4(RLGS_NEVER, 0, 0, 0); ring.b = make_int4(0, 0, 0, 0); }
    }
    // the two front entries of the queue live in registers
    JobRec h0;   // the queue front lives in registers; after a pop the next record is fetched while the tick finishes
    h0.a = h0.b = make_int4(0, 0, 0, 0);
    if (st.Q > 0) h0 = load_rec(D.stack + st.head);

    // the launch stops at its tick budget or at the safety limit max_ticks, whichever comes first (one compare per tick)
    const int d_budget = st.d + tick_budget;
    const int d_stop = (max_ticks > 0 && max_ticks < (long)d_budget) ? (int)max(max_ticks, (long)st.d) : d_budget;
    while (true) {
        if ((J - st.cursor) + st.R == 0) { st.done = 1; break; }   // schedule.py:185 (queue not consulted, q2)
        if (st.d == d_stop) { if (max_ticks > 0 && st.d >= max_ticks) { st.done = 1; st.status = RLGS_ERR_CAPACITY; } break; }
        const int d = st.d;

        // ---------------- arrivals: every job with arrival_tick <= d, pushed to the FRONT in order (q1)
        if (st.cursor < J) {
            int idx = ring_base + lane;
            bool arr = idx >= st.cursor && idx < J && ring.arrival() <= d;
            unsigned ab = __ballot_sync(RLGS_FULL, arr);
            if (ab) {
                int k = __popc(ab);
                bool ring_covers_batch = (st.cursor + k < ring_base + 32) || (ring_base + 32 >= J);
                if (ring_covers_batch) {
                    int first = st.cursor - ring_base;
                    if (arr) store_rec(D.stack + (st.head - k) + (idx - st.cursor), ring);
                    JobRec n0 = shfl_rec(ring, first);
                    h0 = n0;
                } else {
                    // batch runs past the ring: count it from global memory, then copy records
                    int pos = ring_base + 32;
                    while (pos < J) {
                        int i2 = pos + lane;
                        bool a2 = i2 < J && D.trace[i2].arrival_tick <= d;
                        int c2 = __popc(__ballot_sync(RLGS_FULL, a2));
                        k += c2;
                        if (c2 < 32) break;
                        pos += 32;
                    }
                    for (int b = 0; b < k; b += 32) {
                        int i2 = b + lane;
                        if (i2 < k) store_rec(D.stack + (st.head - k) + i2, load_rec(D.trace + st.cursor + i2));
                    }
                    __syncwarp();
                    h0 = load_rec(D.stack + st.head - k);
                }
                if (st.Q == 0) st.bottom_arr = d;
                st.head -= k; st.Q += k; st.cursor += k;
                st.sum_arr += (int64_t)k * d;
                st.head_blocked = 0;
                if (st.Q > st.max_q) st.max_q = st.Q;
                if (st.cursor >= ring_base + 32) {
                    ring_base = st.cursor & ~31;
                    int i3 = ring_base + lane;
                    if (i3 < J) ring = load_rec(D.trace + i3); else { ring.a = make_int4(RLGS_NEVER, 0, 0, 0); }
                }
                __syncwarp();
            }
        }

        // ---------------- one scheduling attempt on the queue head (schedule.py:188-190)
        if (!ENV && st.Q > 0 && st.n_free_nodes >= 1 && !st.head_blocked) {
            PlaceResult pr = yarn_place(s.nv, c, h0, lane, D.place_log, st.log_len, st.n_free_nodes, st.idle_nodes);
            if (pr.ok) {
                int job = h0.index();
                int ndev = h0.tasks() * h0.gpc();
                const int dur_ticks = NET ? netcost_dur_ticks(D, net, job, h0.tasks(), pr.nnodes, lane) : h0.dur();
                int sl = st.free_hint;                                    // pop the free-slot chain, else a fresh slot
                if (sl >= 0) st.free_hint = s.sv.a[sl].z; else sl = st.hw++;
                if (sl >= slot_cap) { st.status = RLGS_ERR_CAPACITY; st.done = 1; break; }
                const int cal = (d + dur_ticks) & (RLGS_CAL_W - 1);
                const int cal_head = s.sv.bkt[cal];
                __syncwarp();                                             // every lane has read the chain heads before lane 0 rewrites them
                if (lane == 0) {
                    s.sv.a[sl] = make_int4(d + dur_ticks, st.start_seq, cal_head, job);
                    s.sv.bkt[cal] = sl;                                   // file under the end tick
                    s.sv.b[sl] = make_int4(pr.node >= 0 ? (pr.node | (h0.tasks() << 16)) : (int)(0xffffu | ((uint32_t)pr.nnodes << 16)),
                                           pr.node >= 0 ? (int)pr.mask : st.log_len, (int)h0.util(), 0);
                    s.sv.memterm[sl] = h0.mem_term();
                    D.start_tick[job] = d;
                    D.place_off[job] = st.log_len;
                }
                st.log_len += pr.nnodes;
                st.start_seq += 1;
                st.busy_gpus += ndev;
                st.mem_sum += h0.mem_term();
                int64_t mu = h0.util() & 0xffff, sd = h0.util() >> 16;
                st.util_mu_sum += mu * ndev;
                st.util_var_sum += sd * sd * ndev;
                st.sum_arr -= h0.arrival();
                st.sum_jct += (int64_t)(d + dur_ticks - h0.arrival());  // end is fixed at start (no preemption)
                st.R += 1; st.Q -= 1; st.head += 1;
                if (st.R > st.max_r) st.max_r = st.R;
                if (st.Q > 0) h0 = load_rec(D.stack + st.head);   // consumed by the next tick's attempt
                __syncwarp();
            } else if (h0.fits()) {
                st.head_blocked = 1;  // a failed attempt has no side effect: skip retries until a release
            }
        }

        if (ENV && st.Q > 0) {
            // ---------------- environment: the policy picks a job inside the look-ahead window
            const int win = min(st.Q, env.window_k);
            int pick = 0;
            if (env.policy == 1) pick = (int)(rlgs_hash3(env.seed, (uint32_t)(rs.replica + blockIdx.x), (uint32_t)d) % (uint32_t)win);
            else if (env.policy == 2) pick = env.actions[blockIdx.x];
            if (pick >= 0 && pick < win && st.n_free_nodes >= 1) {
                JobRec hx = load_rec(D.stack + st.head + pick);
                PlaceResult pr = yarn_place(s.nv, c, hx, lane, D.place_log, st.log_len, st.n_free_nodes, st.idle_nodes);
                if (pr.ok) {
                    int job = hx.index();
                    int ndev = hx.tasks() * hx.gpc();
                    const int dur_ticks = NET ? netcost_dur_ticks(D, net, job, hx.tasks(), pr.nnodes, lane) : hx.dur();
                    int sl = st.free_hint;                                    // pop the free-slot chain, else a fresh slot
                    if (sl >= 0) st.free_hint = s.sv.a[sl].z; else sl = st.hw++;
                    if (sl >= slot_cap) { st.status = RLGS_ERR_CAPACITY; st.done = 1; break; }
                    const int cal = (d + dur_ticks) & (RLGS_CAL_W - 1);
                    const int cal_head = s.sv.bkt[cal];
                    __syncwarp();
                    if (lane == 0) {
                        s.sv.a[sl] = make_int4(d + dur_ticks, st.start_seq, cal_head, job);
                        s.sv.bkt[cal] = sl;                               // file under the end tick
                        s.sv.b[sl] = make_int4(pr.node >= 0 ? (pr.node | (hx.tasks() << 16)) : (int)(0xffffu | ((uint32_t)pr.nnodes << 16)),
                                               pr.node >= 0 ? (int)pr.mask : st.log_len, (int)hx.util(), 0);
                        s.sv.memterm[sl] = hx.mem_term();
                        D.start_tick[job] = d;
                        D.place_off[job] = st.log_len;
                    }
                    st.log_len += pr.nnodes;
                    st.start_seq += 1;
                    st.busy_gpus += ndev;
                    st.mem_sum += hx.mem_term();
                    int64_t mu = hx.util() & 0xffff, sd = hx.util() >> 16;
                    st.util_mu_sum += mu * ndev;
                    st.util_var_sum += sd * sd * ndev;
                    st.sum_arr -= hx.arrival();
                    st.sum_jct += (int64_t)(d + dur_ticks - hx.arrival());
                    // queue.pop(pick): entries in front of it move one place towards the back of the stack
                    JobRec mv; bool m = lane < pick;
                    if (m) mv = load_rec(D.stack + st.head + lane);
                    __syncwarp();
                    if (m) store_rec(D.stack + st.head + lane + 1, mv);
                    __syncwarp();
                    st.R += 1; st.Q -= 1; st.head += 1;
                        if (st.R > st.max_r) st.max_r = st.R;
                    if (st.Q > 0) st.bottom_arr = D.stack[st.head + st.Q - 1].arrival_tick;
                }
            }
        }

        // median loads are issued early; they are consumed when the row is written
        int med_lo_arr = 0, med_hi_arr = 0;
        if (rows_mode && st.Q > 0) {
            med_lo_arr = D.stack[st.head + (st.Q - 1) / 2].arrival_tick;
            med_hi_arr = D.stack[st.head + st.Q / 2].arrival_tick;
        }

        // ---------------- delta_time += 1; step; release finished jobs in start order
        st.d = d + 1;
        {
            // calendar bucket of this tick: jobs whose end == d finish, in start order when there are several
            // (running_jobs dict order, schedule.py:144); the others in the chain end a multiple of RLGS_CAL_W later
            const int bk = st.d & (RLGS_CAL_W - 1);
            for (int more = 1; more;) {
                int sl = s.sv.bkt[bk], prev = -1, best = -1, best_prev = -1, best_seq = RLGS_NEVER, best_next = -1, matches = 0;
                while (sl >= 0) {
                    const int4 e = s.sv.a[sl];                       // one 16-byte load per hop: end, seq, next
                    if (e.x == st.d) { matches++; if (e.y < best_seq) { best_seq = e.y; best = sl; best_prev = prev; best_next = e.z; } }
                    prev = sl; sl = e.z;
                }
                if (best < 0) break;
                more = matches > 1;                                  // walk again only if another job finishes this tick
                __syncwarp();                                        // the walk is over in every lane before lane 0 unlinks
                if (lane == 0) { if (best_prev < 0) s.sv.bkt[bk] = best_next; else s.sv.a[best_prev].z = best_next; }
                __syncwarp();
                fifo_finish_slot(D, s, c, st, best, lane);
            }
        }

        // ---------------- stats row (schedule.py:204-205)
        st.sumQ += st.Q; st.sumR += st.R;
        if (ENV) reward_acc -= (float)(st.Q + st.R);
        if (rows_mode && lane == 0) {
            if (((st.d - 1) & (RLGS_ROW_CHUNK - 1)) == 0 || row_cur == nullptr) row_cur = row_ptr(rs, blockIdx.x, st.d - 1);
            rlgs_row *row = row_cur++;
            int4 w0 = make_int4(st.idle_nodes, st.busy_gpus, st.R, st.Q);
            int4 w1 = make_int4(st.F, st.Q > 0 ? st.d - med_lo_arr : 0, st.Q > 0 ? st.d - med_hi_arr : 0,
                                st.Q > 0 ? st.d - st.bottom_arr : 0);
            int64_t sp = (int64_t)st.Q * st.d - st.sum_arr;
            int4 *o = reinterpret_cast<int4 *>(row);
            o[0] = w0; o[1] = w1;
            o[2] = make_int4((int)(uint32_t)sp, (int)(sp >> 32), (int)(uint32_t)st.mem_sum, (int)(st.mem_sum >> 32));
            o[3] = make_int4((int)(uint32_t)st.util_mu_sum, (int)(st.util_mu_sum >> 32), (int)(uint32_t)st.util_var_sum,
                             (int)(st.util_var_sum >> 32));
        }
    }

    st.events = (int64_t)st.cursor + st.start_seq + st.F;   // arrivals + starts + finishes
    if (!st.done && st.status == RLGS_OK && (J - st.cursor) + st.R == 0) st.done = 1;   // the while condition of schedule.py:185
    if (ENV) {
        // observation: per node free GPUs / cpu / mem, the look-ahead window, queue statistics
        float *o = env.obs + (size_t)blockIdx.x * env.obs_dim;
        for (int i = lane; i < c.N; i += 32) {
            o[i] = (float)__popc(~s.nv.busy[i] & c.gmask);
            o[c.N + i] = (float)(c.cpu_cap - RLGS_CPUS_PER_TASK * s.nv.units[i]);
            o[2 * c.N + i] = (float)(c.mem_cap - RLGS_MEM_PER_TASK * s.nv.units[i]);
        }
        for (int i = lane; i < env.window_k; i += 32) {
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < st.Q) { JobRec r = load_rec(D.stack + st.head + i); w = make_float4((float)r.gpus(), (float)r.tasks(), (float)r.dur(), (float)(st.d - r.arrival())); }
            float *ow = o + 3 * c.N + 4 * i;
            ow[0] = w.x; ow[1] = w.y; ow[2] = w.z; ow[3] = w.w;
        }
        if (lane == 0) {
            float *t = o + 3 * c.N + 4 * env.window_k;
            t[0] = (float)st.Q; t[1] = (float)st.R; t[2] = (float)st.F; t[3] = (float)st.d;
            env.reward[blockIdx.x] = reward_acc;
            env.done[blockIdx.x] = (uint8_t)(st.done != 0);
        }
    }
    fifo_state_io(D, s, c, slot_cap, st, lane, true);
    if (lane == 0) {
        states[blockIdx.x] = st;
        if (st.done) returns[blockIdx.x] = -st.sum_jct;  // episode return, read by the all-gather
    }
}
