// fifo_grp.cuh — the live tick loop of the reference (fifo schedule + yarn placement) as a persistent kernel in which
// a GROUP of LPR lanes (8, 16 or 32) advances one replica, so that one warp carries 32 / LPR replicas.
//
// Restates Scheduler.start() (core/scheduling/schedule.py:178-216) and what it drives each tick:
//   arrivals + front insertion   jobs_manager.py:228-241,115-140; job_queue_manager.py:147-154   (q1)
//   one scheduling attempt       schedule.py:40-60; algorithm.py:189-202                          (q5)
//   yarn fit                     algorithm.py:28-32,301-417; node.py:109-127,200-275; device.py:19-43,67-77 (q8, q11)
//   step (pending / processed)   jobs_manager.py:143-148,65-70; job.py:154-158,183-188
//   finish scan + release        jobs_manager.py:243-250; schedule.py:141-162; node.py:71-91
//   stats row                    schedule.py:95-133; jobs_manager.py:72-87
//
// Why groups: the tick is a chain of dependent scalar bookkeeping (ncu, round 1: issue-bound, every scalar instruction
// executed by 32 lanes for one replica).  With LPR = 8 every warp instruction serves four replicas; the only lane-parallel
// part, the first-fit scan over the node keys, keeps its width because a lane reads four consecutive 32-bit keys with one
// 16-byte shared-memory load (8 lanes x 4 keys = 32 nodes per step, as the 32-lane version had).  Groups synchronise among
// themselves only (__ballot_sync / __shfl_sync / __syncwarp with the group's lane mask): they follow their own control flow
// and reconverge wherever their paths coincide, which is most of the time because nearly every tick of a loaded trace has an
// arrival, a start and a finish.
//
// On chip per replica (shared memory, 16-byte aligned pieces):
//   key[N]    one word per node: idle devices, busy-device mask and free task units (see NodeC below)
//   ever[N/32] node ever used (q3)
//   slot[S]   int4 {end tick | RLGS_NEVER, next slot in the calendar / free chain, node | tasks << 16 (or 0xffff | nnodes << 16),
//                   device mask (or first placement-log entry)},  sjob[S] job index
//   bkt[128]  calendar: head | tail << 16 of the chain of running jobs whose end tick == b (mod 128); a started job is appended
//             at the tail, so a chain is in start order and same-tick finishes come out in the running_jobs dict order
// pending_time and time_processed are not stored: d - arrival and d - start.
#pragma once
#include "rlgs_device.cuh"

#ifndef RLGS_CAL_W
#define RLGS_CAL_W 128
#endif
#define RLGS_NONE16 0xffffu

// ---- group of LPR lanes --------------------------------------------------------------------------------------------
template <int LPR>
struct Grp {
    static constexpr int K = 32 / LPR;   // replicas per warp
    int g, gl, shift;
    unsigned mask;
    __device__ __forceinline__ Grp() {
        const int lane = threadIdx.x & 31;
        g = lane / LPR; gl = lane % LPR; shift = g * LPR;
        mask = LPR == 32 ? 0xffffffffu : (((1u << (LPR & 31)) - 1u) << shift);
    }
    __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(mask, p) >> shift; }   // group-relative bits
    __device__ __forceinline__ int shfl(int v, int src) const { return __shfl_sync(mask, v, src, LPR); }
    __device__ __forceinline__ int shfl_up(int v, int o) const { return __shfl_up_sync(mask, v, o, LPR); }
    __device__ __forceinline__ int shfl_xor(int v, int o) const { return __shfl_xor_sync(mask, v, o, LPR); }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ int incl_scan(int v) const {
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1) { int t = shfl_up(v, o); if (gl >= o) v += t; }
        return v;
    }
    __device__ __forceinline__ int sum(int v) const {
#pragma unroll
        for (int o = LPR / 2; o; o >>= 1) v += shfl_xor(v, o);
        return v;
    }
    __device__ __forceinline__ JobRec shfl_rec(const JobRec &r, int src) const {
        JobRec o;
        o.a.x = shfl(r.a.x, src); o.a.y = shfl(r.a.y, src); o.a.z = shfl(r.a.z, src); o.a.w = shfl(r.a.w, src);
        o.b.x = shfl(r.b.x, src); o.b.y = shfl(r.b.y, src); o.b.z = shfl(r.b.z, src); o.b.w = shfl(r.b.w, src);
        return o;
    }
};

// ---- shared-memory view of one replica -------------------------------------------------------------------------------
// Node word ("key"), the only per-node state the tick touches on its hot path.  Two layouts, chosen at compile time:
//   PK  (nodes with <= 8 GPUs, the reference's default):  idle devices << 25 | busy-device mask << 16 | free task units
//   !PK (9 .. 32 GPUs per node):                          idle devices << 16 | free task units,  busy mask in its own array
// Free task units = min(cpu_cap // 12, mem_cap // 60) - tasks charged (placed + leaked, q8) <= 0x7fff (rlgs_create refuses larger
// nodes).  A task is only ever charged to a node whose key showed room for it (placement and the q8 leak alike), so the count
// never goes negative and nothing is clamped; cpu_free and mem_free follow from it (every task charges exactly 12 / 60).
// In both layouts bit 15 and bit 31 are clear and the high half compares like the idle count, which lets ONE subtraction test
// "idle >= gpus and free units >= tasks" for a node (grp_fit4).
struct GrpSm {
    uint32_t *key;    // [Npad] zero beyond N (a zero key never fits)
    uint32_t *busy;   // [N] (!PK only)
    uint32_t *ever;   // [ceil(N/32)]
    int4 *slot;       // [slot_cap]
    int32_t *sjob;    // [slot_cap]
    uint32_t *bkt;    // [RLGS_CAL_W]
};

__host__ __device__ inline int grp_npad(int N, int lpr) { int q = 4 * lpr; return (N + q - 1) / q * q; }
__host__ __device__ inline bool grp_packed(int G) { return G <= 8; }

__host__ __device__ inline size_t grp_smem_bytes(int N, int G, int slot_cap, int lpr) {
    size_t words = (size_t)grp_npad(N, lpr) + (grp_packed(G) ? 0 : (size_t)N) + (size_t)((N + 31) / 32);
    words = (words + 3) & ~(size_t)3;
    words += 4 * (size_t)slot_cap + (size_t)slot_cap + RLGS_CAL_W;
    return ((words + 3) & ~(size_t)3) * 4;
}

__device__ __forceinline__ GrpSm grp_carve(unsigned char *base, int N, int G, int slot_cap, int lpr) {
    GrpSm s;
    uint32_t *w = reinterpret_cast<uint32_t *>(base);
    s.key = w; w += grp_npad(N, lpr);
    s.busy = w; if (!grp_packed(G)) w += N;
    s.ever = w; w += (N + 31) / 32;
    while ((w - reinterpret_cast<uint32_t *>(base)) & 3) w += 1;
    s.slot = reinterpret_cast<int4 *>(w); w += 4 * slot_cap;
    s.sjob = reinterpret_cast<int32_t *>(w); w += slot_cap;
    s.bkt = w;
    return s;
}

template <bool PK>
struct NodeC {
    static constexpr int IDLE_SHIFT = PK ? 25 : 16;
    __device__ static __forceinline__ uint32_t need_word(int gpus, int tasks) { return ((uint32_t)gpus << IDLE_SHIFT) | (uint32_t)tasks; }
    __device__ static __forceinline__ int idle(uint32_t k) { return (int)(k >> IDLE_SHIFT); }
    __device__ static __forceinline__ int free_units(uint32_t k) { return (int)(k & 0xffffu); }
    __device__ static __forceinline__ uint32_t busy(const GrpSm &s, int i, uint32_t k) { return PK ? ((k >> 16) & 0x1ffu) : s.busy[i]; }
    __device__ static __forceinline__ uint32_t make(int fu, uint32_t busy, const ClusterConst &c) {
        const uint32_t idle = (uint32_t)__popc(~busy & c.gmask);
        return PK ? ((idle << 25) | (busy << 16) | (uint32_t)fu) : ((idle << 16) | (uint32_t)fu);
    }
    __device__ static __forceinline__ void store(const GrpSm &s, int i, int fu, uint32_t busy, const ClusterConst &c) {
        if (!PK) s.busy[i] = busy;
        s.key[i] = make(fu, busy, c);
    }
    __device__ static __forceinline__ int cap(uint32_t k, int gpc) {   // Node.can_fit_num_task (node.py:109-127): tasks this node can take
        return min(idle(k) / gpc, free_units(k));
    }
};

// first of four consecutive node keys with idle devices >= gpus and free units >= tasks (need = NodeC::need_word), else -1
// (algorithm.py:407-409: free devices >= gpus, cpu_free >= 12 T, mem_free >= 60 T).  (k | H) - need keeps bit 15 / bit 31
// exactly when the low / high half did not borrow.
__device__ __forceinline__ int grp_fit4(uint4 k, uint32_t need) {
    const uint32_t H = 0x80008000u;
    const bool o0 = (((k.x | H) - need) & H) == H, o1 = (((k.y | H) - need) & H) == H;
    const bool o2 = (((k.z | H) - need) & H) == H, o3 = (((k.w | H) - need) & H) == H;
    return o0 ? 0 : (o1 ? 1 : (o2 ? 2 : (o3 ? 3 : -1)));
}

struct GrpPlace {
    int ok;          // 1 placed, 0 not placed
    int node;        // single-node: node index; multi-node: -1
    uint32_t mask;   // single-node: device mask
    int nnodes;      // entries appended to the placement log
};

// The reference gates a scheduling attempt on num_free_nodes() >= 1 (schedule.py:40-44; a node is free while cpu_free > 0 or
// mem_free > 0).  The gate cannot change an outcome: when it is closed every node has used up cpu and mem, so every free-unit
// count is 0 and any attempt fails without a side effect (the q8 leak needs free units too).  The tick loop therefore does not
// keep that count.

// Charges k tasks of cpu / mem to `node` (all lanes of the group call with the same node): the q8 leak.
template <int LPR>
__device__ __forceinline__ void grp_charge(const Grp<LPR> &G, GrpSm s, int node, int k) {
    const uint32_t key = s.key[node];
    G.sync();
    s.key[node] = key - (uint32_t)k;      // free units live in the low half and k never exceeds them
}

// The rare, group-local placements (ms_yarn_placement, algorithm.py:28-32): the q8 leak of a task no device accepts and
// try_cross_node_alloc_ms for jobs wider than a node.  Every lane of the group calls; the result is group-uniform.
template <int LPR, bool PK>
__device__ __forceinline__ GrpPlace grp_place_rare(const Grp<LPR> &G, GrpSm s, const ClusterConst &c, const JobRec &j, int2 *place_log,
                                                   int log_pos, int &idle_nodes) {
    typedef NodeC<PK> NC;
    GrpPlace r; r.ok = 0; r.node = -1; r.mask = 0; r.nnodes = 0;
    const int T = j.tasks(), gpc = j.gpc(), need_g = j.gpus();
    const bool fits = j.fits();
    const int Npad = grp_npad(c.N, LPR);
    if (need_g <= c.G) {
        // no device accepts the task (device.py:67-77): every candidate node of try_single_node_alloc_ms charges T tasks of
        // cpu / mem and keeps them (q8)
        for (int i = 0; i < c.N; ++i) {
            const uint32_t k = s.key[i];
            if (NC::idle(k) >= need_g && NC::free_units(k) >= T) grp_charge(G, s, i, T);
        }
        return r;
    }
    // ---- try_cross_node_alloc_ms (algorithm.py:301-393): walk the nodes in id order, each takes min(capacity, remaining) tasks;
    //      the fit+1 attempt of q11 has no side effect
    if (!fits) {   // the first task attempt on every node with capacity >= 1 leaks one task of cpu / mem (q8)
        for (int i = 0; i < c.N; ++i)
            if (NC::cap(s.key[i], gpc) >= 1) grp_charge(G, s, i, 1);
        return r;
    }
    int remaining = T, nodes_assigned = 0;
    for (int base = 0; base < Npad && remaining > 0; base += 4 * LPR) {
        const uint4 k = *reinterpret_cast<const uint4 *>(s.key + base + 4 * G.gl);
        const int c0 = NC::cap(k.x, gpc), c1 = NC::cap(k.y, gpc), c2 = NC::cap(k.z, gpc), c3 = NC::cap(k.w, gpc);
        const int sum4 = c0 + c1 + c2 + c3;
        const int incl = G.incl_scan(sum4);
        const int tot = G.shfl(incl, LPR - 1);
        int before = incl - sum4, cnt = 0;
        cnt += min(c0, remaining - before) > 0; before += c0;
        cnt += min(c1, remaining - before) > 0; before += c1;
        cnt += min(c2, remaining - before) > 0; before += c2;
        cnt += min(c3, remaining - before) > 0;
        nodes_assigned += G.sum(cnt);
        remaining -= min(remaining, tot);
    }
    if (remaining > 0 || nodes_assigned < j.least()) return r;   // rollback: no net effect (algorithm.py:376-386)
    // commit: recompute the same takes and apply them (each lane owns its four nodes: distinct addresses)
    remaining = T;
    int written = 0;
    G.sync();
    for (int base = 0; base < Npad && remaining > 0; base += 4 * LPR) {
        const int i0 = base + 4 * G.gl;
        const uint4 k = *reinterpret_cast<const uint4 *>(s.key + i0);
        const int cap[4] = {NC::cap(k.x, gpc), NC::cap(k.y, gpc), NC::cap(k.z, gpc), NC::cap(k.w, gpc)};
        const int sum4 = cap[0] + cap[1] + cap[2] + cap[3];
        const int incl = G.incl_scan(sum4);
        const int tot = G.shfl(incl, LPR - 1);
        int before = incl - sum4;
        int take[4], cnt = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) { take[q] = max(0, min(cap[q], remaining - before)); before += cap[q]; cnt += take[q] > 0; }
        const int cincl = G.incl_scan(cnt);
        int pos = log_pos + written + cincl - cnt, didle = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (take[q] > 0) {
                const int i = i0 + q;
                const uint32_t kq = q == 0 ? k.x : (q == 1 ? k.y : (q == 2 ? k.z : k.w));
                const uint32_t busy = NC::busy(s, i, kq);
                const uint32_t taken = lowest_bits(~busy & c.gmask, take[q] * gpc);
                NC::store(s, i, NC::free_units(kq) - take[q], busy | taken, c);
                const uint32_t bit = 1u << (i & 31);
                if (!(atomicOr(&s.ever[i >> 5], bit) & bit)) didle += 1;   // several lanes share an `ever` word
                place_log[pos++] = make_int2(i | (take[q] << 16), (int)taken);
            }
        }
        idle_nodes -= G.sum(didle);
        written += G.shfl(cincl, LPR - 1);
        remaining -= min(remaining, tot);
    }
    G.sync();
    r.ok = 1; r.nnodes = written;
    return r;
}

// Node.release_allocated_resources (node.py:71-91) for a multi-node job: one placement-log entry per lane, distinct nodes
template <int LPR, bool PK>
__device__ __forceinline__ int grp_release_multi(const Grp<LPR> &G, GrpSm s, const ClusterConst &c, const int2 *log, int nn) {
    typedef NodeC<PK> NC;
    int ndev = 0;
    G.sync();
    for (int b = G.gl; b < nn; b += LPR) {
        const int2 e = log[b];
        const int node = e.x & 0xffff, tasks = (e.x >> 16) & 0xffff;
        const uint32_t k = s.key[node];
        NC::store(s, node, NC::free_units(k) + tasks, NC::busy(s, node, k) & ~(uint32_t)e.y, c);
        ndev += __popc((uint32_t)e.y);
    }
    ndev = G.sum(ndev);
    G.sync();
    return ndev;
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
    int32_t obs_dim;          // 3N + 5*window_k + 4
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
__device__ __forceinline__ int netcost_dur_ticks(const RepDesc &D, const NetCost &net, int job, int tasks, int nodes_used, bool writer) {
    const int J = D.J;
    double dur = D.net_in[job];
    if (tasks > 1) {
        double model_per_sec = D.net_in[J + job] / net.bandwidth;
        double nodes_induced_sec = (double)nodes_used * net.latency;
        double iteration_round_trip = D.net_in[2 * J + job] * 2.0;
        dur += (model_per_sec + nodes_induced_sec) * iteration_round_trip;
    }
    if (writer) D.dur_out[job] = dur;
    double c = ceil(dur);
    return c < 1.0 ? 1 : (c > 1.0e9 ? 1000000000 : (int)c);
}

// Saves / restores the shared-memory state of a replica (bounded launches and env steps).  The node keys, the calendar and
// the free-slot chain are derived data: they are rebuilt on restore (calendar chains in start order, from start_tick[job]).
template <int LPR, bool PK>
__device__ __forceinline__ void grp_state_io(const Grp<LPR> &G, const RepDesc &D, GrpSm s, const ClusterConst &c, int slot_cap, RepState &st, bool save) {
    typedef NodeC<PK> NC;
    const int N = c.N;
    if (save)
        for (int i = G.gl; i < N; i += LPR) {
            const uint32_t k = s.key[i];
            D.node_save[i] = c.base_units - NC::free_units(k); D.node_save[N + i] = (int32_t)NC::busy(s, i, k);   // tasks charged, busy mask
        }
    for (int i = G.gl; i < (N + 31) / 32; i += LPR) {
        if (save) D.node_save[2 * N + i] = (int32_t)s.ever[i]; else s.ever[i] = (uint32_t)D.node_save[2 * N + i];
    }
    const int hw = st.hw;
    for (int i = G.gl; i < hw; i += LPR) {
        if (save) { int4 e = s.slot[i]; D.slot_save[i] = make_int4(e.x, s.sjob[i], e.z, e.w); }
        else { int4 x = D.slot_save[i]; s.slot[i] = make_int4(x.x, (int)RLGS_NONE16, x.z, x.w); s.sjob[i] = x.y; }
    }
    G.sync();
    if (!save) {
        const int Npad = grp_npad(N, LPR);
        for (int i = G.gl; i < Npad; i += LPR) {
            if (i < N) NC::store(s, i, c.base_units - D.node_save[i], (uint32_t)D.node_save[N + i], c); else s.key[i] = 0u;
        }
        for (int i = G.gl; i < RLGS_CAL_W; i += LPR) s.bkt[i] = RLGS_NONE16 | (RLGS_NONE16 << 16);
        G.sync();
        int free_head = -1;
        if (G.gl == 0) {
            for (int i = hw - 1; i >= 0; --i) {
                const int4 e = s.slot[i];
                if (e.x == RLGS_NEVER) { s.slot[i].y = free_head < 0 ? (int)RLGS_NONE16 : free_head; free_head = i; continue; }
                // sorted insert by start tick: chains are short and restores are rare
                const int b = e.x & (RLGS_CAL_W - 1), my_start = D.start_tick[s.sjob[i]];
                uint32_t hb = s.bkt[b];
                int cur = (int)(hb & 0xffff), prev = (int)RLGS_NONE16;
                while (cur != (int)RLGS_NONE16 && D.start_tick[s.sjob[cur]] < my_start) { prev = cur; cur = s.slot[cur].y; }
                s.slot[i].y = cur;
                uint32_t head = hb & 0xffff, tail = hb >> 16;
                if (prev == (int)RLGS_NONE16) head = (uint32_t)i; else s.slot[prev].y = i;
                if (cur == (int)RLGS_NONE16) tail = (uint32_t)i;
                s.bkt[b] = head | (tail << 16);
            }
        }
        st.free_hint = G.shfl(free_head, 0);
        G.sync();
    }
}

// 16-byte wire row (rlgs_row16, include/rlgs.h): the per-tick state that is not an integral of the start / finish event stream
__device__ __forceinline__ int4 pack_row16(int idle_nodes, int finished, int queued, int maxp, int med_lo, int med_hi) {
    const uint32_t w0 = (uint32_t)idle_nodes | ((uint32_t)finished << 12);
    const uint32_t w1 = (uint32_t)queued | ((uint32_t)(maxp & 0xfff) << 20);
    const uint32_t w2 = (uint32_t)(maxp >> 12) | ((uint32_t)(med_lo & 0xfffff) << 12);
    const uint32_t w3 = (uint32_t)(med_lo >> 20) | ((uint32_t)med_hi << 4);
    return make_int4((int)w0, (int)w1, (int)w2, (int)w3);
}

// address of row i of replica `rep` in the chunk-major row store; called once per 4096 rows, kept out of line so that the tick
// pays one decrement + branch for it
template <int ROW_BYTES>
__device__ __noinline__ unsigned char *row_address(const RowStore &rs, int rep, int64_t i) {
    return reinterpret_cast<unsigned char *>(rs.chunks[i >> RLGS_ROW_CHUNK_LOG]) +
           (((size_t)(rs.replica + rep) << RLGS_ROW_CHUNK_LOG) + (size_t)(i & (RLGS_ROW_CHUNK - 1))) * ROW_BYTES;
}

#ifndef RLGS_GRP_MIN_BLOCKS
#define RLGS_GRP_MIN_BLOCKS 16
#endif

// ---- the tick loop, in LOCKSTEP over the groups of a warp --------------------------------------------------------------
// Control flow is warp-uniform: a phase runs when ANY group of the warp needs it (__any_sync over the full warp) and every
// lane guards its effects with its own group's predicate.  All hot-path collectives therefore use the full-warp mask with
// width-LPR segments, every warp instruction serves all 32 / LPR replicas, and nothing depends on where divergent groups
// would reconverge (measured with per-group masks and free control flow: 9.4 active threads per instruction at LPR = 8,
// i.e. the groups simply took turns).  Rare paths — a job wider than a node, the q8 leak, an arrival batch longer than the
// register ring, the environment's queue.pop(pick) — stay group-local inside per-group branches.
#define RLGS_FULLMASK 0xffffffffu
// ROWS: 0 = no rows, 1 = 64-byte rlgs_row per tick, 2 = 16-byte rlgs_row16 per tick, 3 = 12-byte rlgs_row12 per tick,
//       4 = 16-byte rlgs_row16e per tick (row12 + the tick's start event)
template <int LPR, bool PK, bool ENV, int ROWS, bool NET>
__global__ void __launch_bounds__(32, RLGS_GRP_MIN_BLOCKS) fifo_grp_kernel(const RepDesc *__restrict__ descs, RepState *__restrict__ states, int n_rep,
                                                                           ClusterConst c, int slot_cap, int tick_budget, RowStore rs,
                                                                           int64_t *__restrict__ returns, int64_t max_ticks, EnvIO env, NetCost net) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef NodeC<PK> NC;
    const Grp<LPR> G;
    constexpr unsigned GBITS = LPR == 32 ? 0xffffffffu : ((1u << (LPR & 31)) - 1u);
#define GBALLOT(p) ((__ballot_sync(RLGS_FULLMASK, (p)) >> G.shift) & GBITS)
#define GSHFL(v, src) __shfl_sync(RLGS_FULLMASK, (v), (src), LPR)
    // the groups of a warp take replicas a quarter (half) of the launch apart: traces are attached to contiguous replica
    // ranges, so the replicas sharing a warp usually follow different traces
    const int rep_raw = G.g * (int)gridDim.x + (int)blockIdx.x;
    const bool valid = rep_raw < n_rep;
    const int rep = valid ? rep_raw : 0;          // lanes of an empty group keep executing (collectives are warp-wide) but touch nothing
    const RepDesc D = descs[rep];
    RepState st = states[rep];
    bool act = valid && !st.done && st.status == RLGS_OK;   // this group's replica still ticks in this launch
    const bool resident = act;                              // ... and its state has to be written back at the end
    const bool writer = G.gl == 0;
    if (ENV && valid && !act && writer) { env.reward[rep] = 0.f; env.done[rep] = 1; }
    const GrpSm s = grp_carve(smem_raw + (size_t)G.g * grp_smem_bytes(c.N, c.G, slot_cap, LPR), c.N, c.G, slot_cap, LPR);
    float reward_acc = 0.f;
    constexpr int ROW_BYTES = (ROWS == 2 || ROWS == 4) ? 16 : (ROWS == 3 ? 12 : (ROWS == 5 ? 4 : 64));
    unsigned char *row_cur = nullptr;   // next row of this replica inside the current chunk
    int rows_left = 0;                  // rows that still fit the chunk behind row_cur
    if (ROWS && (int64_t)st.d + tick_budget > (int64_t)rs.n_chunks * RLGS_ROW_CHUNK)   // the launch stops when the allocated chunks are full
        tick_budget = (int)max((int64_t)0, (int64_t)rs.n_chunks * RLGS_ROW_CHUNK - st.d);
    const int Npad = grp_npad(c.N, LPR);
    if (act) {
        if (st.d == 0) {   // first launch of a run: empty cluster, no running jobs
            for (int i = G.gl; i < Npad; i += LPR) {
                if (i < c.N) NC::store(s, i, c.base_units, 0u, c); else s.key[i] = 0u;
            }
            for (int i = G.gl; i < (c.N + 31) / 32; i += LPR) s.ever[i] = 0u;
            for (int i = G.gl; i < RLGS_CAL_W; i += LPR) s.bkt[i] = RLGS_NONE16 | (RLGS_NONE16 << 16);
        } else {
            grp_state_io<LPR, PK>(G, D, s, c, slot_cap, st, false);
        }
    }
    __syncwarp();
    const int J = D.J;

    // register ring over the trace: lane l of the group holds job ring_base + l (one coalesced read per LPR arrivals)
    int ring_base = st.cursor - (st.cursor % LPR);
    JobRec ring;
    ring.a = make_int4(RLGS_NEVER, 0, 0, 0); ring.b = make_int4(0, 0, 0, 0);
    if (act && ring_base + G.gl < J) ring = load_rec(D.trace + ring_base + G.gl);
    // arrival tick of the first job behind the ring: tells whether a batch that fills the ring ends there.  The validity flag
    // is kept apart so that the (predicated) load is not followed by a select on its result — that was a full load wait.
    int ring_next = RLGS_NEVER;
    bool ring_next_ok = act && ring_base + LPR < J;
    if (ring_next_ok) ring_next = D.trace[ring_base + LPR].arrival_tick;
    // The queue front lives in registers.  After a pop the next record is fetched into h0n and only becomes h0 at the next
    // attempt: the load then has the rest of the tick to complete.  (Loading straight into h0 made the compiler copy a loaded
    // word right behind the load — a full global-load wait at every start, 10 % of all stall samples in the v9 capture.)
    JobRec h0, h0n;
    h0.a = h0.b = h0n.a = h0n.b = make_int4(0, 0, 0, 0);
    bool h0_pending = false;
    if (act && st.Q > 0) h0 = load_rec(D.stack + st.head);

    // the launch stops at its tick budget or at the safety limit max_ticks, whichever comes first (one compare per tick)
    const int d_budget = st.d + tick_budget;
    const int d_stop = (max_ticks > 0 && max_ticks < (long)d_budget) ? (int)max(max_ticks, (long)st.d) : d_budget;
    while (true) {
        if (act) {
            if ((J - st.cursor) + st.R == 0) { st.done = 1; act = false; }   // schedule.py:185 (queue not consulted, q2)
            else if (st.d == d_stop) { if (max_ticks > 0 && st.d >= max_ticks) { st.done = 1; st.status = RLGS_ERR_CAPACITY; } act = false; }
        }
        if (!__any_sync(RLGS_FULLMASK, act)) break;
        const int d = st.d;

        // ---------------- arrivals: every job with arrival_tick <= d, pushed to the FRONT in order (q1)
        {
            const int idx = ring_base + G.gl;
            const bool arr = act && idx >= st.cursor && ring.arrival() <= d;   // entries beyond J carry RLGS_NEVER
            const unsigned ab = GBALLOT(arr);
            if (__any_sync(RLGS_FULLMASK, ab != 0u)) {
                int k = __popc(ab);
                const bool covered = (st.cursor + k < ring_base + LPR) || !ring_next_ok || ring_next > d;
                const int first = (st.cursor - ring_base) & (LPR - 1);
                JobRec n0;
                n0.a.x = GSHFL(ring.a.x, first); n0.a.y = GSHFL(ring.a.y, first); n0.a.z = GSHFL(ring.a.z, first); n0.a.w = GSHFL(ring.a.w, first);
                n0.b.x = GSHFL(ring.b.x, first); n0.b.y = GSHFL(ring.b.y, first); n0.b.z = GSHFL(ring.b.z, first); n0.b.w = GSHFL(ring.b.w, first);
                if (ab) {
                    if (covered) {
                        if (arr) store_rec(D.stack + (st.head - k) + (idx - st.cursor), ring);
                    } else {
                        // the batch runs past the ring (rare): count it from global memory, then copy the records; group-local
                        int pos = ring_base + LPR;
                        while (pos < J) {
                            const int i2 = pos + G.gl;
                            const int c2 = __popc(G.ballot(i2 < J && D.trace[i2].arrival_tick <= d));
                            k += c2;
                            if (c2 < LPR) break;
                            pos += LPR;
                        }
                        for (int b = G.gl; b < k; b += LPR) store_rec(D.stack + (st.head - k) + b, load_rec(D.trace + st.cursor + b));
                    }
                    h0 = n0;   // the first arrived job = trace[cursor], which the ring holds in lane `first`
                    h0_pending = false;
                    if (st.Q == 0) st.bottom_arr = d;
                    st.head -= k; st.Q += k; st.cursor += k;
                    if (ROWS == 1) st.sum_arr += (int64_t)k * d;
                    st.head_blocked = 0;
                    if (st.Q > st.max_q) st.max_q = st.Q;
                    if (st.cursor >= ring_base + LPR) {
                        ring_base = st.cursor - (st.cursor % LPR);
                        const int i3 = ring_base + G.gl;
                        if (i3 < J) ring = load_rec(D.trace + i3); else ring.a = make_int4(RLGS_NEVER, 0, 0, 0);
                        ring_next_ok = ring_base + LPR < J;
                        if (ring_next_ok) ring_next = D.trace[ring_base + LPR].arrival_tick;
                    }
                }
                __syncwarp();   // the stack records are visible to the whole group (median loads, env window, next head)
            }
        }

        // ---------------- one scheduling attempt (schedule.py:188-190): the queue head, or the policy's pick inside the window
        int pick = 0;
        int started = 0;   // 1 + the job started at this tick (rlgs_row16e)
        bool attempt;
        if (!ENV) attempt = act && st.Q > 0 && !st.head_blocked;
        else {
            const int win = min(st.Q, env.window_k);
            if (env.policy == 1 && win > 0) pick = (int)(rlgs_hash3(env.seed, (uint32_t)(rs.replica + rep), (uint32_t)d) % (uint32_t)win);
            else if (env.policy == 2 && valid) pick = env.actions[rep];
            attempt = act && st.Q > 0 && pick >= 0 && pick < win;
        }
        if (__any_sync(RLGS_FULLMASK, attempt)) {
            if (h0_pending) { h0 = h0n; h0_pending = false; }
            JobRec hx = h0;
            if (ENV && attempt && pick > 0) hx = load_rec(D.stack + st.head + pick);
            const int T = hx.tasks(), gpc = hx.gpc(), need_g = hx.gpus();
            const bool one_node = attempt && need_g <= c.G;
            GrpPlace pr; pr.ok = 0; pr.node = -1; pr.mask = 0; pr.nnodes = 0;
            // ---- try_single_node_alloc_ms (algorithm.py:396-417): first node in id order with enough idle devices, cpu and mem
            {
                const uint32_t need = NC::need_word(need_g, T);
                bool searching = one_node && hx.fits();
                int node = -1;
                for (int base = 0; base < Npad && __any_sync(RLGS_FULLMASK, searching); base += 4 * LPR) {
                    const uint4 k4 = *reinterpret_cast<const uint4 *>(s.key + base + 4 * G.gl);
                    const int f = searching ? grp_fit4(k4, need) : -1;
                    const unsigned b = GBALLOT(f >= 0);
                    const int l = b ? __ffs(b) - 1 : 0;
                    const int fl = GSHFL(f, l);
                    if (searching && b) { node = base + 4 * l + fl; searching = false; }
                }
                const bool hit = node >= 0;
                if (__any_sync(RLGS_FULLMASK, hit)) {
                    const int ni = hit ? node : 0;
                    const uint32_t kn = s.key[ni];
                    const uint32_t busy = NC::busy(s, ni, kn);
                    const uint32_t ew = s.ever[ni >> 5], bit = 1u << (ni & 31);
                    // devices in id order (node.py:209-219): with a lane per device, lane i of the group decides bit i (no loop);
                    // groups narrower than a node fall back to the bit-clearing loop
                    const uint32_t fm = ~busy & c.gmask;
                    const int want = hit ? T * gpc : 0;
                    uint32_t taken;
                    if (PK || LPR == 32) taken = GBALLOT(((fm >> G.gl) & 1u) && __popc(fm & ((1u << G.gl) - 1u)) < want);
                    else taken = lowest_bits(fm, want);
                    __syncwarp();                               // every lane has read the old node state
                    if (hit) {
                        if (!(ew & bit)) { st.idle_nodes--; s.ever[ni >> 5] = ew | bit; }
                        NC::store(s, ni, NC::free_units(kn) - T, busy | taken, c);
                        if (writer) D.place_log[st.log_len] = make_int2(ni | (T << 16), (int)taken);
                        pr.ok = 1; pr.node = ni; pr.mask = taken; pr.nnodes = 1;
                    }
                }
            }
            // ---- rare, group-local: the q8 leak of a task no device accepts, and jobs wider than a node (algorithm.py:301-393)
            if (attempt && (!one_node || !hx.fits()))
                pr = grp_place_rare<LPR, PK>(G, s, c, hx, D.place_log, st.log_len, st.idle_nodes);
            bool ok = pr.ok != 0;
            if (__any_sync(RLGS_FULLMASK, ok)) {
                const int job = hx.index();
                int dur_ticks = hx.dur();
                if (NET && ok) dur_ticks = netcost_dur_ticks(D, net, job, T, pr.nnodes, writer);
                int sl = st.free_hint, free_next = -1;                    // pop the free-slot chain, else a fresh slot
                if (ok) {
                    if (sl >= 0) { const int nx = s.slot[sl].y; free_next = nx == (int)RLGS_NONE16 ? -1 : nx; }
                    else sl = st.hw;
                    if (sl >= slot_cap) { st.status = RLGS_ERR_SLOTS; st.done = 1; act = false; ok = false; }
                }
                const int end = d + dur_ticks, cal = end & (RLGS_CAL_W - 1);
                const uint32_t hb = s.bkt[cal];
                const uint32_t tail = hb >> 16;
                __syncwarp();                                             // every lane has read the chain state
                if (ok) {
                    if (st.free_hint >= 0) st.free_hint = free_next; else st.hw += 1;
                    s.slot[sl] = make_int4(end, (int)RLGS_NONE16, pr.node >= 0 ? (pr.node | (T << 16)) : (int)(0xffffu | ((uint32_t)pr.nnodes << 16)),
                                           pr.node >= 0 ? (int)pr.mask : st.log_len);
                    s.sjob[sl] = job;
                    if (tail == RLGS_NONE16) s.bkt[cal] = (uint32_t)sl | ((uint32_t)sl << 16);   // append: chains stay in start order
                    else { s.slot[tail].y = sl; s.bkt[cal] = (hb & 0xffffu) | ((uint32_t)sl << 16); }
                    if (writer) { D.start_tick[job] = d; D.place_off[job] = st.log_len; }
                    st.log_len += pr.nnodes;
                    st.start_seq += 1;
                    started = job + 1;
                    if (ROWS == 1) {
                        const int ndev = T * gpc;
                        st.busy_gpus += ndev;
                        st.mem_sum += hx.mem_term();
                        const int64_t mu = hx.util() & 0xffff, sd = hx.util() >> 16;
                        st.util_mu_sum += mu * ndev;
                        st.util_var_sum += sd * sd * ndev;
                        st.sum_arr -= hx.arrival();
                    }
                    st.sum_jct += (int64_t)(end - hx.arrival());   // the end is fixed at start (no preemption under fifo)
                    if (ENV && pick > 0) {
                        // queue.pop(pick): entries in front of it move one place towards the back of the stack (group-local)
                        for (int b0 = 0; b0 < pick; b0 += LPR) {   // window_k <= 32 may exceed the group width
                            const int i = pick - 1 - b0 - G.gl;    // from the back so that a chunk never overwrites an unread entry
                            JobRec mv;
                            if (i >= 0) mv = load_rec(D.stack + st.head + i);
                            G.sync();
                            if (i >= 0) store_rec(D.stack + st.head + i + 1, mv);
                            G.sync();
                        }
                    }
                    st.R += 1; st.Q -= 1; st.head += 1;
                    if (st.R > st.max_r) st.max_r = st.R;
                    if (st.Q > 0) {
                        h0n = load_rec(D.stack + st.head); h0_pending = true;   // consumed by the next attempt
                        if (ENV) st.bottom_arr = D.stack[st.head + st.Q - 1].arrival_tick;
                    }
                }
            }
            if (!ENV && attempt && !pr.ok && hx.fits()) st.head_blocked = 1;   // a failed attempt has no side effect: skip retries until a release
        }

        // median loads are issued early; they are consumed when the row is written
        int med_lo_arr = 0, med_hi_arr = 0;
        if (ROWS && ROWS != 5 && act && st.Q > 0) {   // rlgs_row4e carries no pending times: the host replays the queue
            med_lo_arr = D.stack[st.head + (st.Q - 1) / 2].arrival_tick;
            med_hi_arr = D.stack[st.head + st.Q / 2].arrival_tick;
        }

        // ---------------- delta_time += 1; step; release the jobs whose end tick is now, in start order (schedule.py:141-162)
        if (act) st.d = d + 1;
        {
            const int bk = st.d & (RLGS_CAL_W - 1);
            const uint32_t hb = s.bkt[bk];
            uint32_t head = hb & 0xffffu, tail = hb >> 16;
            int sl = act ? (int)head : (int)RLGS_NONE16, prev = (int)RLGS_NONE16;
            bool changed = false;
            while (__any_sync(RLGS_FULLMASK, sl != (int)RLGS_NONE16)) {
                const bool v = sl != (int)RLGS_NONE16;
                const int4 e = s.slot[v ? sl : 0];                  // one 16-byte load per hop: end, next, placement
                const int nx = e.y;
                const bool match = v && e.x == st.d;
                if (__any_sync(RLGS_FULLMASK, match)) {
                    const uint32_t place = (uint32_t)e.z, mask = (uint32_t)e.w;
                    const bool m1 = match && (place & 0xffffu) != 0xffffu;   // single-node job (the hot case)
                    const int ni = m1 ? (int)(place & 0xffffu) : 0;
                    const uint32_t kn = s.key[ni];
                    const uint32_t busy = NC::busy(s, ni, kn) & ~mask;
                    const int job = s.sjob[v ? sl : 0];
                    __syncwarp();                                   // every lane has read the slot, the job and the old node state
                    int ndev = 0;
                    if (m1) {                                       // Node.release_allocated_resources (node.py:71-91)
                        NC::store(s, ni, NC::free_units(kn) + (int)(place >> 16), busy, c);
                        ndev = __popc(mask);
                    } else if (match) {
                        ndev = grp_release_multi<LPR, PK>(G, s, c, D.place_log + (int)mask, (int)(place >> 16));   // group-local, rare
                    }
                    if (match) {
                        // unlink from the calendar chain, push on the free-slot chain (all lanes of the group store the same words)
                        if (prev == (int)RLGS_NONE16) head = (uint32_t)nx; else s.slot[prev].y = nx;
                        if (nx == (int)RLGS_NONE16) tail = (uint32_t)prev;
                        changed = true;
                        s.slot[sl] = make_int4(RLGS_NEVER, st.free_hint < 0 ? (int)RLGS_NONE16 : st.free_hint, 0, 0);
                        st.free_hint = sl;
                        if (ROWS == 1) {
                            const JobRec jr = load_rec(D.trace + job);   // the row sums need the job's constants again
                            st.busy_gpus -= ndev;
                            st.mem_sum -= jr.mem_term();
                            const int64_t mu = jr.util() & 0xffff, sd = jr.util() >> 16;
                            st.util_mu_sum -= mu * ndev;
                            st.util_var_sum -= sd * sd * ndev;
                        }
                        if (writer) { D.end_tick[job] = st.d; D.finish_order[st.F] = job; }
                        st.F += 1; st.R -= 1;
                        st.head_blocked = 0;   // resources were freed: the queue head may fit now
                    }
                }
                if (v) { if (!match) prev = sl; sl = nx; }
            }
            if (changed) s.bkt[bk] = head | (tail << 16);
        }

        // ---------------- stats row (schedule.py:204-205)
        if (act) {
            st.sumQ += st.Q; st.sumR += st.R;
            if (ENV) reward_acc -= (float)(st.Q + st.R);
            if (ROWS) {
                if (--rows_left < 0) {   // first row of the launch or of a chunk
                    row_cur = row_address<ROW_BYTES>(rs, rep, st.d - 1);
                    rows_left = RLGS_ROW_CHUNK - 1 - ((st.d - 1) & (RLGS_ROW_CHUNK - 1));
                }
                const int maxp = st.Q > 0 ? st.d - st.bottom_arr : 0, mlo = st.Q > 0 ? st.d - med_lo_arr : 0, mhi = st.Q > 0 ? st.d - med_hi_arr : 0;
                if (ROWS == 5) {   // rlgs_row4e: idle_nodes:12 | started:1 | queued[18:0]:19
                    if (writer) *reinterpret_cast<uint32_t *>(row_cur) = (uint32_t)st.idle_nodes | (started ? 0x1000u : 0u) | ((uint32_t)st.Q << 13);
                } else if (ROWS == 2) {
                    if (writer) *reinterpret_cast<int4 *>(row_cur) = pack_row16(st.idle_nodes, st.F, st.Q, maxp, mlo, mhi);
                } else if (ROWS == 3 || ROWS == 4) {   // rlgs_row12 / rlgs_row16e: lanes 0..2 (0..3) of the group store one word each
                    const uint32_t idle = (uint32_t)st.idle_nodes;
                    const uint32_t wv = G.gl == 0 ? ((uint32_t)maxp | (idle << 24)) : (G.gl == 1 ? ((uint32_t)mlo | ((idle >> 8) << 24)) : (G.gl == 2 ? (uint32_t)mhi : (uint32_t)started));
                    if (G.gl < (ROWS == 4 ? 4 : 3)) reinterpret_cast<uint32_t *>(row_cur)[G.gl] = wv;
                } else if (writer) {
                    const int64_t sp = (int64_t)st.Q * st.d - st.sum_arr;
                    int4 *o = reinterpret_cast<int4 *>(row_cur);
                    o[0] = make_int4(st.idle_nodes, st.busy_gpus, st.R, st.Q);
                    o[1] = make_int4(st.F, mlo, mhi, maxp);
                    o[2] = make_int4((int)(uint32_t)sp, (int)(sp >> 32), (int)(uint32_t)st.mem_sum, (int)(st.mem_sum >> 32));
                    o[3] = make_int4((int)(uint32_t)st.util_mu_sum, (int)(st.util_mu_sum >> 32), (int)(uint32_t)st.util_var_sum,
                                     (int)(st.util_var_sum >> 32));
                }
                row_cur += ROW_BYTES;
            }
        }
    }
#undef GBALLOT
#undef GSHFL
    if (!resident) return;                        // the remaining code is per group (group-local synchronisation only)
    if (ROWS >= 2 && ROWS <= 4 && st.d >= (1 << 24)) { st.status = RLGS_ERR_WIRE; st.done = 1; }   // pending times no longer fit the 24-bit wire fields

    st.events = (int64_t)st.cursor + st.start_seq + st.F;   // arrivals + starts + finishes
    if (!st.done && st.status == RLGS_OK && (J - st.cursor) + st.R == 0) st.done = 1;   // the while condition of schedule.py:185
    if (ENV) {
        // observation: per node free GPUs / cpu / mem, the look-ahead window, queue statistics
        float *o = env.obs + (size_t)rep * env.obs_dim;
        for (int i = G.gl; i < c.N; i += LPR) {
            const uint32_t key = s.key[i];
            const int units = c.base_units - NC::free_units(key);   // tasks charged to the node: cpu_used = 12 u, mem_used = 60 u
            o[i] = (float)NC::idle(key);
            o[c.N + i] = (float)(c.cpu_cap - RLGS_CPUS_PER_TASK * units);
            o[2 * c.N + i] = (float)(c.mem_cap - RLGS_MEM_PER_TASK * units);
        }
        for (int i = G.gl; i < env.window_k; i += LPR) {
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            float idx = -1.f;
            if (i < st.Q) { JobRec r = load_rec(D.stack + st.head + i); w = make_float4((float)r.gpus(), (float)r.tasks(), (float)r.dur(), (float)(st.d - r.arrival())); idx = (float)r.index(); }
            float *ow = o + 3 * c.N + 5 * i;
            ow[0] = w.x; ow[1] = w.y; ow[2] = w.z; ow[3] = w.w; ow[4] = idx;
        }
        if (writer) {
            float *t = o + 3 * c.N + 5 * env.window_k;
            t[0] = (float)st.Q; t[1] = (float)st.R; t[2] = (float)st.F; t[3] = (float)st.d;
            env.reward[rep] = reward_acc;
            env.done[rep] = (uint8_t)(st.done != 0);
        }
    }
    grp_state_io<LPR, PK>(G, D, s, c, slot_cap, st, true);
    if (writer) {
        states[rep] = st;
        if (st.done) returns[rep] = -st.sum_jct;   // episode return, read by the all-gather
    }
}
