// rlgs_device.cuh — device-side data layout and warp-level building blocks shared by the simulation
// kernels (sm_100a).  One warp advances one replica; cluster/node state is staged in shared memory,
// the per-job structure-of-records streams from / to HBM.
//
// Reference semantics restated here (paths relative to the reference repo):
//   node occupancy predicates   infra/node.py:51-60,99-127,146-171,200-221
//   device accept rule          infra/device.py:19-43,67-77   (yarn never shares a device)
//   yarn single / cross fit     core/scheduling/algorithm.py:28-32,301-417
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rlgs.h"

#define RLGS_FULL 0xffffffffu
#define RLGS_CPUS_PER_TASK 12  // core/jobs/job.py:105
#define RLGS_MEM_PER_TASK 60   // core/jobs/job.py:106
#define RLGS_NEVER 0x7fffffff

// Per-replica scalars that survive between kernel launches (chunked runs, env steps).
struct RepState {
    int32_t d;            // current tick (delta_time of schedule.py:180)
    int32_t cursor;       // jobs of the trace already arrived (JobTraceReader "generated" column)
    int32_t head;         // index of the queue front inside the stack array
    int32_t Q, R, F;      // queued / running / finished job counts
    int32_t hw;           // slot high-water mark
    int32_t idle_nodes, busy_gpus, n_free_nodes;
    int32_t start_seq;    // starts so far (orders same-tick finishes like the running_jobs dict)
    int32_t bottom_arr;   // arrival tick of the oldest queued job (max_pending_time)
    int32_t free_hint;    // a free slot index or -1
    int32_t head_blocked; // last fit attempt of the current head failed and nothing was released since
    int32_t done, status;
    int32_t max_q, max_r, log_len, pad0;
    int64_t mem_sum, util_mu_sum, util_var_sum;
    int64_t sum_arr;      // sum of arrival ticks of queued jobs: sum_pending = Q*d - sum_arr
    int64_t sumQ, sumR, sum_jct, events;
};

// Per-replica pointers (all device memory; computed on the host once).
struct RepDesc {
    const rlgs_job *trace;  // [J] 32-byte records in arrival order (may be shared between replicas)
    rlgs_job *stack;        // [J] fifo queue as a stack growing towards index 0 (front insertion, q1)
    int32_t *start_tick;    // [J]
    int32_t *end_tick;      // [J]
    int32_t *finish_order;  // [J]
    int32_t *place_off;     // [J] first entry of the job's placement in place_log (-1 = never placed)
    int2 *place_log;        // [log_cap] (node | ntasks<<16, device mask) per (job, node), in node order
    int32_t *node_save;     // [3N + ceil(N/32)] cpu_used, mem_used, busy mask, ever-used bitmap
    int4 *slot_save;        // [2*slot_cap]
    const double *net_in;   // [3][J]: duration, model MB, iterations (network-cost inputs; may be null)
    double *dur_out;        // [J] duration after network costs (null unless enabled)
    int32_t J;
    int32_t log_cap;
};

// PS/worker transfer cost added to a job's duration when it is placed (core/network/network_service.py:3-39).
struct NetCost {
    int32_t enabled, pad;
    double bandwidth, latency;   // --bandwidth (MB/s), --internode_latency (s)
};

struct ClusterConst {
    int32_t N, G;          // nodes, GPUs per node
    int32_t cpu_cap, mem_cap;
    int32_t base_units;    // min(cpu_cap // 12, mem_cap // 60): tasks an empty node's cpu and mem can take
    int32_t free_limit;    // a node is "free" (cpu_free > 0 or mem_free > 0, node.py:59) while its charged units < this
    uint32_t gmask;        // G low bits set
    int32_t D;             // N*G
    int32_t free_floor;    // base_units - free_limit: a node is "free" while its free task units exceed this (fifo_grp.cuh)
};

// chunk-major row store shared by all kernels: row i of replica r lives in chunk i / RLGS_ROW_CHUNK
#define RLGS_ROW_CHUNK_LOG 12
#define RLGS_ROW_CHUNK (1 << RLGS_ROW_CHUNK_LOG)
struct RowStore {
    rlgs_row *const *chunks;  // device array of chunk base pointers, each [n_replicas][RLGS_ROW_CHUNK]
    int32_t n_chunks;
    int32_t replica;          // global replica index of blockIdx.x == 0
};
__device__ __forceinline__ rlgs_row *row_ptr(const RowStore &rs, int replica_local, int64_t i) {
    return rs.chunks[i >> RLGS_ROW_CHUNK_LOG] + ((size_t)(rs.replica + replica_local) << RLGS_ROW_CHUNK_LOG) + (i & (RLGS_ROW_CHUNK - 1));
}

// volatile: read the special register once and keep the value — under register pressure the compiler otherwise re-reads
// SR_TID.X (S2R, tens of cycles) wherever `lane` is used (4.7 % of the instructions of dlas_gpu_kernel in the r02 capture)
__device__ __forceinline__ int lane_id() { int l; asm volatile("mov.u32 %0, %%laneid;" : "=r"(l)); return l; }

// mask of the k lowest set bits of `freemask` (devices are taken in device-id order, node.py:209-219)
__device__ __forceinline__ uint32_t lowest_bits(uint32_t freemask, int k) {
    if (k == 1) return freemask & (0u - freemask);   // most jobs take one device
    uint32_t rem = freemask;
    for (int i = 0; i < k; ++i) rem &= rem - 1;
    return freemask ^ rem;
}
// The same for a full, converged warp whose lanes all hold the same (freemask, k): lane i decides bit i, no loop.
__device__ __forceinline__ uint32_t lowest_bits_warp(uint32_t freemask, int k, int lane) {
    const bool keep = ((freemask >> lane) & 1u) && __popc(freemask & ((1u << lane) - 1u)) < k;
    return __ballot_sync(RLGS_FULL, keep);
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(RLGS_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__device__ __forceinline__ int4 shfl_int4(int4 v, int src) {
    int4 r;
    r.x = __shfl_sync(RLGS_FULL, v.x, src);
    r.y = __shfl_sync(RLGS_FULL, v.y, src);
    r.z = __shfl_sync(RLGS_FULL, v.z, src);
    r.w = __shfl_sync(RLGS_FULL, v.w, src);
    return r;
}

// Decoded view of a 32-byte job record held as two int4 (see rlgs_job in include/rlgs.h).
struct JobRec {
    int4 a, b;
    __device__ __forceinline__ int arrival() const { return a.x; }
    __device__ __forceinline__ int dur() const { return a.y; }
    __device__ __forceinline__ int gpus() const { return a.z & 0xffff; }
    __device__ __forceinline__ int tasks() const { return (a.z >> 16) & 0xffff; }
    __device__ __forceinline__ int gpc() const { return a.w & 0xffff; }
    __device__ __forceinline__ int least() const { return (a.w >> 16) & 0x7fff; }
    __device__ __forceinline__ bool fits() const { return (a.w >> 31) & 1; }
    __device__ __forceinline__ int64_t mem_term() const { return (int64_t)(((uint64_t)(uint32_t)b.y << 32) | (uint32_t)b.x); }
    __device__ __forceinline__ uint32_t util() const { return (uint32_t)b.z; }
    __device__ __forceinline__ int index() const { return b.w; }
};

// Shared-memory view of the simulated cluster of one replica.
// Every task charges exactly 12 cpus and 60 memory units to its node (core/jobs/job.py:105-106) — also on the
// q8 leak path — so cpu_used = 12 u and mem_used = 60 u for one per-node counter u ("units"):
//   cpu_free // 12 = cpu_cap // 12 - u,  mem_free // 60 = mem_cap // 60 - u   (both clamp at <= 0 the same way)
//   is_free  <=>  u < max(ceil(cpu_cap / 12), ceil(mem_cap / 60))
// which removes every division from the per-tick path.
struct NodeView {
    int32_t *units;   // tasks charged to the node (placed + leaked)
    uint32_t *busy;   // busy-device bitmask per node
    uint32_t *ever;   // bitmap: node ever held a placed job (Node.placed_jobs is never cleared, q3)
    uint32_t *key;    // idle devices << 16 | tasks the free cpu/mem can take: the two numbers every fit test needs
};

__host__ __device__ inline int rlgs_ceil_div_pos(int x, int d) { return x > 0 ? (x + d - 1) / d : 0; }

// key = popc(idle devices) << 16 | min(cpu_free // 12, mem_free // 60) clamped to [0, 65535]
__device__ __forceinline__ uint32_t node_key(int units, uint32_t busy, const ClusterConst &c) {
    int t = min(max(c.base_units - units, 0), 0xffff);
    return ((uint32_t)__popc(~busy & c.gmask) << 16) | (uint32_t)t;
}

__device__ __forceinline__ bool node_is_free(int units, const ClusterConst &c) {
    return units < c.free_limit;  // cpu_free > 0 or mem_free > 0 (infra/node.py:59-60)
}
