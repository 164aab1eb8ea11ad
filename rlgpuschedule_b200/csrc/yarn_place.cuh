// yarn_place.cuh — the yarn placement fit as a warp-ballot scan over the per-node occupancy
// counters and busy-device bitmasks staged in shared memory.
//
// Restates (reference paths):
//   ms_yarn_placement            core/scheduling/algorithm.py:28-32
//   try_single_node_alloc_ms     core/scheduling/algorithm.py:396-417  (+ Node.try_alloc_job infra/node.py:245-275)
//   try_cross_node_alloc_ms      core/scheduling/algorithm.py:301-393  (q11: tries fit+1 tasks, no side effect)
//   Node.try_reserve_and_placed_task  infra/node.py:200-221  (q8: cpu/mem charged before the device loop,
//                                     never undone when no device accepts the task)
// Node state = charged task units + busy-device mask + the derived fit key (see NodeView in rlgs_device.cuh).
// Lane l owns nodes l, l+32, ...; node ids ascend with the index, so "first node in id order" is
// the lowest set bit of the first non-empty ballot.
#pragma once
#include "rlgs_device.cuh"

struct PlaceResult {
    int ok;          // 1 placed, 0 not placed
    int node;        // single-node: node index; multi-node: -1
    uint32_t mask;   // single-node: device mask
    int nnodes;      // entries appended to the placement log
};

// Charges `k` tasks worth of cpu/mem to the nodes selected by `pred` (one node per lane per call),
// keeps n_free_nodes consistent.  Used by the q8 leak paths.
__device__ __forceinline__ void charge_nodes(NodeView nv, const ClusterConst &c, int i, bool pred, int k, int &n_free_nodes) {
    bool was = false, now = false;
    if (pred) {
        int u = nv.units[i];
        was = node_is_free(u, c);
        u += k;
        now = node_is_free(u, c);
        nv.units[i] = u; nv.key[i] = node_key(u, nv.busy[i], c);
    }
    n_free_nodes += __popc(__ballot_sync(RLGS_FULL, now)) - __popc(__ballot_sync(RLGS_FULL, was));
}

__device__ __forceinline__ int node_task_capacity(NodeView nv, const ClusterConst &c, int i, int gpc) {
    // min(free_devices // gpc, cpu_free // 12, mem_free // 60), python floor division; anything <= 0
    // places nothing (Node.can_fit_num_task, infra/node.py:109-127)
    (void)c;
    uint32_t k = nv.key[i];
    return min((int)(k >> 16) / gpc, (int)(k & 0xffff));
}

// Tries to place job `j` under yarn.  On success the node counters / masks are updated and the
// placement is appended to place_log[log_pos ...].  All lanes must call; result is warp-uniform.
__device__ __forceinline__ PlaceResult yarn_place(NodeView nv, const ClusterConst &c, const JobRec &j, int lane,
                                                  int2 *place_log, int log_pos, int &n_free_nodes, int &idle_nodes) {
    PlaceResult r; r.ok = 0; r.node = -1; r.mask = 0; r.nnodes = 0;
    const int T = j.tasks(), gpc = j.gpc(), need_g = j.gpus();
    const bool fits = j.fits();
    if (need_g <= c.G) {
        // ---- single node: first node (id order) with enough idle devices, cpu and mem ----
        int node = -1;
        for (int base = 0; base < c.N; base += 32) {
            int i = base + lane;
            bool ok = false;
            if (i < c.N) {
                // idle devices >= gpus and cpu_free >= 12T and mem_free >= 60T (algorithm.py:407-409), from the node key
                uint32_t k = nv.key[i];
                ok = (int)(k >> 16) >= need_g && (int)(k & 0xffff) >= T;
            }
            if (!fits) {  // no device accepts the task: every candidate node leaks T tasks of cpu/mem (q8)
                charge_nodes(nv, c, i, ok, T, n_free_nodes);
                continue;
            }
            unsigned b = __ballot_sync(RLGS_FULL, ok);
            if (b) { node = base + __ffs(b) - 1; break; }
        }
        __syncwarp();
        if (node < 0) return r;
        int u = nv.units[node];
        uint32_t busy = nv.busy[node];
        uint32_t taken = lowest_bits_warp(~busy & c.gmask, T * gpc, lane);   // the warp is converged here, node / busy are uniform
        bool was = node_is_free(u, c);
        u += T;
        n_free_nodes += (int)node_is_free(u, c) - (int)was;
        uint32_t ew = nv.ever[node >> 5], bit = 1u << (node & 31);
        if (!(ew & bit)) idle_nodes--;
        __syncwarp();
        if (lane == 0) {
            nv.units[node] = u; nv.busy[node] = busy | taken; nv.ever[node >> 5] = ew | bit;
            nv.key[node] = node_key(u, busy | taken, c);
            place_log[log_pos] = make_int2(node | (T << 16), (int)taken);
        }
        __syncwarp();
        r.ok = 1; r.node = node; r.mask = taken; r.nnodes = 1;
        return r;
    }
    // ---- cross node: walk nodes in id order, each takes min(capacity, remaining) tasks ----
    if (!fits) {  // the first task attempt on every node with capacity >= 1 leaks one task of cpu/mem
        for (int base = 0; base < c.N; base += 32) {
            int i = base + lane;
            bool ok = (i < c.N) && node_task_capacity(nv, c, i, gpc) >= 1;
            charge_nodes(nv, c, i, ok, 1, n_free_nodes);
        }
        __syncwarp();
        return r;
    }
    int remaining = T, nodes_assigned = 0;
    for (int base = 0; base < c.N && remaining > 0; base += 32) {
        int i = base + lane;
        int cap = (i < c.N) ? max(node_task_capacity(nv, c, i, gpc), 0) : 0;
        int incl = warp_incl_scan(cap, lane);
        int take = max(0, min(cap, remaining - (incl - cap)));
        int tot = __shfl_sync(RLGS_FULL, incl, 31);
        nodes_assigned += __popc(__ballot_sync(RLGS_FULL, take > 0));
        remaining -= min(remaining, tot);
    }
    if (remaining > 0 || nodes_assigned < j.least()) return r;  // rollback: no net effect (algorithm.py:376-386)
    // commit pass: recompute the same takes and apply them
    remaining = T;
    int written = 0;
    for (int base = 0; base < c.N && remaining > 0; base += 32) {
        int i = base + lane;
        int cap = (i < c.N) ? max(node_task_capacity(nv, c, i, gpc), 0) : 0;
        int incl = warp_incl_scan(cap, lane);
        int take = max(0, min(cap, remaining - (incl - cap)));
        int tot = __shfl_sync(RLGS_FULL, incl, 31);
        unsigned tb = __ballot_sync(RLGS_FULL, take > 0);
        bool was = false, now = false;
        if (take > 0) {
            int u = nv.units[i];
            uint32_t busy = nv.busy[i];
            uint32_t taken = lowest_bits(~busy & c.gmask, take * gpc);
            was = node_is_free(u, c);
            u += take;
            now = node_is_free(u, c);
            nv.units[i] = u; nv.busy[i] = busy | taken; nv.key[i] = node_key(u, busy | taken, c);
            place_log[log_pos + written + __popc(tb & ((1u << lane) - 1))] = make_int2(i | (take << 16), (int)taken);
        }
        n_free_nodes += __popc(__ballot_sync(RLGS_FULL, now)) - __popc(__ballot_sync(RLGS_FULL, was));
        uint32_t ew = nv.ever[base >> 5];
        idle_nodes -= __popc(tb & ~ew);
        __syncwarp();
        if (lane == 0) nv.ever[base >> 5] = ew | tb;
        written += __popc(tb);
        remaining -= min(remaining, tot);
    }
    __syncwarp();
    r.ok = 1; r.nnodes = written;
    return r;
}

// Releases one placement entry (Node.release_allocated_resources, infra/node.py:71-91).
// One entry per lane (`active` lanes), distinct nodes.
__device__ __forceinline__ void release_entry(NodeView nv, const ClusterConst &c, bool active, int2 e, int &n_free_nodes) {
    bool was = false, now = false;
    if (active) {
        int node = e.x & 0xffff, tasks = (e.x >> 16) & 0xffff;
        int u = nv.units[node];
        was = node_is_free(u, c);
        u -= tasks;
        now = node_is_free(u, c);
        uint32_t busy = nv.busy[node] & ~(uint32_t)e.y;
        nv.units[node] = u; nv.busy[node] = busy; nv.key[node] = node_key(u, busy, c);
    }
    n_free_nodes += __popc(__ballot_sync(RLGS_FULL, now)) - __popc(__ballot_sync(RLGS_FULL, was));
}

// Same release for a single-node job: every lane computes the same values, lane 0 stores them.
__device__ __forceinline__ void release_single(NodeView nv, const ClusterConst &c, int node, int tasks, uint32_t mask, int lane, int &n_free_nodes) {
    int u = nv.units[node];
    const bool was = node_is_free(u, c);
    u -= tasks;
    n_free_nodes += (int)node_is_free(u, c) - (int)was;
    const uint32_t busy = nv.busy[node] & ~mask;
    __syncwarp();
    if (lane == 0) { nv.units[node] = u; nv.busy[node] = busy; nv.key[node] = node_key(u, busy, c); }
}
