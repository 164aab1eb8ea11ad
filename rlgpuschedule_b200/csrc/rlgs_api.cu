// rlgs_api.cu — host side of the C ABI declared in include/rlgs.h (librlgs.so).
//
// Owns device memory, launches the simulation kernels, keeps the per-tick / per-event rows in a
// chunk-major device store (chunk k = rows [k*4096, (k+1)*4096) of every replica, contiguous) and,
// in rows_mode FULL, streams chunk k to a pinned host mirror on a copy stream while chunk k+1 is
// being simulated.  No torch types, no CPU fallback: every entry point fails with RLGS_ERR_CUDA
// when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/rlgs.h"
#include "fifo_grp.cuh"
#include "legacy_sched.cuh"
#include "pack_horus.cuh"

static thread_local char g_err[512] = "";

static int32_t fail(int32_t code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(e_ == cudaErrorMemoryAllocation ? RLGS_ERR_OOM : RLGS_ERR_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e_));                                                  \
    } while (0)

static const int N_PLANES = 6;   // start, end, finish_order, {place_off | pending_time}, preempt, resume
static const int MAX_CHUNKS = 1 << 16;

struct TraceBuf {
    rlgs_job *dev = nullptr;
    double *net = nullptr;      // [3][cap_n] network-cost inputs
    double *dur_out = nullptr;  // [count][cap_n]
    PackJob *pack = nullptr;    // [n] horus placement inputs
    PlusFeat *feat = nullptr;   // [n] horus+ k-means features
    void *pack_slab = nullptr;  // per-replica working set of the pack kernels for this trace
    void *slab = nullptr;       // per-replica working set of the fifo / legacy kernels for this trace
    int32_t n = 0, cap_n = 0;
    int64_t log_cap = 0, cap_log = 0;
    int32_t max_arrival = 0;
    int first = 0, count = 0;   // replica range this trace is attached to
    std::vector<rlgs_job> host; // host copy of the records: the expansion of the 16-byte wire rows needs the per-job constants
};

struct Group {            // a contiguous range of replicas driven through one CUDA stream
    int first = 0, count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t k_end = nullptr;
};

// Host copy of the per-replica progress, schedule independent
struct Progress {
    int64_t rows = 0;
    int done = 0, status = 0;
};

struct rlgs_sim {
    rlgs_cluster_spec spec;
    rlgs_opts opts;
    ClusterConst cc;
    LegParams lp;
    int R = 0, device = 0;
    int lpr = 32;           // lanes of a warp per replica (fifo tick loop)
    int wire = 0;           // 0 = rows are kept as rlgs_row, 1 = rlgs_row16, 2 = rlgs_row12, 3 = rlgs_row16e, 4 = rlgs_row4e (RLGS_ROWFMT_*)
    std::vector<char> derived;   // per replica: job planes 0..2 in h_jobs were rebuilt from the event rows
    int planes_mask = 0;    // bit k: job plane k is on the host (h_jobs)
    size_t row_bytes = sizeof(rlgs_row);
    bool env_ready = false; // rlgs_env_reset ran since the last rlgs_load_trace
    int64_t env_ticks = 0;  // upper bound of the ticks simulated since rlgs_env_reset (sizes the row store while stepping)
    int xp_replica = -1;    // replica whose prefix sums are cached below (expansion of wire rows)
    std::vector<int64_t> xp[9];
    std::vector<int32_t> xp_start;   // start ticks of the started jobs, ascending
    std::vector<int32_t> xp_pend;    // rlgs_row4e: max_pending, median_lo, median_hi of every row of xp_replica (queue replay)
    bool legacy = false;
    bool pack = false;      // horus schedule + horus placement (pack_horus.cuh)
    PackParams pp;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    void *user_stream = nullptr;
    bool use_user_stream = false;   // NULL is a valid handle: the legacy default stream
    cudaEvent_t ev_fork = nullptr;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<Group> groups;
    std::vector<TraceBuf> traces;
    std::vector<int> rep_trace;
    // fifo
    std::vector<RepDesc> h_desc;
    RepDesc *d_desc = nullptr;
    RepState *d_state = nullptr, *h_state = nullptr, *h_init = nullptr;
    // legacy
    std::vector<LegDesc> h_ldesc;
    LegDesc *d_ldesc = nullptr;
    LegState *d_lstate = nullptr, *h_lstate = nullptr, *h_linit = nullptr;
    // pack
    std::vector<PackDesc> h_pdesc;
    PackDesc *d_pdesc = nullptr;
    PackState *d_pstate = nullptr, *h_pstate = nullptr, *h_pinit = nullptr;
    int slot_cap = 0;
    // chunk-major row store
    std::vector<rlgs_row *> d_chunks, h_chunks;
    std::vector<char> h_chunk_valid;
    rlgs_row **d_chunk_ptrs = nullptr;
    // job planes [N_PLANES][R][Jmax]
    int32_t *d_jobs = nullptr, *h_jobs = nullptr;
    bool h_jobs_pinned = false;   // pinned when rlgs_run copies tables itself (opts.fetch_jobs); pageable (lazily committed) when only read on demand
    size_t jobs_bytes = 0;
    bool env_used = false;  // job tables were produced by environment steps (picks inside the window: start = end - dur still holds)
    int32_t Jmax = 0;
    int64_t *d_returns = nullptr, *h_returns = nullptr;
    bool ran = false;
    float last_ms = 0.f;
    int last_launches = 0;
    int64_t rows_hint = 0;   // rows of the longest replica in the previous run: sizes the speculative pipeline
};

static void free_host_jobs(rlgs_sim *s) {
    if (!s->h_jobs) return;
    if (s->h_jobs_pinned) cudaFreeHost(s->h_jobs); else free(s->h_jobs);
    s->h_jobs = nullptr;
}
static int32_t alloc_host_jobs(rlgs_sim *s, bool pinned) {
    if (s->h_jobs && (s->h_jobs_pinned || !pinned)) return RLGS_OK;
    free_host_jobs(s);
    if (pinned) { CU(cudaMallocHost(&s->h_jobs, s->jobs_bytes)); }
    else { s->h_jobs = static_cast<int32_t *>(calloc(s->jobs_bytes, 1)); if (!s->h_jobs) return fail(RLGS_ERR_OOM, "host allocation of the job tables failed"); }
    s->h_jobs_pinned = pinned;
    return RLGS_OK;
}

extern "C" int32_t rlgs_version(void) { return RLGS_VERSION; }
extern "C" const char *rlgs_last_error(void) { return g_err; }

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static void free_host_jobs(rlgs_sim *s);
static int32_t alloc_host_jobs(rlgs_sim *s, bool pinned);
static size_t chunk_bytes(const rlgs_sim *s) { return s->row_bytes * (size_t)RLGS_ROW_CHUNK * (size_t)s->R; }

extern "C" int32_t rlgs_create(const rlgs_cluster_spec *spec, const rlgs_opts *opts, rlgs_sim **out) {
    if (!spec || !opts || !out) return fail(RLGS_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    int64_t N = (int64_t)spec->num_switch * spec->num_node_p_switch;
    if (spec->num_switch < 1 || spec->num_node_p_switch < 1 || N > 4096)
        return fail(RLGS_ERR_BAD_ARG, "cluster must have 1..4096 nodes (got %lld)", (long long)N);
    if (spec->num_gpu_p_node < 1 || spec->num_gpu_p_node > 32)
        return fail(RLGS_ERR_BAD_ARG, "num_gpu_p_node must be 1..32 (got %d)", spec->num_gpu_p_node);
    if (opts->n_replicas < 1) return fail(RLGS_ERR_BAD_ARG, "n_replicas must be >= 1");
    const int sched = opts->schedule;
    if (sched != RLGS_SCHED_FIFO && sched != RLGS_SCHED_SJF && sched != RLGS_SCHED_DLAS_GPU && sched != RLGS_SCHED_DLAS &&
        sched != RLGS_SCHED_SHORTEST && sched != RLGS_SCHED_SHORTEST_GPU && sched != RLGS_SCHED_HORUS && sched != RLGS_SCHED_GANDIVA && sched != RLGS_SCHED_HORUS_PLUS)
        return fail(RLGS_ERR_UNSUPPORTED, "schedule id %d is not implemented on the device path", sched);
    const bool is_sjf_family = sched == RLGS_SCHED_SJF || sched == RLGS_SCHED_SHORTEST || sched == RLGS_SCHED_SHORTEST_GPU;
    if ((sched == RLGS_SCHED_FIFO || is_sjf_family) && opts->placement != RLGS_PLACE_YARN)
        return fail(RLGS_ERR_UNSUPPORTED, "placement id %d is not implemented for schedule id %d", opts->placement, sched);
    const bool is_pack = sched == RLGS_SCHED_HORUS || sched == RLGS_SCHED_GANDIVA || sched == RLGS_SCHED_HORUS_PLUS;
    if (sched == RLGS_SCHED_HORUS_PLUS && (opts->num_queue < 1 || opts->num_queue > PACK_MAX_Q)) return fail(RLGS_ERR_BAD_ARG, "horus+: num_queue must be 1..%d", PACK_MAX_Q);
    if (!is_pack && opts->placement == RLGS_PLACE_HORUS)
        return fail(RLGS_ERR_UNSUPPORTED, "the pack placement needs the horus or gandiva schedule (schedule.py:47 passes the schedule name "
                    "to the placement's score table, so fifo + horus raises KeyError in the reference)");
    if (is_pack && opts->placement != RLGS_PLACE_HORUS && opts->placement != RLGS_PLACE_YARN)
        return fail(RLGS_ERR_UNSUPPORTED, "placement id %d is not implemented for schedule id %d", opts->placement, sched);
    if (is_pack && (opts->num_buffer < 0 || opts->num_buffer > 32)) return fail(RLGS_ERR_BAD_ARG, "num_buffer must be 0..32");
    const bool is_dlas = sched == RLGS_SCHED_DLAS_GPU || sched == RLGS_SCHED_DLAS;
    if (is_dlas) {
        if (opts->num_queue < 1 || opts->num_queue > RLGS_MAX_QUEUES) return fail(RLGS_ERR_BAD_ARG, "num_queue must be 1..%d", RLGS_MAX_QUEUES);
        for (int q = 0; q + 1 < opts->num_queue; ++q)
            if (opts->queue_limit[q] < 1) return fail(RLGS_ERR_BAD_ARG, "queue_limit[%d] must be >= 1", q);
    }
    if (opts->enable_network_costs && sched != RLGS_SCHED_FIFO) return fail(RLGS_ERR_UNSUPPORTED, "network costs are implemented for the fifo tick loop only");
    if (opts->enable_network_costs && !(opts->bandwidth > 0)) return fail(RLGS_ERR_BAD_ARG, "bandwidth must be > 0");
    if (opts->rows_format < RLGS_ROWFMT_WIDE || opts->rows_format > RLGS_ROWFMT_EVENT4) return fail(RLGS_ERR_BAD_ARG, "rows_format must be one of RLGS_ROWFMT_*");
    if (opts->rows_format != RLGS_ROWFMT_WIDE && sched != RLGS_SCHED_FIFO) return fail(RLGS_ERR_UNSUPPORTED, "the wire rows belong to the fifo tick loop");
    if (opts->rows_format != RLGS_ROWFMT_WIDE && N > 4095) return fail(RLGS_ERR_UNSUPPORTED, "a wire row holds at most 4095 nodes");
    if (opts->rows_format == RLGS_ROWFMT_EVENT4 && opts->enable_network_costs) return fail(RLGS_ERR_UNSUPPORTED, "RLGS_ROWFMT_EVENT4 needs end = start + dur_ticks: no network costs");
    if (opts->fetch_jobs < 0 || opts->fetch_jobs > 2) return fail(RLGS_ERR_BAD_ARG, "fetch_jobs must be 0, 1 or 2");
    if (opts->fetch_jobs == 2 && (sched != RLGS_SCHED_FIFO || opts->enable_network_costs)) return fail(RLGS_ERR_UNSUPPORTED, "fetch_jobs = 2 needs the fifo tick loop without network costs (start = end - dur_ticks)");
    const int lpr_in = opts->lanes_per_replica;
    if (lpr_in != 0 && lpr_in != 8 && lpr_in != 16 && lpr_in != 32) return fail(RLGS_ERR_BAD_ARG, "lanes_per_replica must be 0, 8, 16 or 32");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(RLGS_ERR_CUDA, "no CUDA device: %s (librlgs has no CPU fallback)", cudaGetErrorString(e));
    if (opts->device < 0 || opts->device >= ndev) return fail(RLGS_ERR_BAD_ARG, "device %d out of range", opts->device);
    CU(cudaSetDevice(opts->device));
    rlgs_sim *s = new (std::nothrow) rlgs_sim();
    if (!s) return fail(RLGS_ERR_OOM, "host allocation failed");
    s->spec = *spec; s->opts = *opts; s->R = opts->n_replicas; s->device = opts->device;
    s->legacy = sched != RLGS_SCHED_FIFO;
    s->pack = is_pack;
    memset(&s->pp, 0, sizeof s->pp);
    s->pp.num_buffer = opts->num_buffer > 0 ? opts->num_buffer : 5;
    s->pp.rng_on = opts->pack_rng != 0; s->pp.seed = opts->pack_seed; s->pp.gandiva = sched == RLGS_SCHED_GANDIVA;
    s->pp.plus_k = sched == RLGS_SCHED_HORUS_PLUS ? opts->num_queue : 0; s->pp.plus_seed = opts->pack_seed;
    s->pp.nodes_per_rack = spec->num_node_p_switch; s->pp.racks = spec->num_switch; s->pp.max_ticks = opts->max_ticks;
    s->cc.N = (int)N; s->cc.G = spec->num_gpu_p_node; s->cc.cpu_cap = spec->num_cpu_p_node; s->cc.mem_cap = spec->mem_p_node;
    s->cc.gmask = spec->num_gpu_p_node == 32 ? 0xffffffffu : ((1u << spec->num_gpu_p_node) - 1u);
    s->cc.D = s->cc.N * s->cc.G;
    s->cc.base_units = std::max(0, std::min(s->cc.cpu_cap > 0 ? s->cc.cpu_cap / RLGS_CPUS_PER_TASK : 0, s->cc.mem_cap > 0 ? s->cc.mem_cap / RLGS_MEM_PER_TASK : 0));
    s->cc.free_limit = std::max(rlgs_ceil_div_pos(s->cc.cpu_cap, RLGS_CPUS_PER_TASK), rlgs_ceil_div_pos(s->cc.mem_cap, RLGS_MEM_PER_TASK));
    s->cc.free_floor = s->cc.base_units - s->cc.free_limit;
    if (sched == RLGS_SCHED_FIFO && s->cc.base_units > 0x7fff) { delete s; return fail(RLGS_ERR_UNSUPPORTED, "a node takes at most 32767 tasks (num_cpu_p_node / 12, mem_p_node / 60)"); }
    memset(&s->lp, 0, sizeof s->lp);
    s->lp.nq = is_dlas ? opts->num_queue : 1;
    s->lp.gputime = sched == RLGS_SCHED_DLAS_GPU;
    s->lp.sort_mode = sched == RLGS_SCHED_SHORTEST ? 1 : (sched == RLGS_SCHED_SHORTEST_GPU ? 2 : 0);
    for (int q = 0; q < RLGS_MAX_QUEUES; ++q) s->lp.limit[q] = opts->queue_limit[q];
    s->lp.total_gpu = s->cc.D; s->lp.num_node = s->cc.N; s->lp.gpus_per_node = s->cc.G; s->lp.max_time = opts->max_ticks;
    s->slot_cap = opts->slot_cap > 0 ? opts->slot_cap : std::min(128, std::max(32, s->cc.D));
    s->slot_cap = (s->slot_cap + 31) & ~31;
    if (s->slot_cap > 65504) { delete s; return fail(RLGS_ERR_BAD_ARG, "slot_cap must be <= 65504"); }
    s->wire = opts->rows_format;
    s->row_bytes = s->wire == RLGS_ROWFMT_WIRE16 ? sizeof(rlgs_row16) : (s->wire == RLGS_ROWFMT_WIRE12 ? sizeof(rlgs_row12) : (s->wire == RLGS_ROWFMT_EVENT16 ? sizeof(rlgs_row16e) : (s->wire == RLGS_ROWFMT_EVENT4 ? sizeof(rlgs_row4e) : sizeof(rlgs_row))));
    s->derived.assign(s->R, 0);
    // lanes per replica: a warp carries 32 / lpr replicas.  Few replicas -> wide groups (more SMs busy, shortest tick);
    // many replicas -> narrow groups (every warp instruction serves 4 replicas).  148 SMs x >= 8 warps each.
    s->lpr = lpr_in ? lpr_in : (s->R >= 148 * 8 * 4 ? 8 : (s->R >= 148 * 8 * 2 ? 16 : 32));
    if (opts->enable_network_costs) s->lpr = 32;   // the float64 network-cost variant is compiled for one replica per warp
    while (s->lpr < 32 && (32 / s->lpr) * grp_smem_bytes(s->cc.N, s->cc.G, s->slot_cap, s->lpr) > 227 * 1024) s->lpr *= 2;   // the warp's replicas must fit one SM
    s->rep_trace.assign(s->R, -1);
    s->h_desc.assign(s->R, RepDesc{});
    s->h_ldesc.assign(s->R, LegDesc{});
    if (s->pack) s->h_pdesc.assign(s->R, PackDesc{});
    int ng = opts->n_streams > 0 ? opts->n_streams : (s->R >= 8 * 148 ? 4 : (s->R >= 2 * 148 ? 2 : 1));
    ng = std::max(1, std::min(ng, s->R));
    cudaError_t ce = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (ce == cudaSuccess) ce = x; };
    ok(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    ok(cudaEventCreate(&s->ev_fork));
    for (int g = 0; g < ng; ++g) {
        Group G;
        G.first = (int)((int64_t)s->R * g / ng); G.count = (int)((int64_t)s->R * (g + 1) / ng) - G.first;
        ok(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
        ok(cudaEventCreate(&G.k_end));
        s->groups.push_back(G);
    }
    ok(cudaMalloc(&s->d_desc, sizeof(RepDesc) * s->R));
    ok(cudaMalloc(&s->d_state, sizeof(RepState) * s->R));
    ok(cudaMalloc(&s->d_ldesc, sizeof(LegDesc) * s->R));
    ok(cudaMalloc(&s->d_lstate, sizeof(LegState) * s->R));
    ok(cudaMalloc(&s->d_returns, sizeof(int64_t) * s->R));
    ok(cudaMalloc(&s->d_chunk_ptrs, sizeof(rlgs_row *) * MAX_CHUNKS));
    ok(cudaMallocHost(&s->h_returns, sizeof(int64_t) * s->R));
    ok(cudaMallocHost(&s->h_state, sizeof(RepState) * s->R));
    ok(cudaMallocHost(&s->h_init, sizeof(RepState) * s->R));
    ok(cudaMallocHost(&s->h_lstate, sizeof(LegState) * s->R));
    ok(cudaMallocHost(&s->h_linit, sizeof(LegState) * s->R));
    if (s->pack) {
        ok(cudaMalloc(&s->d_pdesc, sizeof(PackDesc) * s->R));
        ok(cudaMalloc(&s->d_pstate, sizeof(PackState) * s->R));
        ok(cudaMallocHost(&s->h_pstate, sizeof(PackState) * s->R));
        ok(cudaMallocHost(&s->h_pinit, sizeof(PackState) * s->R));
    }
    if (ce != cudaSuccess) {
        rlgs_destroy(s);
        return fail(ce == cudaErrorMemoryAllocation ? RLGS_ERR_OOM : RLGS_ERR_CUDA, "rlgs_create: %s", cudaGetErrorString(ce));
    }
    *out = s;
    return RLGS_OK;
}

extern "C" void rlgs_destroy(rlgs_sim *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    for (auto &t : s->traces) { cudaFree(t.dev); cudaFree(t.net); cudaFree(t.dur_out); cudaFree(t.pack); cudaFree(t.feat); cudaFree(t.pack_slab); cudaFree(t.slab); }
    cudaFree(s->d_pdesc); cudaFree(s->d_pstate);
    if (s->h_pstate) cudaFreeHost(s->h_pstate);
    if (s->h_pinit) cudaFreeHost(s->h_pinit);
    for (auto p : s->d_chunks) cudaFree(p);
    for (auto p : s->h_chunks) if (p) cudaFreeHost(p);
    cudaFree(s->d_desc); cudaFree(s->d_state); cudaFree(s->d_ldesc); cudaFree(s->d_lstate); cudaFree(s->d_jobs);
    cudaFree(s->d_returns); cudaFree(s->d_chunk_ptrs);
    free_host_jobs(s);
    if (s->h_returns) cudaFreeHost(s->h_returns);
    if (s->h_state) cudaFreeHost(s->h_state);
    if (s->h_init) cudaFreeHost(s->h_init);
    if (s->h_lstate) cudaFreeHost(s->h_lstate);
    if (s->h_linit) cudaFreeHost(s->h_linit);
    for (auto &G : s->groups) {
        if (G.k_end) cudaEventDestroy(G.k_end);
        if (G.stream) cudaStreamDestroy(G.stream);
    }
    for (auto e : s->ev_pool) cudaEventDestroy(e);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int32_t rlgs_set_stream(rlgs_sim *s, void *cuda_stream, int32_t enable) {
    if (!s) return fail(RLGS_ERR_BAD_ARG, "null handle");
    s->user_stream = cuda_stream;
    s->use_user_stream = enable != 0;
    return RLGS_OK;
}

extern "C" int32_t rlgs_load_trace(rlgs_sim *s, int32_t first, int32_t count, const rlgs_job *jobs, int32_t n,
                                   const rlgs_netcost_inputs *net) {
    if (!s || !jobs) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (n < 1) return fail(RLGS_ERR_BAD_ARG, "trace has no jobs (the reference asserts on an empty job table, log_manager.py:138)");
    if (first < 0 || count < 1 || first + count > s->R) return fail(RLGS_ERR_BAD_ARG, "replica range [%d,%d) out of 0..%d", first, first + count, s->R);
    if (s->opts.enable_network_costs && (!net || !net->duration || !net->model_mb || !net->iterations))
        return fail(RLGS_ERR_BAD_ARG, "enable_network_costs needs duration / model_mb / iterations arrays");
    if (s->wire && n >= (1 << 20)) return fail(RLGS_ERR_WIRE, "a wire row counts at most 2^20 - 1 jobs: use RLGS_ROWFMT_WIDE");
    CU(cudaSetDevice(s->device));
    TraceBuf tb;
    tb.n = n;
    int prev = 0;
    for (int32_t i = 0; i < n; ++i) {
        const rlgs_job &j = jobs[i];
        if (j.tasks < 1 || j.gpus_per_task < 1 || j.gpus < 1)
            return fail(RLGS_ERR_BAD_ARG, "job %d: gpus/tasks/gpus_per_task must be >= 1 (the reference raises on such rows)", i);
        if ((int)j.tasks * j.gpus_per_task > (int)j.gpus) return fail(RLGS_ERR_BAD_ARG, "job %d: tasks*gpus_per_task > gpus", i);
        if (j.arrival_tick < prev) return fail(RLGS_ERR_BAD_ARG, "job %d: arrival ticks must be non-decreasing", i);
        if (j.dur_ticks < 1) return fail(RLGS_ERR_BAD_ARG, "job %d: dur_ticks must be >= 1", i);
        if (j.index != i) return fail(RLGS_ERR_BAD_ARG, "job %d: index field must equal the position", i);
        if (j.tasks > 32767) return fail(RLGS_ERR_BAD_ARG, "job %d: more than 32767 tasks", i);
        prev = j.arrival_tick;
        tb.log_cap += std::min<int64_t>(j.tasks, s->cc.N);
    }
    tb.max_arrival = prev; tb.first = first; tb.count = count;
    // a reload of the same replica range with a trace that fits the existing buffers only re-uploads
    // the records (the e2e path: one host->device copy per step, no allocation)
    for (size_t t = 0; t < s->traces.size() && !s->pack; ++t) {
        TraceBuf &old = s->traces[t];
        bool attached = old.dev != nullptr && old.first == first && old.count == count;
        for (int r = 0; r < count && attached; ++r) attached = s->rep_trace[first + r] == (int)t;   // a later load may have re-pointed some of them
        if (attached && n <= old.cap_n && tb.log_cap <= old.cap_log) {
            CU(cudaMemcpy(old.dev, jobs, sizeof(rlgs_job) * (size_t)n, cudaMemcpyHostToDevice));
            if (s->opts.enable_network_costs) {
                CU(cudaMemcpy(old.net, net->duration, 8 * (size_t)n, cudaMemcpyHostToDevice));
                CU(cudaMemcpy(old.net + n, net->model_mb, 8 * (size_t)n, cudaMemcpyHostToDevice));
                CU(cudaMemcpy(old.net + 2 * (size_t)n, net->iterations, 8 * (size_t)n, cudaMemcpyHostToDevice));
                for (int r = 0; r < count; ++r) s->h_desc[first + r].dur_out = old.dur_out + (size_t)r * n;
            }
            old.n = n; old.log_cap = tb.log_cap; old.max_arrival = tb.max_arrival;
            if (s->wire || s->opts.fetch_jobs == 2) old.host.assign(jobs, jobs + n);
            s->env_ready = false; s->xp_replica = -1;
            for (int r = 0; r < count; ++r) {
                s->h_desc[first + r].J = n; s->h_desc[first + r].log_cap = (int32_t)std::min<int64_t>(tb.log_cap, 0x7fffffff);
                s->h_ldesc[first + r].J = n;
            }
            s->ran = false;
            return RLGS_OK;
        }
    }
    tb.cap_n = n; tb.cap_log = tb.log_cap;
    if (s->wire || s->opts.fetch_jobs == 2) tb.host.assign(jobs, jobs + n);
    s->env_ready = false; s->xp_replica = -1;
    CU(cudaMalloc(&tb.dev, sizeof(rlgs_job) * (size_t)n));
    cudaError_t e = cudaMemcpy(tb.dev, jobs, sizeof(rlgs_job) * (size_t)n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(tb.dev); return fail(RLGS_ERR_CUDA, "trace upload: %s", cudaGetErrorString(e)); }
    if (s->opts.enable_network_costs) {
        CU(cudaMalloc(&tb.net, 8 * 3 * (size_t)n));
        CU(cudaMalloc(&tb.dur_out, 8 * (size_t)n * (size_t)count));
        CU(cudaMemcpy(tb.net, net->duration, 8 * (size_t)n, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(tb.net + n, net->model_mb, 8 * (size_t)n, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(tb.net + 2 * (size_t)n, net->iterations, 8 * (size_t)n, cudaMemcpyHostToDevice));
    }
    // the replicas of this range leave their previous trace buffers; a buffer nobody points at any more is released
    for (int r = 0; r < count; ++r) s->rep_trace[first + r] = -1;
    int tid = -1;
    for (size_t t = 0; t < s->traces.size(); ++t) {
        TraceBuf &old = s->traces[t];
        if (old.dev) {
            bool used = false;
            for (int r = 0; r < s->R && !used; ++r) used = s->rep_trace[r] == (int)t;
            if (used) continue;
            cudaFree(old.dev); cudaFree(old.net); cudaFree(old.dur_out); cudaFree(old.pack); cudaFree(old.feat); cudaFree(old.pack_slab); cudaFree(old.slab);
            old = TraceBuf();
        }
        if (tid < 0) tid = (int)t;
    }
    if (tid < 0) { tid = (int)s->traces.size(); s->traces.push_back(tb); } else s->traces[tid] = tb;
    unsigned char *slab = nullptr;
    if (s->pack) {
        for (int r = 0; r < count; ++r) { s->h_pdesc[first + r] = PackDesc{}; s->h_pdesc[first + r].trace = tb.dev; s->h_pdesc[first + r].J = n; s->h_ldesc[first + r].J = n; }
    } else if (!s->legacy) {
        // per-replica working set: queue stack | placement log | node_save | slot_save
        int nw = 3 * s->cc.N + (s->cc.N + 31) / 32;
        size_t a0 = align_up(sizeof(rlgs_job) * (size_t)n, 256), a1 = align_up(sizeof(int2) * (size_t)std::max<int64_t>(tb.log_cap, 1), 256);
        size_t a2 = align_up(4 * (size_t)nw, 256), a3 = align_up(sizeof(int4) * 2 * (size_t)s->slot_cap, 256);
        size_t per = a0 + a1 + a2 + a3;
        CU(cudaMalloc(&slab, per * (size_t)count));
        for (int r = 0; r < count; ++r) {
            unsigned char *p = slab + per * (size_t)r;
            RepDesc &D = s->h_desc[first + r];
            D.trace = tb.dev; D.J = n; D.log_cap = (int32_t)std::min<int64_t>(tb.log_cap, 0x7fffffff);
            D.stack = reinterpret_cast<rlgs_job *>(p); p += a0;
            D.place_log = reinterpret_cast<int2 *>(p); p += a1;
            D.node_save = reinterpret_cast<int32_t *>(p); p += a2;
            D.slot_save = reinterpret_cast<int4 *>(p);
            D.net_in = tb.net; D.dur_out = tb.dur_out ? tb.dur_out + (size_t)r * n : nullptr;
        }
    } else {
        // per-replica working set: 2 entry buffers | pending scratch | 2 demotion scratches | end list | placement scratch
        size_t ae = align_up(sizeof(Ent) * (size_t)n, 256), al = align_up(4 * (size_t)n, 256), ap = align_up(sizeof(int2) * (size_t)s->cc.N, 256);
        bool dlas = s->opts.schedule == RLGS_SCHED_DLAS_GPU || s->opts.schedule == RLGS_SCHED_DLAS;
        const bool sorts = s->lp.sort_mode != 0;   // shortest / shortest-gpu need one scratch buffer for the stable sort
        size_t per = (dlas ? 5 : (sorts ? 2 : 1)) * ae + al + ap;
        CU(cudaMalloc(&slab, per * (size_t)count));
        for (int r = 0; r < count; ++r) {
            unsigned char *p = slab + per * (size_t)r;
            LegDesc &D = s->h_ldesc[first + r];
            D.trace = tb.dev; D.J = n; D.cap = n;
            D.buf[0] = reinterpret_cast<Ent *>(p); p += ae;
            if (dlas) {
                D.buf[1] = reinterpret_cast<Ent *>(p); p += ae;
                D.scratch_p = reinterpret_cast<Ent *>(p); p += ae;
                D.scratch_d[0] = reinterpret_cast<Ent *>(p); p += ae;
                D.scratch_d[1] = reinterpret_cast<Ent *>(p); p += ae;
            } else {
                D.buf[1] = D.scratch_p = D.scratch_d[0] = D.scratch_d[1] = nullptr;
                if (sorts) { D.scratch_p = reinterpret_cast<Ent *>(p); p += ae; }
            }
            D.end_list = reinterpret_cast<int32_t *>(p); p += al;
            D.place_scratch = reinterpret_cast<int2 *>(p);
        }
    }
    s->traces[tid].slab = slab;
    for (int r = 0; r < count; ++r) s->rep_trace[first + r] = tid;
    s->ran = false;
    return RLGS_OK;
}

extern "C" int32_t rlgs_load_pack_inputs(rlgs_sim *s, int32_t first, int32_t count, const rlgs_pack_inputs *in, int32_t n) {
    if (!s || !in || !in->util_avg || !in->util_sd || !in->task_mem || !in->heap_cap) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (!s->pack) return fail(RLGS_ERR_STATE, "pack inputs belong to the horus placement");
    if (in->mem_shift < 0 || in->mem_shift > 30 || in->gpu_mem_cap_mib < 1) return fail(RLGS_ERR_BAD_ARG, "mem_shift must be 0..30 and gpu_mem_cap_mib >= 1");
    TraceBuf *tb = nullptr;
    for (auto &t : s->traces) if (t.first == first && t.count == count && t.n == n && !t.pack) tb = &t;
    if (!tb) return fail(RLGS_ERR_STATE, "no trace of %d jobs loaded for replicas [%d,%d) without pack inputs", n, first, first + count);
    CU(cudaSetDevice(s->device));
    std::vector<rlgs_job> jobs((size_t)n);
    CU(cudaMemcpy(jobs.data(), tb->dev, sizeof(rlgs_job) * (size_t)n, cudaMemcpyDeviceToHost));
    std::vector<PackJob> pj((size_t)n);
    int64_t sum_tasks = 0;
    for (int i = 0; i < n; ++i) {
        if (jobs[i].tasks > PACK_MAX_TASKS) return fail(RLGS_ERR_UNSUPPORTED, "job %d has %d tasks; the pack placement handles up to %d per job", i, (int)jobs[i].tasks, PACK_MAX_TASKS);
        if (in->heap_cap[i] < 0 || in->heap_cap[i] > PACK_MAX_HEAP) return fail(RLGS_ERR_UNSUPPORTED, "job %d: used_gpus %d outside 0..%d", i, in->heap_cap[i], PACK_MAX_HEAP);
        if (in->task_mem[i] < 0) return fail(RLGS_ERR_BAD_ARG, "job %d: negative memory", i);
        pj[i].util_avg = in->util_avg[i]; pj[i].util_sd = in->util_sd[i]; pj[i].mem = in->task_mem[i];
        pj[i].heap_cap = in->heap_cap[i]; pj[i].task_off = (int32_t)sum_tasks;
        sum_tasks += jobs[i].tasks;
    }
    CU(cudaMalloc(&tb->pack, sizeof(PackJob) * (size_t)n));
    CU(cudaMemcpy(tb->pack, pj.data(), sizeof(PackJob) * (size_t)n, cudaMemcpyHostToDevice));
    const int K = s->pp.plus_k;
    if (K > 0) {
        if (!in->util_max || !in->mem_avg_mib || !in->used_gpus) return fail(RLGS_ERR_BAD_ARG, "horus+ needs util_max, mem_avg_mib and used_gpus (the k-means features of core/jobs/utils.py:4-22)");
        std::vector<PlusFeat> ft((size_t)n);
        const double unit_mib = 1.0 / (double)((int64_t)1 << in->mem_shift);
        for (int i = 0; i < n; ++i) {
            PlusFeat &f = ft[i];
            f.f[0] = (double)jobs[i].tasks; f.f[1] = in->util_avg[i]; f.f[2] = (double)jobs[i].gpus_per_task; f.f[3] = in->used_gpus[i];
            f.f[4] = in->util_max[i]; f.f[5] = in->mem_avg_mib[i]; f.f[6] = (double)in->task_mem[i] * unit_mib;
            double t = f.f[0]; for (int k = 1; k < 7; ++k) t += f.f[k];   // transform_to_dist: left to right
            f.tdist = t;
        }
        CU(cudaMalloc(&tb->feat, sizeof(PlusFeat) * (size_t)n));
        CU(cudaMemcpy(tb->feat, ft.data(), sizeof(PlusFeat) * (size_t)n, cudaMemcpyHostToDevice));
    }
    const size_t KQ = (size_t)std::max(K, 1);
    const size_t N = (size_t)s->cc.N, Dv = (size_t)s->cc.D, J = (size_t)n, W = (N + 31) / 32;
    // per-replica working set, 256-byte aligned pieces in this order
    const size_t sz[] = {4 * N, 4 * N, 4 * N, 4 * Dv, 8 * Dv, 8 * PACK_DEV_SLOTS * Dv, 4 * J * W, 8 * J * KQ, 4 * J * KQ, 4 * J, 4 * J, 4 * J, 4 * J,
                         4 * PACK_CAL_W, 2 * (size_t)std::max<int64_t>(sum_tasks, 1), 4 * J,
                         4 * J, 4 * J, 4 * J, 4 * J, 4 * J, 4 * J, 4 * J, 4 * J, 4 * PACK_CAL_W, 4 * J,
                         4 * N, 4 * N, 4 * W, 8 * (size_t)std::max<int64_t>(sum_tasks, 1), 4 * J,
                         4 * J, 4 * J, 4 * J, 8 * J};
    size_t per = 0;
    for (size_t v : sz) per += align_up(v, 256);
    unsigned char *slab = nullptr;
    CU(cudaMalloc(&slab, per * (size_t)count));
    tb->pack_slab = slab;
    const int64_t unit = (int64_t)1 << in->mem_shift;
    for (int r = 0; r < count; ++r) {
        unsigned char *p = slab + per * (size_t)r;
        PackDesc &D = s->h_pdesc[first + r];
        int k = 0;
        auto take = [&](size_t) { unsigned char *q = p; p += align_up(sz[k++], 256); return q; };
        D.trace = tb->dev; D.pj = tb->pack; D.J = n; D.W = (int32_t)W;
        D.units = (int32_t *)take(0); D.ntk = (int32_t *)take(0); D.npj = (int32_t *)take(0); D.dn = (int32_t *)take(0);
        D.dm = (int64_t *)take(0); D.ent = (int2 *)take(0); D.pjbits = (uint32_t *)take(0); D.qkey = (double *)take(0);
        D.qjob = (int32_t *)take(0); D.lprev = (int32_t *)take(0); D.lnext = (int32_t *)take(0); D.pend = (int32_t *)take(0);
        D.cnext = (int32_t *)take(0); D.chead = (int32_t *)take(0); D.tnode = (int16_t *)take(0); D.fin = (int32_t *)take(0);
        D.imask = (uint32_t *)take(0); D.bmask = (uint32_t *)take(0); D.jflag = (int32_t *)take(0); D.pproc = (int32_t *)take(0);
        D.nstart = (int32_t *)take(0); D.qtick = (int32_t *)take(0); D.cbk = (int32_t *)take(0); D.snext = (int32_t *)take(0);
        D.shead = (int32_t *)take(0); D.sat = (int32_t *)take(0);
        D.ybusy = (uint32_t *)take(0); D.ykey = (uint32_t *)take(0); D.yever = (uint32_t *)take(0); D.plog = (int2 *)take(0); D.pcnt = (int32_t *)take(0);
        D.kjobs = (int32_t *)take(0); D.kassign = (int32_t *)take(0); D.kold = (int32_t *)take(0); D.ktmp = (double *)take(0); D.feat = tb->feat;
        D.cap_units = (int64_t)in->gpu_mem_cap_mib * unit; D.margin_units = 500 * unit;
        D.cap_mib = (double)in->gpu_mem_cap_mib; D.unit_mib = 1.0 / (double)unit;
    }
    s->ran = false;
    return RLGS_OK;
}

// (re)allocates the [N_PLANES][R][Jmax] job-output arrays and points every replica at its rows
static int32_t setup_job_arrays(rlgs_sim *s) {
    int32_t Jmax = 0;
    for (int r = 0; r < s->R; ++r) {
        if (s->rep_trace[r] < 0) return fail(RLGS_ERR_STATE, "replica %d has no trace (call rlgs_load_trace)", r);
        Jmax = std::max(Jmax, s->legacy ? s->h_ldesc[r].J : s->h_desc[r].J);
    }
    if (Jmax != s->Jmax || !s->d_jobs) {
        cudaFree(s->d_jobs); s->d_jobs = nullptr;
        free_host_jobs(s);
        s->Jmax = Jmax;
        s->jobs_bytes = sizeof(int32_t) * N_PLANES * (size_t)s->R * (size_t)Jmax;
        CU(cudaMalloc(&s->d_jobs, s->jobs_bytes));
    }
    size_t plane = (size_t)s->R * (size_t)Jmax;
    for (int r = 0; r < s->R; ++r) {
        if (!s->legacy) {
            RepDesc &D = s->h_desc[r];
            D.start_tick = s->d_jobs + 0 * plane + (size_t)r * Jmax;
            D.end_tick = s->d_jobs + 1 * plane + (size_t)r * Jmax;
            D.finish_order = s->d_jobs + 2 * plane + (size_t)r * Jmax;
            D.place_off = s->d_jobs + 3 * plane + (size_t)r * Jmax;
        } else {
            for (int k = 0; k < N_PLANES; ++k) s->h_ldesc[r].planes[k] = s->d_jobs + k * plane + (size_t)r * Jmax;
            if (s->pack) {
                if (!s->h_pdesc[r].pj) return fail(RLGS_ERR_STATE, "replica %d has no pack inputs (call rlgs_load_pack_inputs)", r);
                for (int k = 0; k < N_PLANES; ++k) s->h_pdesc[r].planes[k] = s->h_ldesc[r].planes[k];
            }
        }
    }
    return RLGS_OK;
}

static int32_t add_chunks(rlgs_sim *s, int upto, bool host_too) {
    if (upto > MAX_CHUNKS) return fail(RLGS_ERR_CAPACITY, "row store would exceed %d chunks", MAX_CHUNKS);
    bool grew = false;
    while ((int)s->d_chunks.size() < upto) {
        rlgs_row *p = nullptr;
        CU(cudaMalloc(&p, chunk_bytes(s)));
        s->d_chunks.push_back(p); s->h_chunks.push_back(nullptr); s->h_chunk_valid.push_back(0);
        grew = true;
    }
    if (host_too)
        for (int k = 0; k < upto; ++k)
            if (!s->h_chunks[k]) CU(cudaMallocHost(&s->h_chunks[k], chunk_bytes(s)));
    if (grew) CU(cudaMemcpy(s->d_chunk_ptrs, s->d_chunks.data(), sizeof(rlgs_row *) * s->d_chunks.size(), cudaMemcpyHostToDevice));
    return RLGS_OK;
}

static NetCost netcost_of(const rlgs_sim *s) {
    NetCost n; n.enabled = s->opts.enable_network_costs != 0; n.pad = 0; n.bandwidth = s->opts.bandwidth; n.latency = s->opts.internode_latency;
    return n;
}

// one instantiation of the fifo tick loop: LPR lanes per replica, packed / split node words, env / rows / network-cost variants
template <int LPR, bool PK, bool ENV, int ROWS, bool NET>
static cudaError_t launch_grp(rlgs_sim *s, int first, int count, int budget, const RowStore &rs, const EnvIO &io, cudaStream_t st) {
    constexpr int K = 32 / LPR;
    const size_t smem = K * grp_smem_bytes(s->cc.N, s->cc.G, s->slot_cap, LPR);
    static size_t attr_set[64] = {0};   // per device: largest dynamic shared-memory size already allowed for this instantiation
    const int dev = s->device & 63;
    if (smem > attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fifo_grp_kernel<LPR, PK, ENV, ROWS, NET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set[dev] = smem;
    }
    fifo_grp_kernel<LPR, PK, ENV, ROWS, NET><<<(count + K - 1) / K, 32, smem, st>>>(s->d_desc + first, s->d_state + first, count, s->cc, s->slot_cap, budget, rs,
                                                                                s->d_returns + first, s->opts.max_ticks, io, netcost_of(s));
    return cudaGetLastError();
}

// rows: 0 = none, 1 = rlgs_row, 2 = rlgs_row16, 3 = rlgs_row12, 4 = rlgs_row16e, 5 = rlgs_row4e (no network costs).  The network-cost and env-with-rows variants exist for one replica per warp only
// (rlgs_create pins lanes_per_replica to 32 for them).
template <int LPR, bool PK>
static cudaError_t launch_fifo_lp(rlgs_sim *s, int first, int count, int budget, int rows, bool env, const EnvIO &io, const RowStore &rs, cudaStream_t st) {
    if (s->opts.enable_network_costs != 0) {
        if (LPR != 32) return cudaErrorNotSupported;
        if (env) return rows ? launch_grp<32, PK, true, 1, true>(s, first, count, budget, rs, io, st) : launch_grp<32, PK, true, 0, true>(s, first, count, budget, rs, io, st);
        if (rows == 0) return launch_grp<32, PK, false, 0, true>(s, first, count, budget, rs, io, st);
        if (rows == 1) return launch_grp<32, PK, false, 1, true>(s, first, count, budget, rs, io, st);
        if (rows == 2) return launch_grp<32, PK, false, 2, true>(s, first, count, budget, rs, io, st);
        if (rows == 3) return launch_grp<32, PK, false, 3, true>(s, first, count, budget, rs, io, st);
        if (rows == 5) return cudaErrorNotSupported;
        return launch_grp<32, PK, false, 4, true>(s, first, count, budget, rs, io, st);
    }
    if (env && rows) return LPR == 32 && !s->wire ? launch_grp<32, PK, true, 1, false>(s, first, count, budget, rs, io, st) : cudaErrorNotSupported;
    if (env) return launch_grp<LPR, PK, true, 0, false>(s, first, count, budget, rs, io, st);
    if (rows == 0) return launch_grp<LPR, PK, false, 0, false>(s, first, count, budget, rs, io, st);
    if (rows == 1) return launch_grp<LPR, PK, false, 1, false>(s, first, count, budget, rs, io, st);
    if (rows == 2) return launch_grp<LPR, PK, false, 2, false>(s, first, count, budget, rs, io, st);
    if (rows == 3) return launch_grp<LPR, PK, false, 3, false>(s, first, count, budget, rs, io, st);
    if (rows == 5) return launch_grp<LPR, PK, false, 5, false>(s, first, count, budget, rs, io, st);
    return launch_grp<LPR, PK, false, 4, false>(s, first, count, budget, rs, io, st);
}

static cudaError_t launch_fifo(rlgs_sim *s, int first, int count, int budget, int rows, bool env, const EnvIO &io, const RowStore &rs, cudaStream_t st) {
    const bool pk = grp_packed(s->cc.G);
    if (s->lpr == 8) return pk ? launch_fifo_lp<8, true>(s, first, count, budget, rows, env, io, rs, st) : launch_fifo_lp<8, false>(s, first, count, budget, rows, env, io, rs, st);
    if (s->lpr == 16) return pk ? launch_fifo_lp<16, true>(s, first, count, budget, rows, env, io, rs, st) : launch_fifo_lp<16, false>(s, first, count, budget, rows, env, io, rs, st);
    return pk ? launch_fifo_lp<32, true>(s, first, count, budget, rows, env, io, rs, st) : launch_fifo_lp<32, false>(s, first, count, budget, rows, env, io, rs, st);
}

static void launch(rlgs_sim *s, int first, int count, int budget, bool rows, cudaStream_t st) {
    RowStore rs; rs.chunks = rows ? s->d_chunk_ptrs : nullptr; rs.n_chunks = (int)s->d_chunks.size(); rs.replica = first;
    if (!s->legacy) {
        EnvIO none; memset(&none, 0, sizeof none);
        launch_fifo(s, first, count, budget, rows ? 1 + s->wire : 0, false, none, rs, st);   // the caller reads cudaGetLastError
    } else {
        LegParams lp = s->lp; lp.event_budget = budget;
        const bool dense = s->R > 148 * 24;   // more replicas than 24 warps per SM: the 64-register builds of the per-event kernels (legacy_sched.cuh)
        if (s->pack) {
            PackParams pp = s->pp; pp.tick_budget = budget;
#define RLGS_LAUNCH_PACK(G, Y, PL) pack_horus_kernel<G, Y, PL><<<count, 32, pack_smem_bytes(s->cc.N), st>>>(s->d_pdesc + first, s->d_pstate + first, pp, s->cc, rs, s->d_returns + first)
            const bool yarn = s->opts.placement == RLGS_PLACE_YARN;
            if (pp.plus_k > 0) { if (yarn) RLGS_LAUNCH_PACK(false, true, true); else RLGS_LAUNCH_PACK(false, false, true); }
            else if (pp.gandiva) { if (yarn) RLGS_LAUNCH_PACK(true, true, false); else RLGS_LAUNCH_PACK(true, false, false); }
            else { if (yarn) RLGS_LAUNCH_PACK(false, true, false); else RLGS_LAUNCH_PACK(false, false, false); }
#undef RLGS_LAUNCH_PACK
        } else if (s->opts.schedule == RLGS_SCHED_DLAS_GPU || s->opts.schedule == RLGS_SCHED_DLAS) {
            if (dense) dlas_gpu_kernel<32><<<count, 32, 0, st>>>(s->d_ldesc + first, s->d_lstate + first, lp, rs, s->d_returns + first);
            else dlas_gpu_kernel<20><<<count, 32, 0, st>>>(s->d_ldesc + first, s->d_lstate + first, lp, rs, s->d_returns + first);
        } else {
            if (dense) sjf_yarn_kernel<32><<<count, 32, sjf_smem_bytes(s->cc.N), st>>>(s->d_ldesc + first, s->d_lstate + first, lp, s->cc, rs, s->d_returns + first);
            else sjf_yarn_kernel<16><<<count, 32, sjf_smem_bytes(s->cc.N), st>>>(s->d_ldesc + first, s->d_lstate + first, lp, s->cc, rs, s->d_returns + first);
        }
    }
}

static Progress progress_of(const rlgs_sim *s, int r) {
    Progress p;
    if (s->pack) { p.rows = s->h_pstate[r].d; p.done = s->h_pstate[r].done; p.status = s->h_pstate[r].status; }
    else if (!s->legacy) { p.rows = s->h_state[r].d; p.done = s->h_state[r].done; p.status = s->h_state[r].status; }
    else { p.rows = s->h_lstate[r].n_rows; p.done = s->h_lstate[r].done; p.status = s->h_lstate[r].status; }
    return p;
}

static cudaError_t get_event(rlgs_sim *s, size_t i, cudaEvent_t *out) {
    while (s->ev_pool.size() <= i) { cudaEvent_t e; cudaError_t ce = cudaEventCreate(&e); if (ce != cudaSuccess) return ce; s->ev_pool.push_back(e); }
    *out = s->ev_pool[i];
    return cudaSuccess;
}

extern "C" int32_t rlgs_run(rlgs_sim *s) {
    if (!s) return fail(RLGS_ERR_BAD_ARG, "null handle");
    CU(cudaSetDevice(s->device));
    int32_t rc = setup_job_arrays(s);
    if (rc) return rc;
    cudaStream_t main_st = s->use_user_stream ? (cudaStream_t)s->user_stream : s->stream;
    const int mode = s->opts.rows_mode;
    const bool rows = mode != RLGS_ROWS_NONE, eager_rows = mode == RLGS_ROWS_FULL, eager_jobs = s->opts.fetch_jobs != 0;
    const int R = s->R;
    if (!s->legacy) {
        size_t smem = (32 / s->lpr) * grp_smem_bytes(s->cc.N, s->cc.G, s->slot_cap, s->lpr);
        if (smem > 227 * 1024) return fail(RLGS_ERR_CAPACITY, "cluster state needs %zu B of shared memory per warp (> 227 KB)", smem);
    } else if (!s->pack && s->lp.nq == 1 && s->opts.schedule != RLGS_SCHED_DLAS_GPU && s->opts.schedule != RLGS_SCHED_DLAS) {
        CU(cudaFuncSetAttribute(sjf_yarn_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sjf_smem_bytes(s->cc.N)));
        CU(cudaFuncSetAttribute(sjf_yarn_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sjf_smem_bytes(s->cc.N)));
    }
    int32_t max_arrival = 0;
    for (int r = 0; r < R; ++r) {
        if (!s->legacy) {
            RepState z; memset(&z, 0, sizeof z);
            z.head = s->h_desc[r].J; z.idle_nodes = s->cc.N;
            z.n_free_nodes = (s->cc.cpu_cap > 0 || s->cc.mem_cap > 0) ? s->cc.N : 0;
            z.free_hint = -1;
            s->h_init[r] = z;
        } else {
            LegState z; memset(&z, 0, sizeof z);
            z.next_end = RLGS_NEVER; z.next_jump = RLGS_NEVER;
            s->h_linit[r] = z;
            if (s->pack) {
                PackState q; memset(&q, 0, sizeof q);
                q.idle_nodes = s->cc.N; q.n_free_nodes = (s->cc.cpu_cap > 0 || s->cc.mem_cap > 0) ? s->cc.N : 0;
                q.lhead = q.ltail = q.mlo = -1;
                s->h_pinit[r] = q;
            }
        }
        max_arrival = std::max(max_arrival, s->traces[s->rep_trace[r]].max_arrival);
    }
    int target_chunks = 0;   // chunks launched (and copied) before the host looks at the replica states
    if (rows) {
        int64_t cap = s->opts.rows_cap > 0 ? s->opts.rows_cap : (s->rows_hint > 0 ? s->rows_hint : (int64_t)max_arrival + 4096);
        target_chunks = (int)std::max<int64_t>(1, (cap + RLGS_ROW_CHUNK - 1) / RLGS_ROW_CHUNK);
        rc = add_chunks(s, std::max(target_chunks, (int)s->d_chunks.size()), eager_rows);
        if (rc) return rc;
    }
    if (eager_jobs) { rc = alloc_host_jobs(s, true); if (rc) return rc; }
    std::fill(s->h_chunk_valid.begin(), s->h_chunk_valid.end(), 0);
    s->planes_mask = 0; std::fill(s->derived.begin(), s->derived.end(), 0); s->ran = false; s->xp_replica = -1;

    if (!s->legacy) {
        CU(cudaMemcpyAsync(s->d_state, s->h_init, sizeof(RepState) * R, cudaMemcpyHostToDevice, main_st));
        CU(cudaMemcpyAsync(s->d_desc, s->h_desc.data(), sizeof(RepDesc) * R, cudaMemcpyHostToDevice, main_st));
    } else if (s->pack) {
        CU(cudaMemcpyAsync(s->d_pstate, s->h_pinit, sizeof(PackState) * R, cudaMemcpyHostToDevice, main_st));
        CU(cudaMemcpyAsync(s->d_pdesc, s->h_pdesc.data(), sizeof(PackDesc) * R, cudaMemcpyHostToDevice, main_st));
    } else {
        CU(cudaMemcpyAsync(s->d_lstate, s->h_linit, sizeof(LegState) * R, cudaMemcpyHostToDevice, main_st));
        CU(cudaMemcpyAsync(s->d_ldesc, s->h_ldesc.data(), sizeof(LegDesc) * R, cudaMemcpyHostToDevice, main_st));
    }
    CU(cudaMemsetAsync(s->d_jobs, 0xff, s->jobs_bytes, main_st));

    const bool pipelined = eager_rows && s->opts.ticks_per_launch == 0;
    const int budget = pipelined ? RLGS_ROW_CHUNK : (s->opts.ticks_per_launch > 0 ? s->opts.ticks_per_launch : (1 << 30));
    float total_ms = 0.f;
    int launches = 0;
    int next_chunk = 0;   // pipelined: first chunk whose rows have not been sent to the host yet
    for (;;) {
        size_t ev_i = 0;
        cudaEvent_t e_begin = nullptr, e_last = nullptr;
        if (pipelined) {
            // launches bounded to one row chunk; chunk k travels to the pinned host mirror on the copy stream
            // while chunk k+1 is being simulated.  The host does not wait until everything is enqueued.
            CU(get_event(s, ev_i++, &e_begin));
            CU(cudaEventRecord(e_begin, main_st));
            for (; next_chunk < target_chunks; ++next_chunk) {
                launch(s, 0, R, budget, true, main_st);
                CU(cudaGetLastError());
                launches++;
                CU(get_event(s, ev_i++, &e_last));
                CU(cudaEventRecord(e_last, main_st));
                CU(cudaStreamWaitEvent(s->copy_stream, e_last, 0));
                CU(cudaMemcpyAsync(s->h_chunks[next_chunk], s->d_chunks[next_chunk], chunk_bytes(s), cudaMemcpyDeviceToHost, s->copy_stream));
                s->h_chunk_valid[next_chunk] = 1;
            }
        } else {
            CU(cudaEventRecord(s->ev_fork, main_st));
            for (auto &G : s->groups) {
                CU(cudaStreamWaitEvent(G.stream, s->ev_fork, 0));
                launch(s, G.first, G.count, budget, rows, G.stream);
                CU(cudaGetLastError());
                launches++;
                CU(cudaEventRecord(G.k_end, G.stream));
                CU(cudaStreamWaitEvent(main_st, G.k_end, 0));
            }
        }
        if (s->pack) CU(cudaMemcpyAsync(s->h_pstate, s->d_pstate, sizeof(PackState) * R, cudaMemcpyDeviceToHost, main_st));
        else if (!s->legacy) CU(cudaMemcpyAsync(s->h_state, s->d_state, sizeof(RepState) * R, cudaMemcpyDeviceToHost, main_st));
        else CU(cudaMemcpyAsync(s->h_lstate, s->d_lstate, sizeof(LegState) * R, cudaMemcpyDeviceToHost, main_st));
        CU(cudaStreamSynchronize(main_st));
        if (pipelined) {
            float ms = 0.f;
            if (e_last) CU(cudaEventElapsedTime(&ms, e_begin, e_last));
            total_ms += ms;
        } else {
            float wave = 0.f;
            for (auto &G : s->groups) { float t = 0.f; CU(cudaEventElapsedTime(&t, s->ev_fork, G.k_end)); wave = std::max(wave, t); }
            total_ms += wave;
        }
        bool all_done = true, overflow = false, rows_full = false;
        int bad = -1, bad_code = 0;
        for (int r = 0; r < R; ++r) {
            Progress p = progress_of(s, r);
            if (p.status != RLGS_OK && !overflow) { overflow = true; bad = r; bad_code = p.status; }
            if (!p.done) { all_done = false; if (rows && p.rows >= (int64_t)s->d_chunks.size() * RLGS_ROW_CHUNK) rows_full = true; }
        }
        if (overflow) {
            cudaStreamSynchronize(s->copy_stream);
            if (bad_code == RLGS_ERR_SLOTS) return fail(RLGS_ERR_SLOTS, "replica %d: running-job slot table overflow at slot_cap=%d: recreate with a larger opts.slot_cap", bad, s->slot_cap);
            if (bad_code == RLGS_ERR_WIRE) return fail(RLGS_ERR_WIRE, "replica %d: a value no longer fits the 16-byte wire row: use RLGS_ROWFMT_WIDE", bad);
            if (bad_code != RLGS_ERR_CAPACITY) return fail(bad_code, "replica %d stopped with status %d", bad, bad_code);
            if (s->opts.max_ticks > 0) return fail(RLGS_ERR_CAPACITY, "replica %d reached max_ticks", bad);
            return fail(RLGS_ERR_CAPACITY, "replica %d stopped on a capacity limit (a device table is full)", bad);
        }
        if (all_done) break;
        if (pipelined) {   // the estimate was short: continue one chunk at a time
            target_chunks += 1;
            rc = add_chunks(s, std::max(target_chunks, (int)s->d_chunks.size()), true);
            if (rc) return rc;
        } else if (rows_full) {   // a replica filled the allocated chunks: add some and keep going (state is saved on the device)
            rc = add_chunks(s, (int)s->d_chunks.size() + std::max(1, (int)s->d_chunks.size() / 8), eager_rows);
            if (rc) return rc;
        }
    }
    if (eager_rows && !pipelined) {
        for (size_t k = 0; k < s->d_chunks.size(); ++k) {
            CU(cudaMemcpyAsync(s->h_chunks[k], s->d_chunks[k], chunk_bytes(s), cudaMemcpyDeviceToHost, s->copy_stream));
            s->h_chunk_valid[k] = 1;
        }
    }
    if (eager_jobs) {
        size_t plane = (size_t)R * (size_t)s->Jmax;
        if (s->opts.fetch_jobs == 2) {   // end, finish_order: a finished fifo job started at end - dur_ticks
            CU(cudaMemcpyAsync(s->h_jobs + plane, s->d_jobs + plane, sizeof(int32_t) * 2 * plane, cudaMemcpyDeviceToHost, s->copy_stream));
            s->planes_mask = 6;
        } else {
            CU(cudaMemcpyAsync(s->h_jobs, s->d_jobs, sizeof(int32_t) * 3 * plane, cudaMemcpyDeviceToHost, s->copy_stream));  // start, end, finish_order
            s->planes_mask = 7;
        }
    }
    CU(cudaStreamSynchronize(s->copy_stream));
    s->last_ms = total_ms; s->last_launches = launches;
    s->rows_hint = 0;
    for (int r = 0; r < R; ++r) s->rows_hint = std::max<int64_t>(s->rows_hint, progress_of(s, r).rows);
    for (int r = 0; r < R; ++r) s->h_returns[r] = -(s->pack ? s->h_pstate[r].sum_jct : (s->legacy ? s->h_lstate[r].sum_jct : s->h_state[r].sum_jct));
    s->ran = true;
    for (int r = 0; r < R; ++r) {
        Progress p = progress_of(s, r);
        if (p.status != RLGS_OK) return fail(p.status, "replica %d stopped with status %d after %lld rows", r, p.status, (long long)p.rows);
    }
    return RLGS_OK;
}

extern "C" int32_t rlgs_last_run_ms(rlgs_sim *s, float *kernel_ms, int32_t *n_launches) {
    if (!s || !s->ran) return fail(RLGS_ERR_STATE, "no completed run");
    if (kernel_ms) *kernel_ms = s->last_ms;
    if (n_launches) *n_launches = s->last_launches;
    return RLGS_OK;
}

extern "C" int32_t rlgs_get_summary(rlgs_sim *s, int32_t r, rlgs_summary *out) {
    if (!s || !out) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (!s->ran) return fail(RLGS_ERR_STATE, "no completed run");
    if (r < 0 || r >= s->R) return fail(RLGS_ERR_BAD_ARG, "replica %d out of range", r);
    memset(out, 0, sizeof *out);
    if (s->pack) {
        const PackState &z = s->h_pstate[r];
        out->n_ticks = z.d; out->makespan = z.d; out->sum_jct = z.sum_jct; out->sum_queued = z.sumQ; out->sum_running = z.sumR;
        out->events = z.events; out->n_jobs = s->h_pdesc[r].J; out->n_arrived = z.cursor; out->n_started = z.start_seq;
        out->n_finished = z.F; out->max_queued = z.max_q; out->max_running = z.max_r; out->status = z.status; out->done = z.done;
#ifdef PACK_PROFILE
        static const char *nm[16] = {"arrivals", "q_pop", "score", "heap", "sort", "trials", "real_place", "re_push", "start", "finish", "row", "attempts",
                                     "heap_doomed", "n_doomed_sim", "n_other_sim", "-"};
        for (int k = 0; k < 16; ++k) fprintf(stderr, "prof %-10s %lld\n", nm[k], (long long)z.prof[k]);
#endif
    } else if (!s->legacy) {
        const RepState &z = s->h_state[r];
        out->n_ticks = z.d; out->makespan = z.d; out->sum_jct = z.sum_jct; out->sum_queued = z.sumQ; out->sum_running = z.sumR;
        out->events = z.events; out->n_jobs = s->h_desc[r].J; out->n_arrived = z.cursor; out->n_started = z.start_seq;
        out->n_finished = z.F; out->max_queued = z.max_q; out->max_running = z.max_r; out->status = z.status; out->done = z.done;
    } else {
        const LegState &z = s->h_lstate[r];
        out->n_ticks = z.n_rows; out->makespan = z.t_prev; out->sum_jct = z.sum_jct; out->sum_queued = z.sweep_jobs; out->sum_running = z.demotions;
        out->events = z.events; out->n_jobs = s->h_ldesc[r].J; out->n_arrived = z.cursor; out->n_started = 0;
        out->n_finished = z.F; out->max_queued = z.max_m; out->max_running = 0; out->status = z.status; out->done = z.done;
    }
    return RLGS_OK;
}

// Makes job planes [0, upto) available in h_jobs.  The start plane of a fifo handle that fetched only end + finish_order is
// derived (start = end - dur_ticks) when every started job of every replica has finished; otherwise it is copied like the rest.
static int32_t fetch_planes(rlgs_sim *s, int upto) {
    const int want = (1 << upto) - 1;
    if ((s->planes_mask & want) == want) return RLGS_OK;
    { int32_t rc_a = alloc_host_jobs(s, false); if (rc_a) return rc_a; }
    const size_t plane = (size_t)s->R * (size_t)s->Jmax;
    if (!(s->planes_mask & 1) && (s->planes_mask & 2) && !s->legacy && !s->opts.enable_network_costs) {
        bool derivable = true;
        for (int r = 0; r < s->R && derivable; ++r)
            derivable = s->h_state[r].R == 0 && (int)s->traces[s->rep_trace[r]].host.size() == s->h_desc[r].J;
        if (derivable) {
            for (int r = 0; r < s->R; ++r) {
                const std::vector<rlgs_job> &jobs = s->traces[s->rep_trace[r]].host;
                int32_t *st = s->h_jobs + (size_t)r * s->Jmax;
                const int32_t *en = s->h_jobs + plane + (size_t)r * s->Jmax;
                for (int i = 0; i < (int)jobs.size(); ++i) st[i] = en[i] >= 0 ? en[i] - jobs[i].dur_ticks : -1;
            }
            s->planes_mask |= 1;
        }
    }
    for (int k = 0; k < upto; ++k)
        if (!(s->planes_mask & (1 << k))) {
            CU(cudaMemcpy(s->h_jobs + k * plane, s->d_jobs + k * plane, sizeof(int32_t) * plane, cudaMemcpyDeviceToHost));
            s->planes_mask |= 1 << k;
        }
    return RLGS_OK;
}

static int32_t copy_rows_raw(rlgs_sim *s, int r, int64_t first, int64_t count, unsigned char *out);

// Job planes 0..2 (start, end, finish_order) of replica r in h_jobs: copied from the device, or rebuilt from the replica's event
// rows (RLGS_ROWFMT_EVENT16, no network costs): start_tick = the row that names the job, end_tick = start_tick + dur_ticks when
// the run got that far, finish order = (end tick, start tick), the order release_finished_jobs walks running_jobs (schedule.py:141-162).
// RLGS_ROWFMT_EVENT4: replays the queue of replica r from its 4-byte rows (see rlgs_row4e in include/rlgs.h) and rebuilds job
// planes 0..2 in h_jobs; with `stats` also the pending-time triple of every row (xp_pend).  The queue is kept the way the
// tick loop keeps it: an array filled from the top, the jobs arriving at a tick stored in front of the current front in
// trace order (jobs_manager.py:228-241, q1), the scheduling attempt popping the front (schedule.py:188-190).
static int32_t replay_queue(rlgs_sim *s, int r, bool stats) {
    const std::vector<rlgs_job> &jobs = s->traces[s->rep_trace[r]].host;
    const int J = (int)jobs.size();
    const int64_t n = s->h_state[r].d;
    std::vector<rlgs_row4e> w((size_t)std::max<int64_t>(n, 1));
    if (n > 0) { int32_t rc = copy_rows_raw(s, r, 0, n, reinterpret_cast<unsigned char *>(w.data())); if (rc) return rc; }
    const size_t plane = (size_t)s->R * (size_t)s->Jmax, off = (size_t)r * s->Jmax;
    const bool tables = (s->planes_mask & 7) != 7;   // tables copied from the device are left alone
    int32_t *st = s->h_jobs + off, *en = s->h_jobs + plane + off, *fo = s->h_jobs + 2 * plane + off;
    if (tables) for (int i = 0; i < J; ++i) { st[i] = -1; en[i] = -1; fo[i] = -1; }
    std::vector<int32_t> queue((size_t)std::max(J, 1));
    int front = J, Q = 0, cursor = 0;
    int64_t back_arr = 0;
    std::vector<std::pair<int64_t, int32_t>> fin;   // (end << 32 | start, job)
    if (stats) s->xp_pend.assign((size_t)n * 3, 0);
    for (int64_t i = 0; i < n; ++i) {
        int k = 0;
        while (cursor + k < J && jobs[(size_t)(cursor + k)].arrival_tick <= i) ++k;
        if (k) {
            if (Q == 0) back_arr = i;
            front -= k;
            for (int b = 0; b < k; ++b) queue[(size_t)(front + b)] = cursor + b;
            cursor += k; Q += k;
        }
        const uint32_t x = w[(size_t)i].w;
        if (x & 0x1000u) {
            if (Q == 0) return fail(RLGS_ERR_STATE, "row %lld of replica %d starts a job from an empty queue", (long long)i, r);
            const int j = queue[(size_t)front];
            ++front; --Q;
            const int64_t e = i + jobs[(size_t)j].dur_ticks;
            if (tables) { st[j] = (int32_t)i; if (e <= n) en[j] = (int32_t)e; }
            if (e <= n) fin.push_back(std::make_pair((e << 32) | i, j));
        }
        if ((x >> 13) != ((uint32_t)Q & 0x7ffffu)) return fail(RLGS_ERR_STATE, "row %lld of replica %d: queue length %u on the device, %d in the replay", (long long)i, r, x >> 13, Q);
        if (stats && Q > 0) {
            int32_t *o = &s->xp_pend[(size_t)i * 3];
            o[0] = (int32_t)(i + 1 - back_arr);
            o[1] = (int32_t)(i + 1 - jobs[(size_t)queue[(size_t)(front + (Q - 1) / 2)]].arrival_tick);
            o[2] = (int32_t)(i + 1 - jobs[(size_t)queue[(size_t)(front + Q / 2)]].arrival_tick);
        }
    }
    std::sort(fin.begin(), fin.end());
    if ((int)fin.size() != s->h_state[r].F) return fail(RLGS_ERR_STATE, "replica %d: %zu finished jobs in the event rows, %d on the device", r, fin.size(), s->h_state[r].F);
    if (tables) { for (size_t k = 0; k < fin.size(); ++k) fo[k] = fin[k].second; s->derived[r] = 1; }
    return RLGS_OK;
}

static int32_t ensure_tables(rlgs_sim *s, int r) {
    if ((s->planes_mask & 7) == 7 || s->derived[r]) return RLGS_OK;
    const bool can_derive = (s->wire == RLGS_ROWFMT_EVENT16 || s->wire == RLGS_ROWFMT_EVENT4) && !s->legacy && !s->opts.enable_network_costs && s->opts.rows_mode != RLGS_ROWS_NONE &&
                            (int)s->traces[s->rep_trace[r]].host.size() == s->h_desc[r].J;
    if (!can_derive) return fetch_planes(s, 3);
    { int32_t rc_a = alloc_host_jobs(s, false); if (rc_a) return rc_a; }
    const std::vector<rlgs_job> &jobs = s->traces[s->rep_trace[r]].host;
    const int J = (int)jobs.size();
    const int64_t n = s->h_state[r].d;
    if (s->wire == RLGS_ROWFMT_EVENT4) return replay_queue(s, r, false);
    std::vector<rlgs_row16e> w((size_t)std::max<int64_t>(n, 1));
    if (n > 0) { int32_t rc = copy_rows_raw(s, r, 0, n, reinterpret_cast<unsigned char *>(w.data())); if (rc) return rc; }
    const size_t plane = (size_t)s->R * (size_t)s->Jmax, off = (size_t)r * s->Jmax;
    int32_t *st = s->h_jobs + off, *en = s->h_jobs + plane + off, *fo = s->h_jobs + 2 * plane + off;
    for (int i = 0; i < J; ++i) { st[i] = -1; en[i] = -1; fo[i] = -1; }
    std::vector<std::pair<int64_t, int32_t>> fin;   // (end << 32 | start, job)
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t j1 = w[(size_t)i].w[3];
        if (!j1) continue;
        if ((int64_t)j1 > J) return fail(RLGS_ERR_STATE, "row %lld of replica %d names job %u of %d", (long long)i, r, j1 - 1, J);
        const int j = (int)j1 - 1;
        st[j] = (int32_t)i;
        const int64_t e = i + jobs[(size_t)j].dur_ticks;
        if (e <= n) { en[j] = (int32_t)e; fin.push_back(std::make_pair((e << 32) | i, j)); }
    }
    std::sort(fin.begin(), fin.end());
    if ((int)fin.size() != s->h_state[r].F) return fail(RLGS_ERR_STATE, "replica %d: %zu finished jobs in the event rows, %d on the device", r, fin.size(), s->h_state[r].F);
    for (size_t k = 0; k < fin.size(); ++k) fo[k] = fin[k].second;
    s->derived[r] = 1;
    return RLGS_OK;
}

extern "C" int32_t rlgs_read_job_plane(rlgs_sim *s, int32_t r, int32_t plane_id, int32_t *out) {
    if (!s || !out) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (!s->ran) return fail(RLGS_ERR_STATE, "no completed run");
    if (r < 0 || r >= s->R || plane_id < 0 || plane_id >= N_PLANES) return fail(RLGS_ERR_BAD_ARG, "replica / plane out of range");
    if (!s->legacy && plane_id > RLGS_PLANE_AUX) return fail(RLGS_ERR_BAD_ARG, "plane %d is only recorded by the preemptive schedules", plane_id);
    CU(cudaSetDevice(s->device));
    int32_t rc = plane_id < 3 ? ensure_tables(s, r) : fetch_planes(s, N_PLANES);
    if (rc) return rc;
    size_t plane = (size_t)s->R * (size_t)s->Jmax;
    int J = s->legacy ? s->h_ldesc[r].J : s->h_desc[r].J;
    memcpy(out, s->h_jobs + plane_id * plane + (size_t)r * s->Jmax, 4 * (size_t)J);
    return RLGS_OK;
}

extern "C" int32_t rlgs_read_jobs(rlgs_sim *s, int32_t r, int32_t *finish_order, int32_t *start_tick, int32_t *end_tick,
                                  int32_t *preempt, int32_t *first_node) {
    if (!s) return fail(RLGS_ERR_BAD_ARG, "null handle");
    if (!s->ran) return fail(RLGS_ERR_STATE, "no completed run");
    if (r < 0 || r >= s->R) return fail(RLGS_ERR_BAD_ARG, "replica %d out of range", r);
    CU(cudaSetDevice(s->device));
    int32_t rc = (first_node || (preempt && s->legacy)) ? fetch_planes(s, N_PLANES) : ensure_tables(s, r);   // pack is a legacy-layout family
    if (rc) return rc;
    size_t plane = (size_t)s->R * (size_t)s->Jmax, off = (size_t)r * s->Jmax;
    int J = s->legacy ? s->h_ldesc[r].J : s->h_desc[r].J;
    const int32_t *st = s->h_jobs + off, *en = s->h_jobs + plane + off, *fo = s->h_jobs + 2 * plane + off;
    if (start_tick) memcpy(start_tick, st, 4 * (size_t)J);
    if (end_tick) memcpy(end_tick, en, 4 * (size_t)J);
    if (finish_order) memcpy(finish_order, fo, 4 * (size_t)J);
    if (preempt) {
        if (s->pack) { const int32_t *ns = s->h_jobs + 5 * plane + off; for (int i = 0; i < J; ++i) preempt[i] = ns[i] >= 0 ? ns[i] : (st[i] >= 0 ? 1 : 0); }  // Job.migration_count
        else if (s->legacy) memcpy(preempt, s->h_jobs + 4 * plane + off, 4 * (size_t)J);
        else for (int i = 0; i < J; ++i) preempt[i] = st[i] >= 0 ? 1 : 0;  // Job.migration_count (job.py:171, q6)
    }
    if (first_node) {
        if (s->legacy) return fail(RLGS_ERR_UNSUPPORTED, "placements are recomputed at every event under the preemptive schedules");
        const int32_t *po = s->h_jobs + 3 * plane + off;
        int ll = s->h_state[r].log_len;
        std::vector<int2> log((size_t)std::max(1, ll));
        if (ll > 0) CU(cudaMemcpy(log.data(), s->h_desc[r].place_log, sizeof(int2) * (size_t)ll, cudaMemcpyDeviceToHost));
        for (int i = 0; i < J; ++i) first_node[i] = (po[i] >= 0 && po[i] < ll) ? (log[po[i]].x & 0xffff) : -1;
    }
    return RLGS_OK;
}

// copies `count` rows starting at row `first` of replica r out of the chunk-major store (pinned mirror when valid, else the device)
static int32_t copy_rows_raw(rlgs_sim *s, int r, int64_t first, int64_t count, unsigned char *out) {
    const size_t rb = s->row_bytes;
    int64_t done = 0;
    while (done < count) {
        int64_t i = first + done;
        int k = (int)(i >> RLGS_ROW_CHUNK_LOG);
        int64_t off = i & (RLGS_ROW_CHUNK - 1), len = std::min<int64_t>(count - done, RLGS_ROW_CHUNK - off);
        size_t pos = (((size_t)r << RLGS_ROW_CHUNK_LOG) + (size_t)off) * rb;
        if (s->h_chunk_valid[k]) memcpy(out + done * rb, reinterpret_cast<unsigned char *>(s->h_chunks[k]) + pos, rb * (size_t)len);
        else CU(cudaMemcpy(out + done * rb, reinterpret_cast<unsigned char *>(s->d_chunks[k]) + pos, rb * (size_t)len, cudaMemcpyDeviceToHost));
        done += len;
    }
    return RLGS_OK;
}

static int32_t rows_args_ok(rlgs_sim *s, int32_t r, int64_t first, int64_t count) {
    if (!s->ran) return fail(RLGS_ERR_STATE, "no completed run");
    if (r < 0 || r >= s->R) return fail(RLGS_ERR_BAD_ARG, "replica %d out of range", r);
    if (s->opts.rows_mode == RLGS_ROWS_NONE) return fail(RLGS_ERR_STATE, "rows were not recorded (opts.rows_mode)");
    int64_t n = progress_of(s, r).rows;
    if (first < 0 || count < 0 || first + count > n) return fail(RLGS_ERR_BAD_ARG, "row range [%lld,%lld) out of 0..%lld", (long long)first, (long long)(first + count), (long long)n);
    return RLGS_OK;
}

// Prefix sums behind the expansion of rlgs_row16 (include/rlgs.h): per-job constants summed in start order, in finish order and
// (arrival ticks) in trace order.  Cached for one replica at a time.
static int32_t prepare_expansion(rlgs_sim *s, int r) {
    if (s->xp_replica == r) return RLGS_OK;
    int32_t rc = ensure_tables(s, r);
    if (rc) return rc;
    if (s->wire == RLGS_ROWFMT_EVENT4) { s->xp_replica = -1; rc = replay_queue(s, r, true); if (rc) return rc; }   // the pending-time triples
    const TraceBuf &tb = s->traces[s->rep_trace[r]];
    const int J = tb.n;
    if ((int)tb.host.size() != J) return fail(RLGS_ERR_STATE, "no host copy of the trace of replica %d", r);
    const size_t plane = (size_t)s->R * (size_t)s->Jmax, off = (size_t)r * s->Jmax;
    const int32_t *st = s->h_jobs + off, *fo = s->h_jobs + 2 * plane + off;
    const int F = s->h_state[r].F;
    // start order: one job starts per tick (schedule.py:188-190), so start ticks are distinct; sort the started jobs by them
    std::vector<std::pair<int32_t, int32_t>> so;
    so.reserve((size_t)J);
    for (int i = 0; i < J; ++i) if (st[i] >= 0) so.push_back(std::make_pair(st[i], i));
    std::sort(so.begin(), so.end());
    const size_t S = so.size();
    s->xp_start.resize(S);
    for (size_t k = 0; k < S; ++k) s->xp_start[k] = so[k].first;
    // xp[0..3]: devices, mem_term, util_mu * devices, util_sd^2 * devices in start order; xp[4]: arrival ticks in start order;
    // xp[5..8]: the first four in finish order.  Arrival ticks in trace order are summed on the fly (they are sorted).
    for (int k = 0; k < 5; ++k) s->xp[k].assign(S + 1, 0);
    for (int k = 5; k < 9; ++k) s->xp[k].assign((size_t)F + 1, 0);
    auto consts = [&](int i, int64_t *v) {
        const rlgs_job &j = tb.host[i];
        const int64_t nd = (int64_t)j.tasks * j.gpus_per_task;
        v[0] = nd; v[1] = j.mem_term; v[2] = (int64_t)j.util_mu_q * nd; v[3] = (int64_t)j.util_sd_q * j.util_sd_q * nd;
    };
    int64_t v[4];
    for (size_t k = 0; k < S; ++k) {
        consts(so[k].second, v);
        for (int q = 0; q < 4; ++q) s->xp[q][k + 1] = s->xp[q][k] + v[q];
        s->xp[4][k + 1] = s->xp[4][k] + tb.host[so[k].second].arrival_tick;
    }
    for (int k = 0; k < F; ++k) {
        consts(fo[k], v);
        for (int q = 0; q < 4; ++q) s->xp[5 + q][k + 1] = s->xp[5 + q][k] + v[q];
    }
    s->xp_replica = r;
    return RLGS_OK;
}

static int32_t read_wire_rows(rlgs_sim *s, int32_t r, int64_t first, int64_t count, void *out, int wire) {
    if (!s || !out) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (s->wire != wire) return fail(RLGS_ERR_STATE, "the handle keeps its rows in another format (opts.rows_format = %d)", s->wire);
    int32_t rc = rows_args_ok(s, r, first, count);
    if (rc) return rc;
    CU(cudaSetDevice(s->device));
    return copy_rows_raw(s, r, first, count, reinterpret_cast<unsigned char *>(out));
}
extern "C" int32_t rlgs_read_rows16(rlgs_sim *s, int32_t r, int64_t first, int64_t count, rlgs_row16 *out) { return read_wire_rows(s, r, first, count, out, RLGS_ROWFMT_WIRE16); }
extern "C" int32_t rlgs_read_rows12(rlgs_sim *s, int32_t r, int64_t first, int64_t count, rlgs_row12 *out) { return read_wire_rows(s, r, first, count, out, RLGS_ROWFMT_WIRE12); }
extern "C" int32_t rlgs_read_rows16e(rlgs_sim *s, int32_t r, int64_t first, int64_t count, rlgs_row16e *out) { return read_wire_rows(s, r, first, count, out, RLGS_ROWFMT_EVENT16); }
extern "C" int32_t rlgs_read_rows4e(rlgs_sim *s, int32_t r, int64_t first, int64_t count, rlgs_row4e *out) { return read_wire_rows(s, r, first, count, out, RLGS_ROWFMT_EVENT4); }

extern "C" int32_t rlgs_read_rows(rlgs_sim *s, int32_t r, int64_t first, int64_t count, rlgs_row *out) {
    if (!s || !out) return fail(RLGS_ERR_BAD_ARG, "null argument");
    int32_t rc = rows_args_ok(s, r, first, count);
    if (rc) return rc;
    CU(cudaSetDevice(s->device));
    if (!s->wire) return copy_rows_raw(s, r, first, count, reinterpret_cast<unsigned char *>(out));
    // ---- expansion of the wire rows (see rlgs_row16 / rlgs_row12 in include/rlgs.h)
    rc = prepare_expansion(s, r);
    if (rc) return rc;
    std::vector<unsigned char> w((size_t)count * s->row_bytes);
    rc = copy_rows_raw(s, r, first, count, w.data());
    if (rc) return rc;
    const TraceBuf &tb = s->traces[s->rep_trace[r]];
    const int J = tb.n;
    const size_t plane = (size_t)s->R * (size_t)s->Jmax, off = (size_t)r * s->Jmax;
    const int32_t *en = s->h_jobs + plane + off, *fo = s->h_jobs + 2 * plane + off;
    int64_t arrived = 0, sum_arr_all = 0;   // jobs with arrival_tick <= i and the sum of their arrival ticks
    int64_t n_fin = 0, n_sta = 0;           // rlgs_row12: jobs with end_tick <= i + 1 (finish order is sorted by end tick), with start_tick <= i
    const int64_t S_max = (int64_t)s->xp[0].size() - 1, F_max = (int64_t)s->xp[5].size() - 1;
    auto advance = [&](int64_t i) {
        while (arrived < J && tb.host[(size_t)arrived].arrival_tick <= i) { sum_arr_all += tb.host[(size_t)arrived].arrival_tick; ++arrived; }
        while (n_fin < F_max && en[fo[n_fin]] <= i + 1) ++n_fin;
        while (n_sta < S_max && s->xp_start[(size_t)n_sta] <= i) ++n_sta;
    };
    if (s->wire >= RLGS_ROWFMT_WIRE12 && first > 0) advance(first - 1);   // the cumulative counts are positional: catch up with the rows before `first`
    for (int64_t k = 0; k < count; ++k) {
        const int64_t i = first + k;
        advance(i);
        rlgs_row &o = out[k];
        if (s->wire == RLGS_ROWFMT_WIRE16) {
            const uint32_t *x = reinterpret_cast<const rlgs_row16 *>(w.data())[k].w;
            o.idle_nodes = (int32_t)(x[0] & 0xfffu);
            o.finished = (int32_t)(x[0] >> 12);
            o.queued = (int32_t)(x[1] & 0xfffffu);
            o.max_pending = (int32_t)((x[1] >> 20) | ((x[2] & 0xfffu) << 12));
            o.median_lo = (int32_t)((x[2] >> 12) | ((x[3] & 0xfu) << 20));
            o.median_hi = (int32_t)((x[3] >> 4) & 0xffffffu);
        } else if (s->wire == RLGS_ROWFMT_EVENT4) {
            const uint32_t x = reinterpret_cast<const rlgs_row4e *>(w.data())[k].w;
            const int32_t *pt = &s->xp_pend[(size_t)i * 3];
            o.max_pending = pt[0]; o.median_lo = pt[1]; o.median_hi = pt[2];
            o.idle_nodes = (int32_t)(x & 0xfffu);
            o.finished = (int32_t)n_fin;
            o.queued = (int32_t)(arrived - n_sta);
        } else {
            const uint32_t *x = reinterpret_cast<const uint32_t *>(w.data() + (size_t)k * s->row_bytes);   // rlgs_row12 / rlgs_row16e share w[0..2]
            o.max_pending = (int32_t)(x[0] & 0xffffffu);
            o.median_lo = (int32_t)(x[1] & 0xffffffu);
            o.median_hi = (int32_t)(x[2] & 0xffffffu);
            o.idle_nodes = (int32_t)((x[0] >> 24) | (((x[1] >> 24) & 0xfu) << 8));
            o.finished = (int32_t)n_fin;
            o.queued = (int32_t)(arrived - n_sta);
        }
        const int64_t Fi = o.finished, Q = o.queued, Ri = arrived - Q - Fi, Si = Ri + Fi;
        if (Ri < 0 || Si > S_max || Fi > F_max) return fail(RLGS_ERR_STATE, "row %lld of replica %d is inconsistent with the job tables", (long long)i, r);
        o.running = (int32_t)Ri;
        o.busy_gpus = (int32_t)(s->xp[0][(size_t)Si] - s->xp[5][(size_t)Fi]);
        o.mem_sum = s->xp[1][(size_t)Si] - s->xp[6][(size_t)Fi];
        o.util_mu_sum = s->xp[2][(size_t)Si] - s->xp[7][(size_t)Fi];
        o.util_var_sum = s->xp[3][(size_t)Si] - s->xp[8][(size_t)Fi];
        o.sum_pending = Q * (i + 1) - (sum_arr_all - s->xp[4][(size_t)Si]);
    }
    return RLGS_OK;
}

static int32_t rows_view_any(rlgs_sim *s, int32_t r, int32_t chunk, const void **rows, int64_t *count) {
    if (!s->ran) return fail(RLGS_ERR_STATE, "no completed run");
    if (r < 0 || r >= s->R) return fail(RLGS_ERR_BAD_ARG, "replica %d out of range", r);
    if (s->opts.rows_mode == RLGS_ROWS_NONE) return fail(RLGS_ERR_STATE, "rows were not recorded (opts.rows_mode)");
    int64_t n = progress_of(s, r).rows;
    if (chunk < 0 || (int64_t)chunk * RLGS_ROW_CHUNK >= std::max<int64_t>(n, 1)) return fail(RLGS_ERR_BAD_ARG, "chunk %d out of range", chunk);
    CU(cudaSetDevice(s->device));
    if (!s->h_chunk_valid[chunk]) {
        if (!s->h_chunks[chunk]) CU(cudaMallocHost(&s->h_chunks[chunk], chunk_bytes(s)));
        CU(cudaMemcpy(s->h_chunks[chunk], s->d_chunks[chunk], chunk_bytes(s), cudaMemcpyDeviceToHost));
        s->h_chunk_valid[chunk] = 1;
    }
    *rows = reinterpret_cast<unsigned char *>(s->h_chunks[chunk]) + ((size_t)r << RLGS_ROW_CHUNK_LOG) * s->row_bytes;
    *count = std::min<int64_t>(RLGS_ROW_CHUNK, n - (int64_t)chunk * RLGS_ROW_CHUNK);
    return RLGS_OK;
}

extern "C" int32_t rlgs_rows_view(rlgs_sim *s, int32_t r, int32_t chunk, const rlgs_row **rows, int64_t *count) {
    if (!s || !rows || !count) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (s->wire) return fail(RLGS_ERR_STATE, "the handle keeps wire rows: use rlgs_rows16_view / rlgs_rows12_view, or rlgs_read_rows for expanded rows");
    return rows_view_any(s, r, chunk, reinterpret_cast<const void **>(rows), count);
}

extern "C" int32_t rlgs_rows16_view(rlgs_sim *s, int32_t r, int32_t chunk, const rlgs_row16 **rows, int64_t *count) {
    if (!s || !rows || !count) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (s->wire != RLGS_ROWFMT_WIRE16) return fail(RLGS_ERR_STATE, "the handle keeps its rows in another format (opts.rows_format)");
    return rows_view_any(s, r, chunk, reinterpret_cast<const void **>(rows), count);
}

extern "C" int32_t rlgs_rows16e_view(rlgs_sim *s, int32_t r, int32_t chunk, const rlgs_row16e **rows, int64_t *count) {
    if (!s || !rows || !count) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (s->wire != RLGS_ROWFMT_EVENT16) return fail(RLGS_ERR_STATE, "the handle keeps its rows in another format (opts.rows_format)");
    return rows_view_any(s, r, chunk, reinterpret_cast<const void **>(rows), count);
}

extern "C" int32_t rlgs_rows4e_view(rlgs_sim *s, int32_t r, int32_t chunk, const rlgs_row4e **rows, int64_t *count) {
    if (!s || !rows || !count) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (s->wire != RLGS_ROWFMT_EVENT4) return fail(RLGS_ERR_STATE, "the handle keeps its rows in another format (opts.rows_format)");
    return rows_view_any(s, r, chunk, reinterpret_cast<const void **>(rows), count);
}

extern "C" int32_t rlgs_rows12_view(rlgs_sim *s, int32_t r, int32_t chunk, const rlgs_row12 **rows, int64_t *count) {
    if (!s || !rows || !count) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (s->wire != RLGS_ROWFMT_WIRE12) return fail(RLGS_ERR_STATE, "the handle keeps its rows in another format (opts.rows_format)");
    return rows_view_any(s, r, chunk, reinterpret_cast<const void **>(rows), count);
}

// Host-only expansion of one replica's rlgs_row4e stream (include/rlgs.h).  Same rule as replay_queue above, which serves the
// handle's own reads and is pinned by the GPU tests; this entry point is pinned on the CPU (tests/test_abi.py) against the oracle
// and the reference's golden files.
extern "C" int32_t rlgs_replay_rows4e(const rlgs_job *jobs, int32_t J, const rlgs_row4e *rows, int64_t n, int32_t *st, int32_t *en,
                                      int32_t *fo, int32_t *n_finished, int32_t *pending) {
    if (J < 0 || n < 0 || (J > 0 && !jobs) || (n > 0 && !rows)) return fail(RLGS_ERR_BAD_ARG, "null / negative argument");
    for (int i = 0; i < J; ++i) { if (st) st[i] = -1; if (en) en[i] = -1; if (fo) fo[i] = -1; }
    std::vector<int32_t> queue((size_t)std::max(J, 1));
    int front = J, Q = 0, cursor = 0;
    int64_t back_arr = 0;
    std::vector<std::pair<int64_t, int32_t>> fin;   // (end << 32 | start, job)
    for (int64_t i = 0; i < n; ++i) {
        int k = 0;
        while (cursor + k < J && jobs[cursor + k].arrival_tick <= i) ++k;   // the tick's arrivals go to the front, in trace order (q1)
        if (k) {
            if (Q == 0) back_arr = i;
            front -= k;
            for (int b = 0; b < k; ++b) queue[(size_t)(front + b)] = cursor + b;
            cursor += k; Q += k;
        }
        const uint32_t x = rows[i].w;
        if (x & 0x1000u) {                                                  // the attempt started the front (schedule.py:188-190)
            if (Q == 0) return fail(RLGS_ERR_STATE, "row %lld starts a job from an empty queue", (long long)i);
            const int j = queue[(size_t)front];
            ++front; --Q;
            const int64_t e = i + jobs[j].dur_ticks;
            if (st) st[j] = (int32_t)i;
            if (e <= n) { if (en) en[j] = (int32_t)e; fin.push_back(std::make_pair((e << 32) | i, j)); }
        }
        if ((x >> 13) != ((uint32_t)Q & 0x7ffffu)) return fail(RLGS_ERR_STATE, "row %lld: queue length %u in the row, %d in the replay", (long long)i, x >> 13, Q);
        if (pending) {
            int32_t *o = pending + 3 * i;
            o[0] = o[1] = o[2] = 0;
            if (Q > 0) {
                o[0] = (int32_t)(i + 1 - back_arr);
                o[1] = (int32_t)(i + 1 - jobs[queue[(size_t)(front + (Q - 1) / 2)]].arrival_tick);
                o[2] = (int32_t)(i + 1 - jobs[queue[(size_t)(front + Q / 2)]].arrival_tick);
            }
        }
    }
    std::sort(fin.begin(), fin.end());                                      // release order: end tick, then start tick (schedule.py:141-162)
    if (fo) for (size_t k = 0; k < fin.size(); ++k) fo[k] = fin[k].second;
    if (n_finished) *n_finished = (int32_t)fin.size();
    return RLGS_OK;
}

extern "C" int32_t rlgs_returns(rlgs_sim *s, int64_t *out) {
    if (!s || !out) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (!s->ran) return fail(RLGS_ERR_STATE, "no completed run");
    memcpy(out, s->h_returns, sizeof(int64_t) * (size_t)s->R);
    return RLGS_OK;
}

extern "C" int32_t rlgs_returns_device_ptr(rlgs_sim *s, void **dev_ptr) {
    if (!s || !dev_ptr) return fail(RLGS_ERR_BAD_ARG, "null argument");
    *dev_ptr = s->d_returns;
    return RLGS_OK;
}

// ------------------------------------------------------------------------------------------------
// Vectorised RL environment (replaces the stub model/env.py:1-6 of the reference; semantics are
// build-defined, see fifo_grp.cuh EnvIO).  All pointers are DEVICE pointers (torch tensors);
// calls are asynchronous on the handle's stream (rlgs_set_stream) until rlgs_env_sync.
// ------------------------------------------------------------------------------------------------
extern "C" int32_t rlgs_env_obs_dim(rlgs_sim *s, int32_t window_k, int32_t *dim) {
    if (!s || !dim) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (window_k < 1 || window_k > 32) return fail(RLGS_ERR_BAD_ARG, "window_k must be 1..32");
    *dim = 3 * s->cc.N + 5 * window_k + 4;
    return RLGS_OK;
}

extern "C" int32_t rlgs_env_reset(rlgs_sim *s) {
    if (!s) return fail(RLGS_ERR_BAD_ARG, "null handle");
    if (s->legacy) return fail(RLGS_ERR_UNSUPPORTED, "the environment steps the fifo tick loop");
    CU(cudaSetDevice(s->device));
    int32_t rc = setup_job_arrays(s);
    if (rc) return rc;
    size_t smem = (32 / s->lpr) * grp_smem_bytes(s->cc.N, s->cc.G, s->slot_cap, s->lpr);
    if (smem > 227 * 1024) return fail(RLGS_ERR_CAPACITY, "cluster state needs %zu B of shared memory per warp", smem);
    if (s->opts.rows_mode != RLGS_ROWS_NONE) {   // per-tick rows while stepping: 64-byte rows, one replica per warp
        if (s->lpr != 32 || s->wire) return fail(RLGS_ERR_UNSUPPORTED, "environment steps with rows need lanes_per_replica = 32 and RLGS_ROWFMT_WIDE");
        if (s->opts.rows_mode != RLGS_ROWS_DEVICE) return fail(RLGS_ERR_UNSUPPORTED, "environment steps keep their rows on the device (RLGS_ROWS_DEVICE)");
        rc = add_chunks(s, std::max<int>(1, (int)s->d_chunks.size()), false);
        if (rc) return rc;
        std::fill(s->h_chunk_valid.begin(), s->h_chunk_valid.end(), 0);
    }
    s->env_ticks = 0;
    cudaStream_t st = s->use_user_stream ? (cudaStream_t)s->user_stream : s->stream;
    for (int r = 0; r < s->R; ++r) {
        RepState z; memset(&z, 0, sizeof z);
        z.head = s->h_desc[r].J; z.idle_nodes = s->cc.N;
        z.n_free_nodes = (s->cc.cpu_cap > 0 || s->cc.mem_cap > 0) ? s->cc.N : 0;
        z.free_hint = -1;
        s->h_init[r] = z;
    }
    CU(cudaMemcpyAsync(s->d_state, s->h_init, sizeof(RepState) * s->R, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->d_desc, s->h_desc.data(), sizeof(RepDesc) * s->R, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(s->d_jobs, 0xff, s->jobs_bytes, st));
    CU(cudaStreamSynchronize(st));   // h_init / h_desc may be rewritten by the caller's next call
    s->planes_mask = 0; std::fill(s->derived.begin(), s->derived.end(), 0); s->ran = false; s->env_ready = true;
    return RLGS_OK;
}

extern "C" int32_t rlgs_env_step(rlgs_sim *s, const int32_t *actions, float *obs, float *reward, uint8_t *done,
                                 int32_t policy, int32_t window_k, uint32_t seed, int32_t n_ticks) {
    if (!s || !obs || !reward || !done) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (s->legacy) return fail(RLGS_ERR_UNSUPPORTED, "the environment steps the fifo tick loop");
    if (policy < 0 || policy > 2 || (policy == 2 && !actions)) return fail(RLGS_ERR_BAD_ARG, "policy must be 0, 1 or 2 (2 needs actions)");
    if (window_k < 1 || window_k > 32) return fail(RLGS_ERR_BAD_ARG, "window_k must be 1..32");
    if (n_ticks < 1 || (policy == 2 && n_ticks != 1)) return fail(RLGS_ERR_BAD_ARG, "n_ticks must be >= 1 (exactly 1 with external actions)");
    if (!s->env_ready) return fail(RLGS_ERR_STATE, "rlgs_env_step needs rlgs_env_reset after rlgs_load_trace");
    CU(cudaSetDevice(s->device));
    cudaStream_t st = s->use_user_stream ? (cudaStream_t)s->user_stream : s->stream;
    EnvIO io; io.actions = actions; io.obs = obs; io.reward = reward; io.done = done; io.policy = policy; io.window_k = window_k;
    io.seed = seed; io.obs_dim = 3 * s->cc.N + 5 * window_k + 4;
    RowStore rs; rs.chunks = nullptr; rs.n_chunks = 0; rs.replica = 0;
    const bool rows = s->opts.rows_mode != RLGS_ROWS_NONE;
    if (rows) {
        // rows of the ticks this call may simulate: grow the store up to 16 chunks ahead; a launch stops where the store ends
        int64_t upto = std::min<int64_t>(s->env_ticks + n_ticks, s->env_ticks + 16 * (int64_t)RLGS_ROW_CHUNK);
        int need = (int)((upto + RLGS_ROW_CHUNK - 1) / RLGS_ROW_CHUNK);
        if (need > (int)s->d_chunks.size()) { CU(cudaStreamSynchronize(st)); int32_t rc = add_chunks(s, need, false); if (rc) return rc; }
        s->env_ticks = std::min<int64_t>(upto, (int64_t)s->d_chunks.size() * RLGS_ROW_CHUNK);
        rs.chunks = s->d_chunk_ptrs; rs.n_chunks = (int)s->d_chunks.size();
    }
    CU(launch_fifo(s, 0, s->R, n_ticks, rows ? 1 : 0, true, io, rs, st));
    CU(cudaGetLastError());
    return RLGS_OK;
}

// Observation of the current state without advancing it (a zero-tick launch of the same kernel): what reset() returns.
extern "C" int32_t rlgs_env_observe(rlgs_sim *s, float *obs, float *reward, uint8_t *done, int32_t window_k) {
    if (!s || !obs || !reward || !done) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (s->legacy) return fail(RLGS_ERR_UNSUPPORTED, "the environment steps the fifo tick loop");
    if (window_k < 1 || window_k > 32) return fail(RLGS_ERR_BAD_ARG, "window_k must be 1..32");
    if (!s->env_ready) return fail(RLGS_ERR_STATE, "rlgs_env_observe needs rlgs_env_reset after rlgs_load_trace");
    CU(cudaSetDevice(s->device));
    cudaStream_t st = s->use_user_stream ? (cudaStream_t)s->user_stream : s->stream;
    EnvIO io; io.actions = nullptr; io.obs = obs; io.reward = reward; io.done = done; io.policy = 0; io.window_k = window_k;
    io.seed = 0; io.obs_dim = 3 * s->cc.N + 5 * window_k + 4;
    RowStore rs; rs.chunks = nullptr; rs.n_chunks = 0; rs.replica = 0;
    CU(launch_fifo(s, 0, s->R, 0, 0, true, io, rs, st));
    return RLGS_OK;
}

// Waits for the enqueued steps and refreshes the host copy of the replica states, so that
// rlgs_get_summary / rlgs_read_jobs / rlgs_returns describe the environment's current state.
extern "C" int32_t rlgs_env_sync(rlgs_sim *s) {
    if (!s) return fail(RLGS_ERR_BAD_ARG, "null handle");
    if (s->legacy) return fail(RLGS_ERR_UNSUPPORTED, "the environment steps the fifo tick loop");
    if (!s->env_ready) return fail(RLGS_ERR_STATE, "rlgs_env_sync needs rlgs_env_reset after rlgs_load_trace");
    CU(cudaSetDevice(s->device));
    cudaStream_t st = s->use_user_stream ? (cudaStream_t)s->user_stream : s->stream;
    CU(cudaMemcpyAsync(s->h_state, s->d_state, sizeof(RepState) * s->R, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int r = 0; r < s->R; ++r) {
        s->h_returns[r] = -s->h_state[r].sum_jct;
        if (s->h_state[r].status != RLGS_OK) return fail(s->h_state[r].status, "replica %d stopped with status %d at tick %d", r, s->h_state[r].status, s->h_state[r].d);
    }
    s->planes_mask = 0; std::fill(s->derived.begin(), s->derived.end(), 0); s->ran = true;
    return RLGS_OK;
}

// Duration of every job after network costs (Job.add_network_costs, job.py:196-197): the value the
// reference prints in both duration columns of job.csv.  Jobs that never started keep their input duration.
extern "C" int32_t rlgs_read_durations(rlgs_sim *s, int32_t r, double *out) {
    if (!s || !out) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (!s->ran) return fail(RLGS_ERR_STATE, "no completed run");
    if (r < 0 || r >= s->R) return fail(RLGS_ERR_BAD_ARG, "replica %d out of range", r);
    if (s->legacy || !s->opts.enable_network_costs) return fail(RLGS_ERR_STATE, "network costs are not enabled");
    CU(cudaSetDevice(s->device));
    int J = s->h_desc[r].J;
    std::vector<int32_t> st((size_t)J);
    int32_t rc = rlgs_read_jobs(s, r, nullptr, st.data(), nullptr, nullptr, nullptr);
    if (rc) return rc;
    std::vector<double> in((size_t)J);
    CU(cudaMemcpy(in.data(), s->h_desc[r].net_in, 8 * (size_t)J, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out, s->h_desc[r].dur_out, 8 * (size_t)J, cudaMemcpyDeviceToHost));
    for (int i = 0; i < J; ++i) if (st[i] < 0) out[i] = in[i];
    return RLGS_OK;
}
