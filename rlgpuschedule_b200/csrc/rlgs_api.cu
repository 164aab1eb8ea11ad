// rlgs_api.cu — host side of the C ABI declared in include/rlgs.h (librlgs.so).
//
// Owns device memory, launches the simulation kernels chunk by chunk, streams the per-tick rows
// to a pinned host store while the next chunk computes, and hands results back through plain
// pointers.  No torch types, no CPU fallback: every entry point fails with RLGS_ERR_CUDA when no
// device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/rlgs.h"
#include "fifo_yarn.cuh"

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

struct TraceBuf {
    rlgs_job *dev = nullptr;
    int32_t n = 0, cap_n = 0;
    int64_t log_cap = 0, cap_log = 0;
    int32_t max_arrival = 0;
    int first = 0, count = 0;   // replica range this trace is attached to
};

struct Group {            // a contiguous range of replicas driven through one CUDA stream
    int first = 0, count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t k_begin = nullptr, k_end = nullptr;
};

struct rlgs_sim {
    rlgs_cluster_spec spec;
    rlgs_opts opts;
    ClusterConst cc;
    int R = 0;
    int device = 0;
    cudaStream_t stream = nullptr;  // main stream: state upload, fork / join point
    void *user_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t copy_stream = nullptr;   // host-bound copies of the pipelined (rows FULL) path
    std::vector<cudaEvent_t> ev_pool;
    std::vector<Group> groups;
    std::vector<TraceBuf> traces;
    std::vector<int> rep_trace;  // trace id per replica, -1 = none
    std::vector<RepDesc> h_desc;
    RepDesc *d_desc = nullptr;
    RepState *d_state = nullptr;
    RepState *h_state = nullptr; // pinned [R]
    RepState *h_init = nullptr;  // pinned [R] initial states
    std::vector<void *> slabs;   // per load_trace call
    int slot_cap = 0;
    // rows: device-resident store [R][rows_cap] + pinned host mirror [R][h_cap]
    rlgs_row *d_rows = nullptr;
    int64_t rows_cap = 0;
    rlgs_row *h_rows = nullptr;
    int64_t h_cap = 0;
    bool rows_fetched = false;
    // job tables: device [4][R][Jmax] (start, end, finish_order, place_off) + pinned host mirror
    int32_t *d_jobs = nullptr;
    int32_t *h_jobs = nullptr;
    size_t jobs_bytes = 0;
    bool jobs_fetched = false;
    bool jobs_partial = false;   // start/end/finish_order planes are on the host, place_off is not
    int32_t Jmax = 0;
    int64_t *d_returns = nullptr;
    int64_t *h_returns = nullptr;
    bool ran = false;
    float last_ms = 0.f;
    int last_launches = 0;
};

extern "C" int32_t rlgs_version(void) { return RLGS_VERSION; }
extern "C" const char *rlgs_last_error(void) { return g_err; }

extern "C" int32_t rlgs_create(const rlgs_cluster_spec *spec, const rlgs_opts *opts, rlgs_sim **out) {
    if (!spec || !opts || !out) return fail(RLGS_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    int64_t N = (int64_t)spec->num_switch * spec->num_node_p_switch;
    if (spec->num_switch < 1 || spec->num_node_p_switch < 1 || N > 4096)
        return fail(RLGS_ERR_BAD_ARG, "cluster must have 1..4096 nodes (got %lld)", (long long)N);
    if (spec->num_gpu_p_node < 1 || spec->num_gpu_p_node > 32)
        return fail(RLGS_ERR_BAD_ARG, "num_gpu_p_node must be 1..32 (got %d)", spec->num_gpu_p_node);
    if (opts->n_replicas < 1) return fail(RLGS_ERR_BAD_ARG, "n_replicas must be >= 1");
    if (opts->schedule != RLGS_SCHED_FIFO)
        return fail(RLGS_ERR_UNSUPPORTED, "schedule id %d is not implemented on the device path", opts->schedule);
    if (opts->placement != RLGS_PLACE_YARN)
        return fail(RLGS_ERR_UNSUPPORTED, "placement id %d is not implemented for this schedule", opts->placement);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(RLGS_ERR_CUDA, "no CUDA device: %s (librlgs has no CPU fallback)", cudaGetErrorString(e));
    if (opts->device < 0 || opts->device >= ndev) return fail(RLGS_ERR_BAD_ARG, "device %d out of range", opts->device);
    CU(cudaSetDevice(opts->device));
    rlgs_sim *s = new (std::nothrow) rlgs_sim();
    if (!s) return fail(RLGS_ERR_OOM, "host allocation failed");
    s->spec = *spec; s->opts = *opts; s->R = opts->n_replicas; s->device = opts->device;
    s->cc.N = (int)N; s->cc.G = spec->num_gpu_p_node; s->cc.cpu_cap = spec->num_cpu_p_node; s->cc.mem_cap = spec->mem_p_node;
    s->cc.gmask = spec->num_gpu_p_node == 32 ? 0xffffffffu : ((1u << spec->num_gpu_p_node) - 1u);
    s->cc.D = s->cc.N * s->cc.G;
    s->slot_cap = opts->slot_cap > 0 ? opts->slot_cap : std::min(128, std::max(32, s->cc.D));
    s->slot_cap = (s->slot_cap + 31) & ~31;
    s->rep_trace.assign(s->R, -1);
    s->h_desc.assign(s->R, RepDesc{});
    int ng = opts->n_streams > 0 ? opts->n_streams : (s->R >= 8 * 148 ? 4 : (s->R >= 2 * 148 ? 2 : 1));
    ng = std::max(1, std::min(ng, s->R));
    cudaError_t ce = cudaSuccess;
    auto ok = [&](cudaError_t e) { if (ce == cudaSuccess) ce = e; return e == cudaSuccess; };
    ok(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    ok(cudaEventCreate(&s->ev_fork)); ok(cudaEventCreate(&s->ev_join));
    ok(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    for (int g = 0; g < ng; ++g) {
        Group G;
        G.first = (int)((int64_t)s->R * g / ng); G.count = (int)((int64_t)s->R * (g + 1) / ng) - G.first;
        ok(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
        ok(cudaEventCreate(&G.k_begin)); ok(cudaEventCreate(&G.k_end));
        s->groups.push_back(G);
    }
    ok(cudaMalloc(&s->d_desc, sizeof(RepDesc) * s->R));
    ok(cudaMalloc(&s->d_state, sizeof(RepState) * s->R));
    ok(cudaMalloc(&s->d_returns, sizeof(int64_t) * s->R));
    ok(cudaMallocHost(&s->h_returns, sizeof(int64_t) * s->R));
    ok(cudaMallocHost(&s->h_state, sizeof(RepState) * s->R));
    ok(cudaMallocHost(&s->h_init, sizeof(RepState) * s->R));
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
    for (auto &t : s->traces) cudaFree(t.dev);
    for (void *p : s->slabs) cudaFree(p);
    cudaFree(s->d_desc); cudaFree(s->d_state); cudaFree(s->d_rows); cudaFree(s->d_jobs); cudaFree(s->d_returns);
    if (s->h_rows) cudaFreeHost(s->h_rows);
    if (s->h_jobs) cudaFreeHost(s->h_jobs);
    if (s->h_returns) cudaFreeHost(s->h_returns);
    if (s->h_state) cudaFreeHost(s->h_state);
    if (s->h_init) cudaFreeHost(s->h_init);
    for (auto &G : s->groups) {
        if (G.k_begin) cudaEventDestroy(G.k_begin);
        if (G.k_end) cudaEventDestroy(G.k_end);
        if (G.stream) cudaStreamDestroy(G.stream);
    }
    for (auto e : s->ev_pool) cudaEventDestroy(e);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int32_t rlgs_set_stream(rlgs_sim *s, void *cuda_stream) {
    if (!s) return fail(RLGS_ERR_BAD_ARG, "null handle");
    s->user_stream = cuda_stream;
    return RLGS_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int32_t rlgs_load_trace(rlgs_sim *s, int32_t first, int32_t count, const rlgs_job *jobs, int32_t n,
                                   const rlgs_netcost_inputs *net) {
    if (!s || !jobs) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (n < 1) return fail(RLGS_ERR_BAD_ARG, "trace has no jobs (the reference asserts on an empty job table, log_manager.py:138)");
    if (first < 0 || count < 1 || first + count > s->R) return fail(RLGS_ERR_BAD_ARG, "replica range [%d,%d) out of 0..%d", first, first + count, s->R);
    if (net && s->opts.enable_network_costs) return fail(RLGS_ERR_UNSUPPORTED, "network costs are not implemented for this schedule yet");
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
    for (size_t t = 0; t < s->traces.size(); ++t) {
        TraceBuf &old = s->traces[t];
        if (old.first == first && old.count == count && n <= old.cap_n && tb.log_cap <= old.cap_log) {
            CU(cudaMemcpy(old.dev, jobs, sizeof(rlgs_job) * (size_t)n, cudaMemcpyHostToDevice));
            old.n = n; old.log_cap = tb.log_cap; old.max_arrival = tb.max_arrival;
            for (int r = 0; r < count; ++r) { s->h_desc[first + r].J = n; s->h_desc[first + r].log_cap = (int32_t)std::min<int64_t>(tb.log_cap, 0x7fffffff); }
            s->ran = false;
            return RLGS_OK;
        }
    }
    tb.cap_n = n; tb.cap_log = tb.log_cap;
    CU(cudaMalloc(&tb.dev, sizeof(rlgs_job) * (size_t)n));
    cudaError_t e = cudaMemcpy(tb.dev, jobs, sizeof(rlgs_job) * (size_t)n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(tb.dev); return fail(RLGS_ERR_CUDA, "trace upload: %s", cudaGetErrorString(e)); }
    int tid = (int)s->traces.size();
    s->traces.push_back(tb);
    // per-replica working set: stack | place_log | node_save | slot_save
    int nw = 3 * s->cc.N + (s->cc.N + 31) / 32;
    size_t per = align_up(sizeof(rlgs_job) * (size_t)n, 256) + align_up(sizeof(int2) * (size_t)std::max<int64_t>(tb.log_cap, 1), 256) +
                 align_up(4 * (size_t)nw, 256) + align_up(sizeof(int4) * 2 * (size_t)s->slot_cap, 256);
    unsigned char *slab = nullptr;
    CU(cudaMalloc(&slab, per * (size_t)count));
    s->slabs.push_back(slab);
    for (int r = 0; r < count; ++r) {
        unsigned char *p = slab + per * (size_t)r;
        RepDesc &D = s->h_desc[first + r];
        D.trace = tb.dev; D.J = n; D.log_cap = (int32_t)std::min<int64_t>(tb.log_cap, 0x7fffffff);
        D.stack = reinterpret_cast<rlgs_job *>(p); p += align_up(sizeof(rlgs_job) * (size_t)n, 256);
        D.place_log = reinterpret_cast<int2 *>(p); p += align_up(sizeof(int2) * (size_t)std::max<int64_t>(tb.log_cap, 1), 256);
        D.node_save = reinterpret_cast<int32_t *>(p); p += align_up(4 * (size_t)nw, 256);
        D.slot_save = reinterpret_cast<int4 *>(p);
        s->rep_trace[first + r] = tid;
    }
    s->ran = false;
    return RLGS_OK;
}

// (re)allocates the [4][R][Jmax] job-output arrays and points every replica at its rows
static int32_t setup_job_arrays(rlgs_sim *s) {
    int32_t Jmax = 0;
    for (int r = 0; r < s->R; ++r) {
        if (s->rep_trace[r] < 0) return fail(RLGS_ERR_STATE, "replica %d has no trace (call rlgs_load_trace)", r);
        Jmax = std::max(Jmax, s->h_desc[r].J);
    }
    if (Jmax != s->Jmax || !s->d_jobs) {
        cudaFree(s->d_jobs); s->d_jobs = nullptr;
        if (s->h_jobs) { cudaFreeHost(s->h_jobs); s->h_jobs = nullptr; }
        s->Jmax = Jmax;
        s->jobs_bytes = sizeof(int32_t) * 4 * (size_t)s->R * (size_t)Jmax;
        CU(cudaMalloc(&s->d_jobs, s->jobs_bytes));
    }
    size_t plane = (size_t)s->R * (size_t)Jmax;
    for (int r = 0; r < s->R; ++r) {
        RepDesc &D = s->h_desc[r];
        D.start_tick = s->d_jobs + 0 * plane + (size_t)r * Jmax;
        D.end_tick = s->d_jobs + 1 * plane + (size_t)r * Jmax;
        D.finish_order = s->d_jobs + 2 * plane + (size_t)r * Jmax;
        D.place_off = s->d_jobs + 3 * plane + (size_t)r * Jmax;
    }
    return RLGS_OK;
}

// grows the device row store to `cap` ticks per replica, keeping the first `keep` rows of each replica
static int32_t grow_device_rows(rlgs_sim *s, int64_t cap, int64_t keep) {
    rlgs_row *nw = nullptr;
    CU(cudaMalloc(&nw, sizeof(rlgs_row) * (size_t)cap * (size_t)s->R));
    if (s->d_rows && keep > 0)
        CU(cudaMemcpy2D(nw, sizeof(rlgs_row) * (size_t)cap, s->d_rows, sizeof(rlgs_row) * (size_t)s->rows_cap,
                        sizeof(rlgs_row) * (size_t)keep, (size_t)s->R, cudaMemcpyDeviceToDevice));
    cudaFree(s->d_rows);
    s->d_rows = nw; s->rows_cap = cap;
    return RLGS_OK;
}

static int32_t ensure_host_rows(rlgs_sim *s, int64_t cap) {
    if (cap <= s->h_cap) return RLGS_OK;
    if (s->h_rows) cudaFreeHost(s->h_rows);
    s->h_rows = nullptr; s->h_cap = 0;
    CU(cudaMallocHost(&s->h_rows, sizeof(rlgs_row) * (size_t)cap * (size_t)s->R));
    s->h_cap = cap;
    return RLGS_OK;
}

// enqueue the device->host copies of one group's results on its stream
static int32_t enqueue_fetch(rlgs_sim *s, const Group &G, bool rows, bool jobs, int64_t width) {
    if (rows && width > 0)
        CU(cudaMemcpy2DAsync(s->h_rows + (size_t)G.first * s->h_cap, sizeof(rlgs_row) * (size_t)s->h_cap,
                             s->d_rows + (size_t)G.first * s->rows_cap, sizeof(rlgs_row) * (size_t)s->rows_cap,
                             sizeof(rlgs_row) * (size_t)width, (size_t)G.count, cudaMemcpyDeviceToHost, G.stream));
    if (jobs) {
        size_t plane = (size_t)s->R * (size_t)s->Jmax;
        for (int k = 0; k < 4; ++k)
            CU(cudaMemcpyAsync(s->h_jobs + k * plane + (size_t)G.first * s->Jmax, s->d_jobs + k * plane + (size_t)G.first * s->Jmax,
                               sizeof(int32_t) * (size_t)G.count * (size_t)s->Jmax, cudaMemcpyDeviceToHost, G.stream));
    }
    return RLGS_OK;
}

extern "C" int32_t rlgs_run(rlgs_sim *s) {
    if (!s) return fail(RLGS_ERR_BAD_ARG, "null handle");
    CU(cudaSetDevice(s->device));
    int32_t rc = setup_job_arrays(s);
    if (rc) return rc;
    cudaStream_t main_st = s->user_stream ? (cudaStream_t)s->user_stream : s->stream;
    const int mode = s->opts.rows_mode;
    const bool rows = mode != RLGS_ROWS_NONE, eager_rows = mode == RLGS_ROWS_FULL, eager_jobs = s->opts.fetch_jobs != 0;
    const int R = s->R;
    size_t smem = fifo_smem_bytes(s->cc.N, s->slot_cap);
    if (smem > 227 * 1024) return fail(RLGS_ERR_CAPACITY, "cluster state needs %zu B of shared memory per replica (> 227 KB)", smem);
    CU(cudaFuncSetAttribute(fifo_yarn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    int32_t max_arrival = 0;
    for (int r = 0; r < R; ++r) {
        RepState z; memset(&z, 0, sizeof z);
        z.head = s->h_desc[r].J; z.idle_nodes = s->cc.N;
        z.n_free_nodes = (s->cc.cpu_cap > 0 || s->cc.mem_cap > 0) ? s->cc.N : 0;
        z.free_hint = -1;
        s->h_init[r] = z;
        max_arrival = std::max(max_arrival, s->traces[s->rep_trace[r]].max_arrival);
    }
    if (rows && !s->d_rows) {
        int64_t cap = s->opts.rows_cap > 0 ? s->opts.rows_cap : (int64_t)max_arrival + 4096;
        rc = grow_device_rows(s, cap, 0);
        if (rc) return rc;
    }
    if (eager_rows) { rc = ensure_host_rows(s, s->rows_cap); if (rc) return rc; }
    if (eager_jobs && !s->h_jobs) CU(cudaMallocHost(&s->h_jobs, s->jobs_bytes));
    s->jobs_fetched = false; s->jobs_partial = false; s->rows_fetched = false; s->ran = false;

    CU(cudaMemcpyAsync(s->d_state, s->h_init, sizeof(RepState) * R, cudaMemcpyHostToDevice, main_st));
    CU(cudaMemcpyAsync(s->d_desc, s->h_desc.data(), sizeof(RepDesc) * R, cudaMemcpyHostToDevice, main_st));
    CU(cudaMemsetAsync(s->d_jobs, 0xff, s->jobs_bytes, main_st));

    if (eager_rows && s->opts.ticks_per_launch == 0) {
        // ---- pipelined path: launches bounded to PIPE_TICKS ticks; the rows of chunk k travel to the
        // pinned host store on the copy stream while chunk k+1 is being simulated.  Nothing here waits
        // on the host until every chunk is enqueued.
        const int PIPE_TICKS = 8192;
        size_t plane = (size_t)s->R * (size_t)s->Jmax;
        int64_t next_tick = 0;
        int launches = 0;
        float total_ms = 0.f;
        auto get_event = [&](size_t i, cudaEvent_t *out) -> cudaError_t {
            while (s->ev_pool.size() <= i) { cudaEvent_t e; cudaError_t ce = cudaEventCreate(&e); if (ce != cudaSuccess) return ce; s->ev_pool.push_back(e); }
            *out = s->ev_pool[i];
            return cudaSuccess;
        };
        for (;;) {
            size_t ev_i = 0;
            cudaEvent_t e_begin, e_last = nullptr;
            CU(get_event(ev_i++, &e_begin));
            CU(cudaEventRecord(e_begin, main_st));
            for (; next_tick < s->rows_cap; next_tick += PIPE_TICKS) {
                fifo_yarn_kernel<<<R, 32, smem, main_st>>>(s->d_desc, s->d_state, s->cc, s->slot_cap, PIPE_TICKS, s->d_rows, s->rows_cap,
                                                          s->d_returns, s->opts.max_ticks);
                CU(cudaGetLastError());
                launches++;
                CU(get_event(ev_i++, &e_last));
                CU(cudaEventRecord(e_last, main_st));
                CU(cudaStreamWaitEvent(s->copy_stream, e_last, 0));
                int64_t w = std::min<int64_t>(PIPE_TICKS, s->rows_cap - next_tick);
                CU(cudaMemcpy2DAsync(s->h_rows + next_tick, sizeof(rlgs_row) * (size_t)s->h_cap, s->d_rows + next_tick,
                                     sizeof(rlgs_row) * (size_t)s->rows_cap, sizeof(rlgs_row) * (size_t)w, (size_t)R,
                                     cudaMemcpyDeviceToHost, s->copy_stream));
            }
            CU(cudaMemcpyAsync(s->h_state, s->d_state, sizeof(RepState) * R, cudaMemcpyDeviceToHost, main_st));
            CU(cudaStreamSynchronize(main_st));
            float ms = 0.f;
            if (e_last) CU(cudaEventElapsedTime(&ms, e_begin, e_last));
            total_ms += ms;
            bool all_done = true, overflow = false;
            for (int r = 0; r < R; ++r) {
                const RepState &z = s->h_state[r];
                if (z.status == RLGS_ERR_CAPACITY && !(s->opts.max_ticks > 0 && z.d >= s->opts.max_ticks)) overflow = true;
                if (!z.done) all_done = false;
            }
            if (overflow) { cudaStreamSynchronize(s->copy_stream); return fail(RLGS_ERR_CAPACITY, "running-job slot table overflow at slot_cap=%d: recreate with a larger opts.slot_cap", s->slot_cap); }
            if (all_done) break;
            // some replica filled the row store: double it (device and host) and keep going
            CU(cudaStreamSynchronize(s->copy_stream));
            int64_t old_cap = s->rows_cap;
            rc = grow_device_rows(s, old_cap * 2, old_cap);
            if (rc) return rc;
            rlgs_row *old_h = s->h_rows; int64_t old_hcap = s->h_cap;
            s->h_rows = nullptr; s->h_cap = 0;
            rc = ensure_host_rows(s, s->rows_cap);
            if (rc) { cudaFreeHost(old_h); return rc; }
            for (int r = 0; r < R; ++r) memcpy(s->h_rows + (size_t)r * s->h_cap, old_h + (size_t)r * old_hcap, sizeof(rlgs_row) * (size_t)old_cap);
            cudaFreeHost(old_h);
        }
        if (eager_jobs)
            for (int k = 0; k < 3; ++k)   // start, end, finish_order (place_off stays on the device)
                CU(cudaMemcpyAsync(s->h_jobs + k * plane, s->d_jobs + k * plane, sizeof(int32_t) * plane, cudaMemcpyDeviceToHost, s->copy_stream));
        CU(cudaStreamSynchronize(s->copy_stream));
        s->rows_fetched = true; s->jobs_fetched = false; s->jobs_partial = eager_jobs;
        s->last_ms = total_ms; s->last_launches = launches;
        for (int r = 0; r < R; ++r) s->h_returns[r] = -s->h_state[r].sum_jct;
        s->ran = true;
        for (int r = 0; r < R; ++r)
            if (s->h_state[r].status != RLGS_OK) return fail(s->h_state[r].status, "replica %d stopped with status %d at tick %d", r, s->h_state[r].status, s->h_state[r].d);
        return RLGS_OK;
    }
    const int budget = s->opts.ticks_per_launch > 0 ? s->opts.ticks_per_launch : (1 << 30);
    float total_ms = 0.f;
    int launches = 0;
    bool all_done = false, clean_single_pass = true;
    while (!all_done) {
        CU(cudaEventRecord(s->ev_fork, main_st));
        for (auto &G : s->groups) {
            CU(cudaStreamWaitEvent(G.stream, s->ev_fork, 0));
            CU(cudaEventRecord(G.k_begin, G.stream));
            fifo_yarn_kernel<<<G.count, 32, smem, G.stream>>>(s->d_desc + G.first, s->d_state + G.first, s->cc, s->slot_cap, budget,
                                                             rows ? s->d_rows + (size_t)G.first * s->rows_cap : nullptr,
                                                             s->rows_cap, s->d_returns + G.first, s->opts.max_ticks);
            CU(cudaGetLastError());
            CU(cudaEventRecord(G.k_end, G.stream));
            CU(cudaMemcpyAsync(s->h_state + G.first, s->d_state + G.first, sizeof(RepState) * G.count, cudaMemcpyDeviceToHost, G.stream));
            if (launches == 0 && s->opts.ticks_per_launch == 0) {
                // optimistic: results of this group go to the host as soon as its kernel ends, while the
                // other groups still compute; redone below if a replica had to be continued
                rc = enqueue_fetch(s, G, eager_rows, eager_jobs, std::min(s->rows_cap, s->h_cap));
                if (rc) return rc;
            }
        }
        launches++;
        for (auto &G : s->groups) CU(cudaStreamSynchronize(G.stream));
        float wave_ms = 0.f;
        for (auto &G : s->groups) {
            float t0 = 0.f, t1 = 0.f;
            CU(cudaEventElapsedTime(&t0, s->ev_fork, G.k_begin));
            CU(cudaEventElapsedTime(&t1, s->ev_fork, G.k_end));
            wave_ms = std::max(wave_ms, t1);
            (void)t0;
        }
        total_ms += wave_ms;
        all_done = true;
        bool overflow = false, rows_full = false;
        for (int r = 0; r < R; ++r) {
            const RepState &z = s->h_state[r];
            if (z.status == RLGS_ERR_CAPACITY && !(s->opts.max_ticks > 0 && z.d >= s->opts.max_ticks)) overflow = true;
            if (!z.done) { all_done = false; if (rows && z.d >= s->rows_cap) rows_full = true; }
        }
        if (overflow) return fail(RLGS_ERR_CAPACITY, "running-job slot table overflow at slot_cap=%d: recreate with a larger opts.slot_cap", s->slot_cap);
        if (!all_done) clean_single_pass = false;
        if (rows_full) {
            rc = grow_device_rows(s, s->rows_cap * 2, s->rows_cap);
            if (rc) return rc;
        }
    }
    if (!clean_single_pass || s->opts.ticks_per_launch != 0) {
        // continued run: fetch everything now
        if (eager_rows) { rc = ensure_host_rows(s, s->rows_cap); if (rc) return rc; }
        for (auto &G : s->groups) {
            int64_t w = 0;
            for (int r = G.first; r < G.first + G.count; ++r) w = std::max<int64_t>(w, s->h_state[r].d);
            rc = enqueue_fetch(s, G, eager_rows, eager_jobs, w);
            if (rc) return rc;
        }
        for (auto &G : s->groups) CU(cudaStreamSynchronize(G.stream));
    }
    s->rows_fetched = eager_rows; s->jobs_fetched = eager_jobs;
    s->last_ms = total_ms; s->last_launches = launches * (int)s->groups.size();
    for (int r = 0; r < R; ++r) s->h_returns[r] = -s->h_state[r].sum_jct;
    s->ran = true;
    for (int r = 0; r < R; ++r)
        if (s->h_state[r].status != RLGS_OK) return fail(s->h_state[r].status, "replica %d stopped with status %d at tick %d", r, s->h_state[r].status, s->h_state[r].d);
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
    const RepState &z = s->h_state[r];
    memset(out, 0, sizeof *out);
    out->n_ticks = z.d; out->makespan = z.d; out->sum_jct = z.sum_jct; out->sum_queued = z.sumQ; out->sum_running = z.sumR;
    out->events = z.events; out->n_jobs = s->h_desc[r].J; out->n_arrived = z.cursor; out->n_started = z.start_seq;
    out->n_finished = z.F; out->max_queued = z.max_q; out->max_running = z.max_r; out->status = z.status; out->done = z.done;
    return RLGS_OK;
}

static int32_t fetch_jobs(rlgs_sim *s) {
    if (s->jobs_fetched) return RLGS_OK;
    if (!s->h_jobs) CU(cudaMallocHost(&s->h_jobs, s->jobs_bytes));
    CU(cudaMemcpy(s->h_jobs, s->d_jobs, s->jobs_bytes, cudaMemcpyDeviceToHost));
    s->jobs_fetched = true;
    return RLGS_OK;
}

static int32_t fetch_rows(rlgs_sim *s) {
    if (s->rows_fetched) return RLGS_OK;
    int32_t rc = ensure_host_rows(s, s->rows_cap);
    if (rc) return rc;
    int64_t w = 0;
    for (int r = 0; r < s->R; ++r) w = std::max<int64_t>(w, s->h_state[r].d);
    if (w > 0)
        CU(cudaMemcpy2D(s->h_rows, sizeof(rlgs_row) * (size_t)s->h_cap, s->d_rows, sizeof(rlgs_row) * (size_t)s->rows_cap,
                        sizeof(rlgs_row) * (size_t)w, (size_t)s->R, cudaMemcpyDeviceToHost));
    s->rows_fetched = true;
    return RLGS_OK;
}

extern "C" int32_t rlgs_read_jobs(rlgs_sim *s, int32_t r, int32_t *finish_order, int32_t *start_tick, int32_t *end_tick,
                                  int32_t *preempt, int32_t *first_node) {
    if (!s) return fail(RLGS_ERR_BAD_ARG, "null handle");
    if (!s->ran) return fail(RLGS_ERR_STATE, "no completed run");
    if (r < 0 || r >= s->R) return fail(RLGS_ERR_BAD_ARG, "replica %d out of range", r);
    CU(cudaSetDevice(s->device));
    if (!(s->jobs_partial && !first_node)) {
        int32_t rc = fetch_jobs(s);
        if (rc) return rc;
    }
    size_t plane = (size_t)s->R * (size_t)s->Jmax, off = (size_t)r * s->Jmax;
    int J = s->h_desc[r].J;
    const int32_t *st = s->h_jobs + off, *en = s->h_jobs + plane + off, *fo = s->h_jobs + 2 * plane + off, *po = s->h_jobs + 3 * plane + off;
    if (start_tick) memcpy(start_tick, st, 4 * (size_t)J);
    if (end_tick) memcpy(end_tick, en, 4 * (size_t)J);
    if (finish_order) memcpy(finish_order, fo, 4 * (size_t)J);
    if (preempt) for (int i = 0; i < J; ++i) preempt[i] = st[i] >= 0 ? 1 : 0;  // Job.migration_count (job.py:171, q6)
    if (first_node) {
        // first placement-log entry of every started job (node index), for placement parity tests
        std::vector<int2> log((size_t)std::max(1, s->h_state[r].log_len));
        if (s->h_state[r].log_len > 0)
            CU(cudaMemcpy(log.data(), s->h_desc[r].place_log, sizeof(int2) * (size_t)s->h_state[r].log_len, cudaMemcpyDeviceToHost));
        for (int i = 0; i < J; ++i) first_node[i] = (po[i] >= 0 && po[i] < s->h_state[r].log_len) ? (log[po[i]].x & 0xffff) : -1;
    }
    return RLGS_OK;
}

extern "C" int32_t rlgs_rows_view(rlgs_sim *s, int32_t r, const rlgs_row **rows, int64_t *count) {
    if (!s || !rows || !count) return fail(RLGS_ERR_BAD_ARG, "null argument");
    if (!s->ran) return fail(RLGS_ERR_STATE, "no completed run");
    if (r < 0 || r >= s->R) return fail(RLGS_ERR_BAD_ARG, "replica %d out of range", r);
    if (s->opts.rows_mode == RLGS_ROWS_NONE) return fail(RLGS_ERR_STATE, "rows were not recorded (opts.rows_mode)");
    CU(cudaSetDevice(s->device));
    int32_t rc = fetch_rows(s);
    if (rc) return rc;
    *rows = s->h_rows + (size_t)r * s->h_cap;
    *count = s->h_state[r].d;
    return RLGS_OK;
}

extern "C" int32_t rlgs_read_rows(rlgs_sim *s, int32_t r, int64_t first, int64_t count, rlgs_row *out) {
    const rlgs_row *rows; int64_t n;
    int32_t rc = rlgs_rows_view(s, r, &rows, &n);
    if (rc) return rc;
    if (!out || first < 0 || count < 0 || first + count > n) return fail(RLGS_ERR_BAD_ARG, "row range [%lld,%lld) out of 0..%lld", (long long)first, (long long)(first + count), (long long)n);
    memcpy(out, rows + first, sizeof(rlgs_row) * (size_t)count);
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
