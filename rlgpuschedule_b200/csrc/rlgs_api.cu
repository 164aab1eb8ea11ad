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
    int32_t n = 0;
    int64_t log_cap = 0;
    int32_t max_arrival = 0;
};

struct rlgs_sim {
    rlgs_cluster_spec spec;
    rlgs_opts opts;
    ClusterConst cc;
    int R = 0;
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    bool own_stream = true;
    void *user_stream = nullptr;
    std::vector<TraceBuf> traces;
    std::vector<int> rep_trace;  // trace id per replica, -1 = none
    std::vector<RepDesc> h_desc;
    RepDesc *d_desc = nullptr;
    RepState *d_state = nullptr;
    std::vector<RepState> h_state;
    std::vector<void *> slabs;   // per load_trace call
    int *d_done = nullptr;       // unused counter slot (kept for env)
    int slot_cap = 0;
    int chunk_ticks = 0;
    // rows
    rlgs_row *d_rows[2] = {nullptr, nullptr};
    rlgs_row *h_rows = nullptr;  // pinned [R][h_cap]
    int64_t h_cap = 0;
    std::vector<int64_t> n_rows;
    // job-table mirror
    int32_t *h_jobs = nullptr;   // pinned mirror of the per-replica output arrays
    size_t h_jobs_bytes = 0;
    bool jobs_fetched = false;
    int32_t Jmax = 0;
    int32_t *d_jobs = nullptr;   // [4][R][Jmax]: start, end, finish_order, place_off
    int64_t *d_returns = nullptr;
    int64_t *h_returns = nullptr;
    bool ran = false;
    float last_ms = 0.f;
    int last_launches = 0;
    std::vector<cudaEvent_t> ev;
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
    s->slot_cap = opts->slot_cap > 0 ? opts->slot_cap : std::min(256, std::max(32, s->cc.D));
    s->slot_cap = (s->slot_cap + 31) & ~31;
    s->chunk_ticks = opts->chunk_ticks > 0 ? opts->chunk_ticks : 2048;
    s->rep_trace.assign(s->R, -1);
    s->h_desc.assign(s->R, RepDesc{});
    s->h_state.assign(s->R, RepState{});
    s->n_rows.assign(s->R, 0);
    cudaError_t ce;
    if ((ce = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaMalloc(&s->d_desc, sizeof(RepDesc) * s->R)) != cudaSuccess ||
        (ce = cudaMalloc(&s->d_state, sizeof(RepState) * s->R)) != cudaSuccess ||
        (ce = cudaMalloc(&s->d_returns, sizeof(int64_t) * s->R)) != cudaSuccess ||
        (ce = cudaMallocHost(&s->h_returns, sizeof(int64_t) * s->R)) != cudaSuccess) {
        rlgs_destroy(s);
        return fail(RLGS_ERR_CUDA, "rlgs_create: %s", cudaGetErrorString(ce));
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
    cudaFree(s->d_desc); cudaFree(s->d_state); cudaFree(s->d_rows[0]); cudaFree(s->d_rows[1]);
    cudaFree(s->d_jobs); cudaFree(s->d_returns);
    if (s->h_rows) cudaFreeHost(s->h_rows);
    if (s->h_jobs) cudaFreeHost(s->h_jobs);
    if (s->h_returns) cudaFreeHost(s->h_returns);
    for (auto ev : s->ev) cudaEventDestroy(ev);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
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
    tb.max_arrival = prev;
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
        s->h_jobs_bytes = sizeof(int32_t) * 4 * (size_t)s->R * (size_t)Jmax;
        CU(cudaMalloc(&s->d_jobs, s->h_jobs_bytes));
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

static int32_t ensure_host_rows(rlgs_sim *s, int64_t need_ticks) {
    if (need_ticks <= s->h_cap) return RLGS_OK;
    int64_t cap = std::max<int64_t>(need_ticks, s->h_cap * 2);
    cap = (cap + s->chunk_ticks - 1) / s->chunk_ticks * s->chunk_ticks;
    rlgs_row *nw = nullptr;
    CU(cudaStreamSynchronize(s->copy_stream));
    CU(cudaMallocHost(&nw, sizeof(rlgs_row) * (size_t)cap * (size_t)s->R));
    if (s->h_rows) {
        for (int r = 0; r < s->R; ++r)
            memcpy(nw + (size_t)r * cap, s->h_rows + (size_t)r * s->h_cap, sizeof(rlgs_row) * (size_t)s->h_cap);
        cudaFreeHost(s->h_rows);
    }
    s->h_rows = nw; s->h_cap = cap;
    return RLGS_OK;
}

extern "C" int32_t rlgs_run(rlgs_sim *s) {
    if (!s) return fail(RLGS_ERR_BAD_ARG, "null handle");
    CU(cudaSetDevice(s->device));
    int32_t rc = setup_job_arrays(s);
    if (rc) return rc;
    cudaStream_t st = s->user_stream ? (cudaStream_t)s->user_stream : s->stream;
    const bool rows = s->opts.rows_mode == RLGS_ROWS_FULL;
    const int R = s->R;

    for (int attempt = 0;; ++attempt) {
        size_t smem = fifo_smem_bytes(s->cc.N, s->slot_cap);
        if (smem > 227 * 1024) return fail(RLGS_ERR_CAPACITY, "cluster state needs %zu B of shared memory per replica (> 227 KB)", smem);
        CU(cudaFuncSetAttribute(fifo_yarn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // reset state
        int32_t max_arrival = 0;
        for (int r = 0; r < R; ++r) {
            RepState z; memset(&z, 0, sizeof z);
            z.head = s->h_desc[r].J; z.idle_nodes = s->cc.N;
            z.n_free_nodes = (s->cc.cpu_cap > 0 || s->cc.mem_cap > 0) ? s->cc.N : 0;
            z.free_hint = -1;
            s->h_state[r] = z;
            max_arrival = std::max(max_arrival, s->traces[s->rep_trace[r]].max_arrival);
        }
        CU(cudaMemcpyAsync(s->d_state, s->h_state.data(), sizeof(RepState) * R, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(s->d_desc, s->h_desc.data(), sizeof(RepDesc) * R, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(s->d_jobs, 0xff, s->h_jobs_bytes, st));
        s->jobs_fetched = false;
        int chunk = rows ? s->chunk_ticks : (1 << 30);
        if (rows) {
            for (int b = 0; b < 2; ++b)
                if (!s->d_rows[b]) CU(cudaMalloc(&s->d_rows[b], sizeof(rlgs_row) * (size_t)chunk * (size_t)R));
            rc = ensure_host_rows(s, (int64_t)max_arrival + 2 * chunk);
            if (rc) return rc;
        }
        // events for device timing
        size_t ev_used = 0;
        auto next_event = [&](cudaEvent_t *out) -> cudaError_t {
            if (ev_used == s->ev.size()) { cudaEvent_t e; cudaError_t ce = cudaEventCreate(&e); if (ce != cudaSuccess) return ce; s->ev.push_back(e); }
            *out = s->ev[ev_used++];
            return cudaSuccess;
        };
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
        bool all_done = false, overflow = false;
        int64_t base_tick = 0;
        int launches = 0;
        cudaEvent_t copy_done[2] = {nullptr, nullptr};
        while (!all_done) {
            int b = launches & 1;
            if (rows) {
                // the copy that read d_rows[b] two launches ago must be finished before we overwrite it
                if (copy_done[b]) CU(cudaStreamWaitEvent(st, copy_done[b], 0));
            }
            cudaEvent_t e0, e1;
            CU(next_event(&e0)); CU(next_event(&e1));
            CU(cudaEventRecord(e0, st));
            fifo_yarn_kernel<<<R, 32, smem, st>>>(s->d_desc, s->d_state, s->cc, s->slot_cap, chunk, rows ? s->d_rows[b] : nullptr,
                                                  chunk, s->d_returns, s->opts.max_ticks);
            CU(cudaGetLastError());
            CU(cudaEventRecord(e1, st));
            spans.push_back({e0, e1});
            launches++;
            CU(cudaMemcpyAsync(s->h_state.data(), s->d_state, sizeof(RepState) * R, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            all_done = true;
            int64_t max_d = 0;
            for (int r = 0; r < R; ++r) {
                const RepState &z = s->h_state[r];
                if (!z.done) all_done = false;
                if (z.status == RLGS_ERR_CAPACITY && !(s->opts.max_ticks > 0 && z.d >= s->opts.max_ticks)) overflow = true;
                max_d = std::max<int64_t>(max_d, z.d);
            }
            if (overflow) break;
            if (rows) {
                int64_t width = std::min<int64_t>(chunk, max_d - base_tick);
                if (width > 0) {
                    rc = ensure_host_rows(s, base_tick + chunk);
                    if (rc) return rc;
                    if (!copy_done[b]) CU(next_event(&copy_done[b]));
                    CU(cudaMemcpy2DAsync(s->h_rows + base_tick, sizeof(rlgs_row) * (size_t)s->h_cap, s->d_rows[b],
                                         sizeof(rlgs_row) * (size_t)chunk, sizeof(rlgs_row) * (size_t)width, (size_t)R,
                                         cudaMemcpyDeviceToHost, s->copy_stream));
                    CU(cudaEventRecord(copy_done[b], s->copy_stream));
                }
                base_tick += chunk;
            }
        }
        if (overflow) {
            // a replica ran out of running-job slots: double the on-chip table and start over
            if (s->slot_cap >= s->cc.D || attempt > 8) return fail(RLGS_ERR_CAPACITY, "running-job slot table overflow at slot_cap=%d", s->slot_cap);
            CU(cudaStreamSynchronize(s->copy_stream));
            int new_cap = std::min((s->cc.D + 31) & ~31, s->slot_cap * 2);
            return fail(RLGS_ERR_CAPACITY, "running-job slot table overflow at slot_cap=%d: recreate with opts.slot_cap>=%d", s->slot_cap, new_cap);
        }
        CU(cudaStreamSynchronize(s->copy_stream));
        float ms = 0.f;
        for (auto &sp : spans) { float t = 0.f; CU(cudaEventElapsedTime(&t, sp.first, sp.second)); ms += t; }
        s->last_ms = ms; s->last_launches = launches;
        for (int r = 0; r < R; ++r) { s->n_rows[r] = s->h_state[r].d; s->h_returns[r] = -s->h_state[r].sum_jct; }
        break;
    }
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
    if (!s->h_jobs) CU(cudaMallocHost(&s->h_jobs, s->h_jobs_bytes));
    CU(cudaMemcpy(s->h_jobs, s->d_jobs, s->h_jobs_bytes, cudaMemcpyDeviceToHost));
    s->jobs_fetched = true;
    return RLGS_OK;
}

extern "C" int32_t rlgs_read_jobs(rlgs_sim *s, int32_t r, int32_t *finish_order, int32_t *start_tick, int32_t *end_tick,
                                  int32_t *preempt, int32_t *first_node) {
    if (!s) return fail(RLGS_ERR_BAD_ARG, "null handle");
    if (!s->ran) return fail(RLGS_ERR_STATE, "no completed run");
    if (r < 0 || r >= s->R) return fail(RLGS_ERR_BAD_ARG, "replica %d out of range", r);
    CU(cudaSetDevice(s->device));
    int32_t rc = fetch_jobs(s);
    if (rc) return rc;
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
    if (s->opts.rows_mode != RLGS_ROWS_FULL) return fail(RLGS_ERR_STATE, "rows were not recorded (opts.rows_mode)");
    *rows = s->h_rows + (size_t)r * s->h_cap;
    *count = s->n_rows[r];
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
