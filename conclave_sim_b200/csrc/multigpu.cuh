// multigpu.cuh — the multi-GPU half of the C ABI (included by cabi.cu after csim_ctx / csim_run are defined).
//
// Replaces the process-pool fan-out of the reference (src/simulate.py:330-351, `-w/--workers` :231): the trials of one call
// are sharded over the GPUs of one box by disjoint, contiguous sim-id ranges (a trial's Philox stream is keyed by its global
// sim id, so the shards of N devices reproduce a single-device run bit for bit), and the ONLY exchange is one
// ncclAllReduce(sum, int64) of winner_hist | rounds_hist | totals, enqueued on each device's stream directly behind its
// trial kernel.  Per-trial rows never cross devices: each device copies its rows into the caller's arrays.
//
// NCCL is bound at run time with dlopen — the library has no link-time dependency on it, a single-GPU host needs no NCCL
// at all, and inside a Python process that already loaded the wheel-bundled libnccl.so.2 (torch) the same copy is reused.
#pragma once
#include <dlfcn.h>

// the few NCCL declarations this file needs (nccl.h of NCCL 2.x: stable ABI since 2.0)
typedef struct csimNcclComm* csim_ncclComm_t;
typedef struct { char internal[CSIM_NCCL_ID_BYTES]; } csim_ncclUniqueId;
enum { CSIM_NCCL_SUCCESS = 0, CSIM_NCCL_INT64 = 4, CSIM_NCCL_SUM = 0 };

struct NcclApi {
    void* handle = nullptr;
    int version = 0;
    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(csim_ncclUniqueId*) = nullptr;
    int (*CommInitRank)(csim_ncclComm_t*, int, csim_ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(csim_ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(csim_ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, csim_ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

#define NCCL_TRY(expr)                                                                                  \
    do {                                                                                                \
        int _r = (expr);                                                                                \
        if (_r != CSIM_NCCL_SUCCESS)                                                                    \
            return fail(CSIM_ERR_CUDA, "%s failed: NCCL error %d (%s)", #expr, _r,                      \
                        g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");                       \
    } while (0)

extern "C" csim_status csim_nccl_load(const char* path) {
    if (g_nccl.handle) return CSIM_OK;
    const char* env = getenv("CSIM_NCCL_LIB");
    const char* tries[4] = {path, env, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    std::string errs;
    for (const char* t : tries) {
        if (!t || !*t) continue;
        h = dlopen(t, RTLD_NOW | RTLD_LOCAL);
        if (h) break;
        const char* e = dlerror();
        errs += std::string(t) + ": " + (e ? e : "?") + "; ";
    }
    if (!h) return fail(CSIM_ERR_UNSUPPORTED, "csim_nccl_load: cannot load NCCL (%s) — multi-GPU calls need libnccl.so.2", errs.c_str());
    NcclApi api;
    api.handle = h;
    struct { const char* name; void** slot; } syms[] = {
        {"ncclGetVersion", (void**)&api.GetVersion},     {"ncclGetUniqueId", (void**)&api.GetUniqueId},
        {"ncclCommInitRank", (void**)&api.CommInitRank}, {"ncclCommInitAll", (void**)&api.CommInitAll},
        {"ncclCommDestroy", (void**)&api.CommDestroy},   {"ncclAllReduce", (void**)&api.AllReduce},
        {"ncclGroupStart", (void**)&api.GroupStart},     {"ncclGroupEnd", (void**)&api.GroupEnd},
        {"ncclGetErrorString", (void**)&api.GetErrorString}};
    for (auto& sy : syms) {
        *sy.slot = dlsym(h, sy.name);
        if (!*sy.slot) { dlclose(h); return fail(CSIM_ERR_UNSUPPORTED, "csim_nccl_load: %s not found in the NCCL library", sy.name); }
    }
    api.GetVersion(&api.version);
    g_nccl = api;
    return CSIM_OK;
}

extern "C" int32_t csim_nccl_version(void) { return g_nccl.handle ? g_nccl.version : 0; }

extern "C" csim_status csim_nccl_unique_id(uint8_t id[CSIM_NCCL_ID_BYTES]) {
    if (!id) return fail(CSIM_ERR_INVALID, "csim_nccl_unique_id: id is NULL");
    csim_status s = csim_nccl_load(nullptr);
    if (s != CSIM_OK) return s;
    csim_ncclUniqueId u;
    NCCL_TRY(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, CSIM_NCCL_ID_BYTES);
    return CSIM_OK;
}

extern "C" csim_status csim_nccl_comm_create(csim_ctx* c, const uint8_t id[CSIM_NCCL_ID_BYTES], int32_t n_ranks, int32_t rank, void** comm) {
    if (!c || !id || !comm) return fail(CSIM_ERR_INVALID, "csim_nccl_comm_create: NULL argument");
    *comm = nullptr;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(CSIM_ERR_INVALID, "csim_nccl_comm_create: rank %d of %d", rank, n_ranks);
    csim_status s = csim_nccl_load(nullptr);
    if (s != CSIM_OK) return s;
    DeviceGuard guard;
    CU_TRY(cudaSetDevice(c->device));
    csim_ncclUniqueId u;
    memcpy(u.internal, id, CSIM_NCCL_ID_BYTES);
    csim_ncclComm_t cm = nullptr;
    NCCL_TRY(g_nccl.CommInitRank(&cm, n_ranks, u, rank));
    *comm = cm;
    return CSIM_OK;
}

extern "C" csim_status csim_nccl_comm_destroy(void* comm) {
    if (!comm) return CSIM_OK;
    if (!g_nccl.handle) return fail(CSIM_ERR_STATE, "csim_nccl_comm_destroy: NCCL is not loaded");
    NCCL_TRY(g_nccl.CommDestroy((csim_ncclComm_t)comm));
    return CSIM_OK;
}

// the all-reduce(s) of one rank's histograms, to be called between ncclGroupStart / ncclGroupEnd when several devices are
// driven by one thread
static csim_status allreduce_enqueue(void* comm, int64_t* wh, int64_t* rh, int64_t* tt, size_t n_w, size_t n_r, size_t n_t, cudaStream_t st,
                                     bool* grouped) {
    const bool contiguous = (!rh || rh == wh + n_w) && (!tt || (rh && tt == rh + n_r));
    *grouped = false;
    if (contiguous) {       // winner_hist | rounds_hist | totals as ONE message
        NCCL_TRY(g_nccl.AllReduce(wh, wh, n_w + (rh ? n_r : 0) + (tt ? n_t : 0), CSIM_NCCL_INT64, CSIM_NCCL_SUM, (csim_ncclComm_t)comm, st));
        return CSIM_OK;
    }
    *grouped = true;
    NCCL_TRY(g_nccl.AllReduce(wh, wh, n_w, CSIM_NCCL_INT64, CSIM_NCCL_SUM, (csim_ncclComm_t)comm, st));
    if (rh) NCCL_TRY(g_nccl.AllReduce(rh, rh, n_r, CSIM_NCCL_INT64, CSIM_NCCL_SUM, (csim_ncclComm_t)comm, st));
    if (tt) NCCL_TRY(g_nccl.AllReduce(tt, tt, n_t, CSIM_NCCL_INT64, CSIM_NCCL_SUM, (csim_ncclComm_t)comm, st));
    return CSIM_OK;
}

extern "C" csim_status csim_allreduce_hists(csim_ctx* c, void* comm, int64_t* winner_hist, int64_t* rounds_hist, int64_t* totals,
                                            int32_t n_param_sets, int32_t max_rounds, void* stream) {
    if (!c || !comm || !winner_hist) return fail(CSIM_ERR_INVALID, "csim_allreduce_hists: NULL argument");
    if (c->E <= 0) return fail(CSIM_ERR_STATE, "csim_allreduce_hists: roster not uploaded (the histogram width is E + 1)");
    if (n_param_sets < 1 || max_rounds < 0) return fail(CSIM_ERR_INVALID, "csim_allreduce_hists: bad shape");
    csim_status s = csim_nccl_load(nullptr);
    if (s != CSIM_OK) return s;
    DeviceGuard guard;
    CU_TRY(cudaSetDevice(c->device));
    const size_t n_w = (size_t)n_param_sets * (c->E + 1), n_r = (size_t)n_param_sets * (max_rounds + 1), n_t = (size_t)n_param_sets * 2;
    const bool contiguous = (!rounds_hist || rounds_hist == winner_hist + n_w) && (!totals || (rounds_hist && totals == rounds_hist + n_r));
    if (!contiguous) NCCL_TRY(g_nccl.GroupStart());
    bool grouped = false;
    s = allreduce_enqueue(comm, winner_hist, rounds_hist, totals, n_w, n_r, n_t, (cudaStream_t)stream, &grouped);
    if (!contiguous) NCCL_TRY(g_nccl.GroupEnd());
    return s;
}

extern "C" void csim_shard_range(uint64_t n, int32_t rank, int32_t n_ranks, uint64_t* begin, uint64_t* count) {
    if (n_ranks < 1) n_ranks = 1;
    if (rank < 0) rank = 0;
    const uint64_t base = n / (uint64_t)n_ranks, rem = n % (uint64_t)n_ranks, r = (uint64_t)rank;
    if (begin) *begin = r * base + std::min(r, rem);
    if (count) *count = base + (r < rem ? 1 : 0);
}

// ---------------------------------------------------------------------------------------------
// csim_group: one process, one context + stream + communicator per device
// ---------------------------------------------------------------------------------------------
struct csim_group {
    std::vector<csim_ctx*> ctx;
    std::vector<cudaStream_t> stream;
    std::vector<csim_ncclComm_t> comm;          // empty when the group has one device
    std::vector<DevBuf> hist;                    // winner_hist | rounds_hist | totals of each device, contiguous
};

extern "C" csim_status csim_group_destroy(csim_group* g) {
    if (!g) return CSIM_OK;
    DeviceGuard guard;
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        if (g->ctx[i]) { cudaSetDevice(g->ctx[i]->device); cudaDeviceSynchronize(); }
        if (i < g->comm.size() && g->comm[i] && g_nccl.handle) g_nccl.CommDestroy(g->comm[i]);
        if (i < g->hist.size()) g->hist[i].release();
        if (i < g->stream.size() && g->stream[i]) cudaStreamDestroy(g->stream[i]);
        csim_destroy(g->ctx[i]);
    }
    delete g;
    return CSIM_OK;
}

extern "C" csim_status csim_group_create(csim_group** out, const int32_t* devices, int32_t n_devices) {
    if (!out) return fail(CSIM_ERR_INVALID, "csim_group_create: out is NULL");
    *out = nullptr;
    if (!devices || n_devices < 1) return fail(CSIM_ERR_INVALID, "csim_group_create: need at least one device");
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(CSIM_ERR_INVALID, "csim_group_create: device %d listed twice", devices[i]);
    if (n_devices > 1) {
        csim_status s = csim_nccl_load(nullptr);
        if (s != CSIM_OK) return s;
    }
    DeviceGuard guard;
    csim_group* g = new csim_group();
    g->ctx.assign(n_devices, nullptr);
    g->stream.assign(n_devices, nullptr);
    g->hist.resize(n_devices);
    for (int i = 0; i < n_devices; ++i) {
        csim_status s = csim_create(&g->ctx[i], devices[i]);
        if (s != CSIM_OK) { std::string keep = g_err; csim_group_destroy(g); g_err = keep; return s; }
        cudaSetDevice(devices[i]);
        cudaError_t e = cudaStreamCreateWithFlags(&g->stream[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) { csim_group_destroy(g); return fail(CSIM_ERR_CUDA, "cudaStreamCreate on device %d: %s", devices[i], cudaGetErrorString(e)); }
    }
    if (n_devices > 1) {
        g->comm.assign(n_devices, nullptr);
        std::vector<int> devs(devices, devices + n_devices);
        int r = g_nccl.CommInitAll(g->comm.data(), n_devices, devs.data());
        if (r != CSIM_NCCL_SUCCESS) {
            g->comm.clear();
            csim_group_destroy(g);
            return fail(CSIM_ERR_CUDA, "ncclCommInitAll over %d devices failed: NCCL error %d (%s)", n_devices, r, g_nccl.GetErrorString(r));
        }
    }
    *out = g;
    return CSIM_OK;
}

extern "C" int32_t csim_group_size(const csim_group* g) { return g ? (int32_t)g->ctx.size() : 0; }
extern "C" csim_ctx* csim_group_ctx(csim_group* g, int32_t i) { return (g && i >= 0 && i < (int32_t)g->ctx.size()) ? g->ctx[i] : nullptr; }

extern "C" csim_status csim_group_upload_roster(csim_group* g, int32_t E, const double* x, const int32_t* region, const uint8_t* pap) {
    if (!g) return fail(CSIM_ERR_INVALID, "csim_group_upload_roster: group is NULL");
    for (csim_ctx* c : g->ctx) {
        csim_status s = csim_upload_roster(c, E, x, region, pap);
        if (s != CSIM_OK) return s;
    }
    return CSIM_OK;
}

// host [set][n_sims] rows  <->  device [set][count] rows: one strided copy per array
static cudaError_t copy_rows(void* host, const void* dev, size_t elem_bytes, uint64_t n_sims, uint64_t begin, uint64_t count, int n_sets,
                             bool to_host, cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    char* h = (char*)host + begin * elem_bytes;
    if (to_host) return cudaMemcpy2DAsync(h, n_sims * elem_bytes, dev, count * elem_bytes, count * elem_bytes, n_sets, cudaMemcpyDeviceToHost, st);
    return cudaMemcpy2DAsync((void*)dev, count * elem_bytes, h, n_sims * elem_bytes, count * elem_bytes, n_sets, cudaMemcpyHostToDevice, st);
}

extern "C" csim_status csim_group_run(csim_group* g, const csim_run_args* ra, float* kernel_ms) {
    if (!g || !ra) return fail(CSIM_ERR_INVALID, "csim_group_run: NULL argument");
    if (ra->buffers_on_device) return fail(CSIM_ERR_INVALID, "csim_group_run takes host buffers (each device has its own memory)");
    const int nd = (int)g->ctx.size();
    if (nd == 1 && !kernel_ms) {                      // one device: exactly csim_run on the group's stream
        csim_run_args one = *ra;
        one.stream = g->stream[0];
        return csim_run(g->ctx[0], &one);
    }
    if (ra->n_param_sets < 0) return fail(CSIM_ERR_INVALID, "csim_group_run: n_param_sets < 0");
    const uint64_t n = (uint64_t)ra->n_param_sets * ra->n_sims;
    if (n == 0) return CSIM_OK;
    if (!ra->params) return fail(CSIM_ERR_INVALID, "csim_group_run: params is NULL");
    if (!ra->winner || !ra->rounds) return fail(CSIM_ERR_INVALID, "csim_group_run: winner and rounds outputs are required");
    const int E = g->ctx[0]->E, R = ra->params[0].max_rounds, ns = ra->n_param_sets;
    if (E <= 0) return fail(CSIM_ERR_STATE, "csim_group_run: roster not uploaded");
    if (R < 0) return fail(CSIM_ERR_INVALID, "csim_group_run: max_rounds < 0");
    const size_t n_w = (size_t)ns * (E + 1), n_r = (size_t)ns * (R + 1), n_t = (size_t)ns * 2, n_h = n_w + n_r + n_t;
    DeviceGuard guard;

    // 1. every device: stage, enqueue its shard (device-buffer csim_run only ENQUEUES), histograms into its own buffer
    std::vector<uint64_t> beg(nd), cnt(nd);
    for (int i = 0; i < nd; ++i) {
        csim_ctx* c = g->ctx[i];
        cudaStream_t st = g->stream[i];
        csim_shard_range(ra->n_sims, i, nd, &beg[i], &cnt[i]);
        CU_TRY(cudaSetDevice(c->device));
        CU_TRY(ctx_wait_prev(c, st));
        CU_TRY(g->hist[i].reserve(8 * n_h));
        CU_TRY(cudaMemsetAsync(g->hist[i].p, 0, 8 * n_h, st));
        if (cnt[i] == 0) continue;                                  // fewer trials than devices: this one only joins the all-reduce
        const uint64_t nl = (uint64_t)ns * cnt[i];
        csim_run_args a = *ra;
        a.buffers_on_device = 1; a.stream = st;
        a.sim_id_begin = ra->sim_id_begin + beg[i]; a.n_sims = cnt[i];
        a.auto_batch_sims = ra->auto_batch_sims ? ra->auto_batch_sims : ra->n_sims;      // AUTO chooses for the whole job, not for the shard
        CU_TRY(c->o_winner.reserve(4 * nl)); a.winner = (int32_t*)c->o_winner.p;
        CU_TRY(c->o_rounds.reserve(4 * nl)); a.rounds = (int32_t*)c->o_rounds.p;
        a.reason = nullptr; a.final_votes = nullptr; a.round_tallies = nullptr; a.uniform_tape = nullptr;
        if (ra->reason) { CU_TRY(c->o_reason.reserve(nl)); a.reason = (uint8_t*)c->o_reason.p; }
        if (ra->final_votes) { CU_TRY(c->o_final.reserve(4 * nl * E)); a.final_votes = (int32_t*)c->o_final.p; }
        if (ra->round_tallies) { CU_TRY(c->o_tallies.reserve(4 * nl * (size_t)R * E)); a.round_tallies = (int32_t*)c->o_tallies.p; }
        if (ra->uniform_tape) {
            CU_TRY(c->i_tape.reserve(8 * nl * (size_t)R * E));
            CU_TRY(copy_rows((void*)ra->uniform_tape, c->i_tape.p, 8 * (size_t)R * E, ra->n_sims, beg[i], cnt[i], ns, false, st));
            a.uniform_tape = (const double*)c->i_tape.p;
        }
        int64_t* h = (int64_t*)g->hist[i].p;
        a.winner_hist = h; a.rounds_hist = h + n_w; a.totals = h + n_w + n_r;
        csim_status s = csim_run(c, &a);
        if (s != CSIM_OK) {
            for (int j = 0; j <= i; ++j) { cudaSetDevice(g->ctx[j]->device); cudaStreamSynchronize(g->stream[j]); }
            return s;
        }
    }
    // 2. ONE all-reduce of the packed histograms, enqueued behind each device's trial kernel
    if (nd > 1) {
        NCCL_TRY(g_nccl.GroupStart());
        for (int i = 0; i < nd; ++i) {
            int r = g_nccl.AllReduce(g->hist[i].p, g->hist[i].p, n_h, CSIM_NCCL_INT64, CSIM_NCCL_SUM, g->comm[i], g->stream[i]);
            if (r != CSIM_NCCL_SUCCESS) { g_nccl.GroupEnd(); return fail(CSIM_ERR_CUDA, "ncclAllReduce on device %d: NCCL error %d (%s)", g->ctx[i]->device, r, g_nccl.GetErrorString(r)); }
        }
        NCCL_TRY(g_nccl.GroupEnd());
    }
    // 3. per-trial rows straight into the caller's [set][sim] arrays; the reduced histograms from device 0
    std::vector<int64_t> hh(n_h);
    for (int i = 0; i < nd; ++i) {
        csim_ctx* c = g->ctx[i];
        cudaStream_t st = g->stream[i];
        CU_TRY(cudaSetDevice(c->device));
        if (cnt[i]) {
            CU_TRY(copy_rows(ra->winner, c->o_winner.p, 4, ra->n_sims, beg[i], cnt[i], ns, true, st));
            CU_TRY(copy_rows(ra->rounds, c->o_rounds.p, 4, ra->n_sims, beg[i], cnt[i], ns, true, st));
            if (ra->reason) CU_TRY(copy_rows(ra->reason, c->o_reason.p, 1, ra->n_sims, beg[i], cnt[i], ns, true, st));
            if (ra->final_votes) CU_TRY(copy_rows(ra->final_votes, c->o_final.p, 4 * (size_t)E, ra->n_sims, beg[i], cnt[i], ns, true, st));
            if (ra->round_tallies) CU_TRY(copy_rows(ra->round_tallies, c->o_tallies.p, 4 * (size_t)R * E, ra->n_sims, beg[i], cnt[i], ns, true, st));
        }
        if (i == 0) CU_TRY(cudaMemcpyAsync(hh.data(), g->hist[0].p, 8 * n_h, cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx_mark_busy(c, st));
    }
    for (int i = 0; i < nd; ++i) {
        CU_TRY(cudaSetDevice(g->ctx[i]->device));
        CU_TRY(cudaStreamSynchronize(g->stream[i]));
        if (kernel_ms) {
            float ms = -1.f;
            if (cnt[i] && cudaEventElapsedTime(&ms, g->ctx[i]->ev0, g->ctx[i]->ev1) == cudaSuccess) g->ctx[i]->last_ms = ms;
            kernel_ms[i] = cnt[i] ? ms : 0.f;
        }
    }
    if (ra->winner_hist) for (size_t j = 0; j < n_w; ++j) ra->winner_hist[j] += hh[j];
    if (ra->rounds_hist) for (size_t j = 0; j < n_r; ++j) ra->rounds_hist[j] += hh[n_w + j];
    if (ra->totals) for (size_t j = 0; j < n_t; ++j) ra->totals[j] += hh[n_w + n_r + j];
    return CSIM_OK;
}
