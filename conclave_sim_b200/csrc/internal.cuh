// internal.cuh — host-side structures shared by the translation units of the library (not part of the C ABI).
//
// The library is built from one translation unit per engine (eng_*.cu: the trial kernel, its table builders and its
// launcher) plus cabi.cu (the entry points of include/conclave_b200.h, engine choice, probability matrix, multi-GPU), so
// that a kernel edit recompiles one file and the files compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "conclave_b200.h"
#include "common.cuh"

#ifndef CSIM_GRID_TILE
#define CSIM_GRID_TILE 65535           // gridDim.y / .z limit: the table-build launches are tiled over (set, round)
#endif

// sets the calling thread's csim_last_error() message and returns `st` (defined in cabi.cu)
csim_status fail(csim_status st, const char* fmt, ...);

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return fail(CSIM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n == 0) n = 16;
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// csim_engine_probs: while one is set on the context, the launchers pick the PROBE instantiation of their trial kernel
struct ProbeReq {
    int round = 1;
    const int32_t* prev_vote = nullptr;     // device copies
    const int32_t* prev_count = nullptr;
    const uint8_t* fatigued = nullptr;
    float* scores = nullptr;                // [E][E]
    float* prefix = nullptr;                // [E][cap]
    int cap = 0;
    int nch = 0, chunk = 0;                 // filled by the launcher
};

struct csim_ctx {
    ProbeReq* probe = nullptr;
    int device = 0;
    cudaDeviceProp prop{};
    int E = 0, Epad = 0, G = 0;
    DevBuf x64, x32, region, pap;
    DevBuf sets, betas, nbeta, W, tab, cdf64, cdfk64, cdf32, cdfk32, ptab, accm, accb;
    DevBuf work;                        // per-set trial counters of the dynamically scheduled kernels (WorkQueue)
    double x_mid = 0.0, x_half = 0.0;   // centre and half-range of ideology_score (factorised exp)
    // host-buffer mode staging
    DevBuf o_winner, o_rounds, o_reason, o_final, o_tallies, o_whist, o_rhist, o_totals, i_tape;
    // prob-matrix staging
    DevBuf p_prev, p_cnt, p_fat, p_shares, p_bw, p_out, p_cand, p_pref;
    // page-locked staging of the per-batch parameter tables: the copies to the device are then truly asynchronous, so a
    // device-buffer csim_run returns as soon as its work is enqueued and the host can run ahead of the GPU (copies from
    // pageable memory block until everything earlier in the stream has finished)
    void* h_stage = nullptr; size_t h_stage_cap = 0;
    cudaEvent_t ev_stage = nullptr; bool stage_valid = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_busy = nullptr;      // recorded behind the last work a call enqueued: the next call's stream waits on it
    bool ev_valid = false, busy_valid = false;
    int64_t launches = 0;
    float last_ms = -1.f;
    int last_engine = CSIM_ENGINE_AUTO, last_precision = CSIM_FP32, last_fell = 0;
};

// Entry points switch to the context's device; the caller's current device is restored on every exit path.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// A context serves one stream at a time (its per-batch scratch is overwritten by the next call): every call makes its
// stream wait on what the previous call enqueued, and leaves its own marker behind.
static inline cudaError_t ctx_wait_prev(csim_ctx* c, cudaStream_t st) {
    return c->busy_valid ? cudaStreamWaitEvent(st, c->ev_busy, 0) : cudaSuccess;
}
static inline cudaError_t ctx_mark_busy(csim_ctx* c, cudaStream_t st) {
    cudaError_t e = cudaEventRecord(c->ev_busy, st);
    if (e == cudaSuccess) c->busy_valid = true;
    return e;
}

static inline RosterDev roster_dev(const csim_ctx* c) {
    RosterDev r;
    r.E = c->E; r.Epad = c->Epad; r.G = c->G;
    r.x64 = (const double*)c->x64.p; r.x32 = (const float*)c->x32.p;
    r.region = (const int32_t*)c->region.p; r.pap = (const uint8_t*)c->pap.p;
    return r;
}

// Zeroed per-set trial counters for one launch of a dynamically scheduled kernel (enqueued on `st` ahead of the kernel), and
// the number of trials a warp claims per fetch: small enough that the last claims end together (<= 1/128 of a warp's share, at most 4),
// large enough that the atomics are noise.
static inline csim_status work_queue(csim_ctx* c, int n_sets, uint64_t n_sims, uint64_t n_warps, cudaStream_t st, WorkQueue* wq) {
    const size_t bytes = sizeof(unsigned long long) * (size_t)std::max(n_sets, 1);
    CU_TRY(c->work.reserve(bytes));
    CU_TRY(cudaMemsetAsync(c->work.p, 0, bytes, st));
    wq->next = (unsigned long long*)c->work.p;
    const uint64_t per_warp = n_sims * (uint64_t)std::max(n_sets, 1) / std::max<uint64_t>(n_warps, 1);
    wq->chunk = (int)std::max<uint64_t>(1, std::min<uint64_t>(per_warp / 128, 4));
    if (const char* e = getenv("CSIM_CHUNK")) wq->chunk = std::max(1, atoi(e));
    return CSIM_OK;
}

template <class K>
static csim_status persistent_grid(csim_ctx* c, K kernel, int threads, size_t smem, uint64_t n_trials, int warps_per_cta, int* grid) {
    CU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) return fail(CSIM_ERR_UNSUPPORTED, "kernel does not fit on an SM (threads=%d, smem=%zu)", threads, smem);
    const uint64_t want = (n_trials + warps_per_cta - 1) / warps_per_cta;
    const uint64_t full = (uint64_t)c->prop.multiProcessorCount * per_sm;     // a multiple of the SM count
    *grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, full));
    return CSIM_OK;
}

// launchers, one per engine (eng_*.cu).  Each enqueues its table builders and its trial kernel on `st`, records the
// context's timing events around the trial kernel and returns CSIM_ERR_UNSUPPORTED (with a message) for what it cannot run.
csim_status launch_exact64(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b, const TrialOut& out, cudaStream_t st);
csim_status launch_dense32(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const std::vector<double>& hb, const BatchDev& b,
                           const TrialOut& out, cudaStream_t st);
csim_status launch_table(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b, const TrialOut& out, cudaStream_t st);
csim_status launch_prefix(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b, const TrialOut& out, cudaStream_t st);
bool prefix_supported(const csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, int R);
csim_status launch_full32(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b, const TrialOut& out, cudaStream_t st);
