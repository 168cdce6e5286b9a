// cabi.cu — implementation of include/conclave_b200.h (the C ABI) on top of the kernels.
//
// Host-side responsibilities: validate arguments, derive the per-set constants exactly as the
// reference derives them in fp64 (beta schedule model.py:562-580, vote threshold
// simulate.py:175), size grids as multiples of the SM count, and move buffers when the caller
// hands host memory.  There is no CPU implementation of the hot path in this file.
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
#include "exact64.cuh"
#include "dense32.cuh"
#include "prob32.cuh"
#include "table.cuh"
#include "prefix.cuh"
#include "full32.cuh"

static thread_local std::string g_err;

#ifndef CSIM_GRID_TILE
#define CSIM_GRID_TILE 65535           // gridDim.y / .z limit: the table-build launches are tiled over (set, round)
#endif

static csim_status fail(csim_status st, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return st;
}

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
    DevBuf sets, betas, nbeta, W, tab, cdf64, cdfk64, cdf32, cdfk32, ptab;
    double x_mid = 0.0, x_half = 0.0;   // centre and half-range of ideology_score (factorised exp)
    // host-buffer mode staging
    DevBuf o_winner, o_rounds, o_reason, o_final, o_tallies, o_whist, o_rhist, o_totals, i_tape;
    // prob-matrix staging
    DevBuf p_prev, p_cnt, p_fat, p_shares, p_bw, p_out, p_cand, p_pref;
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
static cudaError_t ctx_wait_prev(csim_ctx* c, cudaStream_t st) {
    return c->busy_valid ? cudaStreamWaitEvent(st, c->ev_busy, 0) : cudaSuccess;
}
static cudaError_t ctx_mark_busy(csim_ctx* c, cudaStream_t st) {
    cudaError_t e = cudaEventRecord(c->ev_busy, st);
    if (e == cudaSuccess) c->busy_valid = true;
    return e;
}

extern "C" int csim_abi_version(void) { return CSIM_ABI_VERSION; }
extern "C" const char* csim_last_error(void) { return g_err.c_str(); }

extern "C" csim_status csim_device_count(int32_t* n_out) {
    if (!n_out) return fail(CSIM_ERR_INVALID, "csim_device_count: n is NULL");
    *n_out = 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(CSIM_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    *n_out = n;
    return CSIM_OK;
}

extern "C" csim_status csim_create(csim_ctx** out, int device) {
    if (!out) return fail(CSIM_ERR_INVALID, "csim_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(CSIM_ERR_CUDA, "csim_create: no usable CUDA device (%s); this engine has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(CSIM_ERR_INVALID, "csim_create: device %d out of range [0,%d)", device, n);
    DeviceGuard guard;
    CU_TRY(cudaSetDevice(device));
    csim_ctx* c = new csim_ctx();
    c->device = device;
    e = cudaGetDeviceProperties(&c->prop, device);
    if (e != cudaSuccess) { delete c; return fail(CSIM_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    if (c->prop.major < 10) {
        int major = c->prop.major, minor = c->prop.minor;
        delete c;
        return fail(CSIM_ERR_UNSUPPORTED, "csim_create: device is sm_%d%d; this library is built for sm_100a only", major, minor);
    }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    cudaEventCreateWithFlags(&c->ev_busy, cudaEventDisableTiming);
    *out = c;
    return CSIM_OK;
}

extern "C" csim_status csim_destroy(csim_ctx* c) {
    if (!c) return CSIM_OK;
    DeviceGuard guard;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();             // nothing of this context may still be in flight when its buffers go
    for (DevBuf* b : {&c->x64, &c->x32, &c->region, &c->pap, &c->sets, &c->betas, &c->nbeta, &c->W, &c->tab, &c->cdf64, &c->cdfk64, &c->cdf32, &c->cdfk32, &c->ptab, &c->o_winner, &c->o_rounds,
                      &c->o_reason, &c->o_final, &c->o_tallies, &c->o_whist, &c->o_rhist, &c->o_totals, &c->i_tape,
                      &c->p_prev, &c->p_cnt, &c->p_fat, &c->p_shares, &c->p_bw, &c->p_out, &c->p_cand, &c->p_pref})
        b->release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_busy) cudaEventDestroy(c->ev_busy);
    delete c;
    return CSIM_OK;
}

extern "C" csim_status csim_device_info(csim_ctx* c, int* sm_count, int* cc_major, int* cc_minor, int* clock_khz) {
    if (!c) return fail(CSIM_ERR_INVALID, "csim_device_info: ctx is NULL");
    if (sm_count) *sm_count = c->prop.multiProcessorCount;
    if (cc_major) *cc_major = c->prop.major;
    if (cc_minor) *cc_minor = c->prop.minor;
    if (clock_khz) { int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device); *clock_khz = khz; }
    return CSIM_OK;
}

extern "C" csim_status csim_upload_roster(csim_ctx* c, int32_t E, const double* x, const int32_t* region, const uint8_t* pap) {
    if (!c) return fail(CSIM_ERR_INVALID, "csim_upload_roster: ctx is NULL");
    if (E <= 0 || !x || !region || !pap) return fail(CSIM_ERR_INVALID, "csim_upload_roster: need E > 0 and non-NULL columns (E=%d)", E);
    DeviceGuard guard;
    CU_TRY(cudaSetDevice(c->device));
    if (c->busy_valid) CU_TRY(cudaEventSynchronize(c->ev_busy));      // a run still in flight reads the old roster
    // dense region codes (equal code == same region), order of first appearance
    std::map<int32_t, int32_t> code;
    std::vector<int32_t> dense(E);
    for (int i = 0; i < E; ++i) {
        auto it = code.find(region[i]);
        if (it == code.end()) it = code.emplace(region[i], (int32_t)code.size()).first;
        dense[i] = it->second;
    }
    std::vector<float> xf(E);
    std::vector<uint8_t> pp(E);
    for (int i = 0; i < E; ++i) { xf[i] = (float)x[i]; pp[i] = pap[i] ? 1 : 0; }
    CU_TRY(c->x64.reserve(sizeof(double) * E));
    CU_TRY(c->x32.reserve(sizeof(float) * E));
    CU_TRY(c->region.reserve(sizeof(int32_t) * E));
    CU_TRY(c->pap.reserve(E));
    CU_TRY(cudaMemcpy(c->x64.p, x, sizeof(double) * E, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(c->x32.p, xf.data(), sizeof(float) * E, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(c->region.p, dense.data(), sizeof(int32_t) * E, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(c->pap.p, pp.data(), E, cudaMemcpyHostToDevice));
    c->E = E; c->Epad = (E + 31) & ~31; c->G = (int)code.size();
    double lo = x[0], hi = x[0];
    for (int i = 1; i < E; ++i) { lo = std::min(lo, x[i]); hi = std::max(hi, x[i]); }
    c->x_mid = 0.5 * (lo + hi); c->x_half = 0.5 * (hi - lo);
    if (!(c->x_mid == c->x_mid) || !(c->x_half == c->x_half)) { c->x_mid = 0.0; c->x_half = 1e300; }
    return CSIM_OK;
}

// model.py:562-580
static double beta_step(const csim_params* p, double beta, int r) {
    if (!p->enable_dynamic_beta || r <= 0) return beta;
    if (p->beta_mode == CSIM_BETA_STEPPED) {
        const int iv = p->beta_increment_interval_rounds;
        if (iv > 0 && r % iv == 0) beta = beta + p->beta_increment_amount;
    } else if (p->beta_growth_rate > 1.0) {
        beta = beta * p->beta_growth_rate;
    }
    return beta;
}

extern "C" csim_status csim_beta_schedule(const csim_params* p, int32_t first_round, int32_t n_rounds, double* out) {
    if (!p || !out || n_rounds < 0 || first_round < 1) return fail(CSIM_ERR_INVALID, "csim_beta_schedule: bad arguments");
    double b = p->initial_beta_weight;
    for (int r = 1; r < first_round + n_rounds; ++r) {
        b = beta_step(p, b, r);
        if (r >= first_round) out[r - first_round] = b;
    }
    return CSIM_OK;
}

extern "C" int32_t csim_need_votes(double thr, int32_t E) {
    const double t = thr * (double)E;               // simulate.py:175
    if (!(t == t)) return INT32_MAX;                // NaN threshold: never reached
    if (t <= 0.0) return 0;
    if (t > 2.0e9) return INT32_MAX;
    int32_t v = (int32_t)floor(t);
    while ((double)v < t) ++v;
    while (v > 0 && (double)(v - 1) >= t) --v;
    return v;
}

// smallest vote count v in [0, E + 1] with (double)v / E >= share (strict == 0) or > share (strict != 0): v / E is monotone
// in v, so "share of the votes < threshold" (simulate.py:111) is exactly "votes < csim_vote_threshold(threshold, E, 0)" and
// "share > threshold" (model.py:397) is exactly "votes >= csim_vote_threshold(threshold, E, 1)" — no division per candidate
extern "C" int32_t csim_vote_threshold(double share, int32_t E, int32_t strict) {
    if (E <= 0) return 0;
    int32_t v = 0;
    if (strict) { while (v <= E && !((double)v / (double)E > share)) ++v; }
    else        { while (v <= E && (double)v / (double)E < share) ++v; }
    return v;
}

static RosterDev roster_dev(const csim_ctx* c) {
    RosterDev r;
    r.E = c->E; r.Epad = c->Epad; r.G = c->G;
    r.x64 = (const double*)c->x64.p; r.x32 = (const float*)c->x32.p;
    r.region = (const int32_t*)c->region.p; r.pap = (const uint8_t*)c->pap.p;
    return r;
}

static DevSet make_set(const csim_params* p, int E, int wiring) {
    DevSet s;
    memset(&s, 0, sizeof s);
    const bool model = wiring == CSIM_WIRING_MODEL;
    s.st1p = 1 + p->stickiness_factor;
    s.bandwagon = p->bandwagon_strength;
    s.bonus = p->regional_affinity_bonus;
    s.pwf = p->papabile_weight_factor;
    s.fat_thr = p->fatigue_vote_share_threshold;
    s.fat_pen = p->fatigue_penalty_factor;
    s.stop_D = p->stop_candidate_threshold_unacceptable_distance;
    s.stop_T = p->stop_candidate_threat_min_vote_share;
    s.stop_boost = p->stop_candidate_boost_factor;
    s.f_st1p = (float)s.st1p; s.f_bandwagon = (float)s.bandwagon; s.f_bonus = (float)s.bonus; s.f_pwf = (float)s.pwf;
    s.f_fat_pen = (float)s.fat_pen; s.f_stop_D = (float)s.stop_D; s.f_stop_boost = (float)s.stop_boost;
    s.need = csim_need_votes(p->supermajority_threshold, E);
    s.max_rounds = p->max_rounds;
    s.rr = p->runoff_threshold_rounds;
    s.k = std::max(0, std::min(p->runoff_candidate_count, E));
    s.has_sticky = model && p->stickiness_factor > 0;
    s.has_bw = model && p->bandwagon_strength > 0;
    s.en_fat = model && p->enable_candidate_fatigue;
    s.fat_rounds = std::max(0, p->fatigue_threshold_rounds);
    s.fat_top_n = std::max(0, p->fatigue_top_n_immune);
    s.en_stop = model && p->enable_stop_candidate;
    // the two share tests of the model as integer vote thresholds: v / E in fp64 is monotone in v, so the comparisons the
    // reference makes on shares (simulate.py:111, model.py:397) are reproduced exactly without a division per candidate
    s.fat_cnt = csim_vote_threshold(s.fat_thr, E, 0);
    s.stop_cnt = csim_vote_threshold(s.stop_T, E, 1);
    return s;
}

// ---------------------------------------------------------------------------------------------
// csim_prob_matrix
// ---------------------------------------------------------------------------------------------
extern "C" csim_status csim_prob_matrix(csim_ctx* c, const csim_params* p, double beta, const int32_t* prev_vote,
                                        const int32_t* prev_count, const uint8_t* fatigued, const double* shares,
                                        const int32_t* cand_index, int32_t n_cand, int32_t precision, void* P_out) {
    if (!c || !p || !P_out) return fail(CSIM_ERR_INVALID, "csim_prob_matrix: NULL argument");
    if (c->E <= 0) return fail(CSIM_ERR_STATE, "csim_prob_matrix: roster not uploaded");
    if (precision != CSIM_FP32 && precision != CSIM_FP64) return fail(CSIM_ERR_INVALID, "csim_prob_matrix: bad precision %d", precision);
    DeviceGuard guard;
    CU_TRY(cudaSetDevice(c->device));
    if (c->busy_valid) CU_TRY(cudaEventSynchronize(c->ev_busy));      // shares sets / betas / W with csim_run (see header: one stream at a time)
    const int E = c->E;
    const int NC = cand_index ? n_cand : E;
    if (cand_index) {
        if (n_cand < 1) return fail(CSIM_ERR_INVALID, "csim_prob_matrix: n_cand must be >= 1 when cand_index is given");
        for (int i = 0; i < n_cand; ++i)
            if (cand_index[i] < 0 || cand_index[i] >= E) return fail(CSIM_ERR_INVALID, "csim_prob_matrix: cand_index[%d] = %d out of range", i, cand_index[i]);
    }
    const int32_t* d_cand = nullptr;
    if (cand_index) { CU_TRY(c->p_cand.reserve(4 * (size_t)n_cand)); CU_TRY(cudaMemcpy(c->p_cand.p, cand_index, 4 * (size_t)n_cand, cudaMemcpyHostToDevice)); d_cand = (const int32_t*)c->p_cand.p; }
    DevSet S = make_set(p, E, CSIM_WIRING_MODEL);   // the standalone model honours every term it is given
    const int32_t* d_prev = nullptr; const int32_t* d_cnt = nullptr; const uint8_t* d_fat = nullptr; const double* d_sh = nullptr;
    if (prev_vote) { CU_TRY(c->p_prev.reserve(4 * E)); CU_TRY(cudaMemcpy(c->p_prev.p, prev_vote, 4 * E, cudaMemcpyHostToDevice)); d_prev = (const int32_t*)c->p_prev.p; }
    if (prev_count) { CU_TRY(c->p_cnt.reserve(4 * E)); CU_TRY(cudaMemcpy(c->p_cnt.p, prev_count, 4 * E, cudaMemcpyHostToDevice)); d_cnt = (const int32_t*)c->p_cnt.p; }
    if (fatigued) { CU_TRY(c->p_fat.reserve(E)); CU_TRY(cudaMemcpy(c->p_fat.p, fatigued, E, cudaMemcpyHostToDevice)); d_fat = (const uint8_t*)c->p_fat.p; }
    if (shares) { CU_TRY(c->p_shares.reserve(8 * E)); CU_TRY(cudaMemcpy(c->p_shares.p, shares, 8 * E, cudaMemcpyHostToDevice)); d_sh = (const double*)c->p_shares.p; }
    RosterDev ro = roster_dev(c);
    if (precision == CSIM_FP64) {
        CU_TRY(c->sets.reserve(sizeof(DevSet)));
        CU_TRY(c->betas.reserve(sizeof(double)));
        CU_TRY(cudaMemcpy(c->sets.p, &S, sizeof S, cudaMemcpyHostToDevice));
        CU_TRY(cudaMemcpy(c->betas.p, &beta, sizeof beta, cudaMemcpyHostToDevice));
        CU_TRY(c->W.reserve(sizeof(double) * (size_t)E * c->Epad));
        CU_TRY(c->p_bw.reserve(8 * E));
        CU_TRY(c->p_out.reserve(sizeof(double) * (size_t)E * NC));
        dim3 g((c->Epad + 127) / 128, E, 1);
        k_base_scores64<<<g, 128>>>(ro, (const DevSet*)c->sets.p, (const double*)c->betas.p, 1, (double*)c->W.p, 0);
        ++c->launches;
        ProbArgs a;
        a.ro = ro; a.S = S; a.W = (const double*)c->W.p; a.prev_vote = d_prev; a.prev_count = d_cnt; a.fatigued = d_fat;
        a.shares = d_sh; a.cand = d_cand; a.ncand = NC; a.bw = (double*)c->p_bw.p; a.P = (double*)c->p_out.p;
        int use_bw = 0;
        if (d_cnt && S.has_bw) {
            int mx = 0;
            for (int i = 0; i < E; ++i) mx = std::max(mx, prev_count[i]);
            if (mx > 0) { use_bw = 1; k_prob_bw64<<<1, 256>>>(a); ++c->launches; }
        }
        k_prob_matrix64<<<(E + 63) / 64, 64>>>(a, use_bw);
        ++c->launches;
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpy(P_out, c->p_out.p, sizeof(double) * (size_t)E * NC, cudaMemcpyDeviceToHost));
    } else {
        CU_TRY(c->p_out.reserve(sizeof(float) * (size_t)E * NC));
        Prob32Args a;
        a.ro = ro; a.S = S; a.nb = (float)(-beta * 1.4426950408889634);
        a.prev_vote = d_prev; a.prev_count = d_cnt; a.fatigued = d_fat; a.shares = d_sh; a.cand = d_cand; a.ncand = NC; a.P = (float*)c->p_out.p;
        k_prob_matrix32<<<(E + 3) / 4, 128>>>(a);
        ++c->launches;
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpy(P_out, c->p_out.p, sizeof(float) * (size_t)E * NC, cudaMemcpyDeviceToHost));
    }
    return CSIM_OK;
}

// ---------------------------------------------------------------------------------------------
// csim_run
// ---------------------------------------------------------------------------------------------
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

static csim_status launch_exact64(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b,
                                  const TrialOut& out, cudaStream_t st) {
    const int E = c->E, R = b.R;
    const size_t wbytes = sizeof(double) * (size_t)b.n_sets * R * E * c->Epad;
    if (wbytes > ((size_t)96 << 30)) return fail(CSIM_ERR_UNSUPPORTED, "fp64 base-score tables need %zu bytes", wbytes);
    CU_TRY(c->W.reserve(wbytes));
    RosterDev ro = roster_dev(c);
    for (int sr0 = 0; sr0 < b.n_sets * R; sr0 += CSIM_GRID_TILE) {       // tiled over the 65535 limit of gridDim.z
        dim3 g((c->Epad + 127) / 128, E, std::min(CSIM_GRID_TILE, b.n_sets * R - sr0));
        k_base_scores64<<<g, 128, 0, st>>>(ro, b.sets, (const double*)c->betas.p, R, (double*)c->W.p, sr0);
        ++c->launches;
    }
    Exact64Args a;
    a.ro = ro; a.b = b; a.out = out; a.W = (const double*)c->W.p; a.wiring = ra->wiring;
    int Fmax = 1, kmax = 1;
    for (const DevSet& s : hs) { Fmax = std::max(Fmax, s.fat_rounds); kmax = std::max(kmax, s.k); }
    a.Fmax = Fmax; a.kmax = kmax;
    a.warp_smem = exact64_warp_smem(E, Fmax, kmax);
    int wpc = (int)std::min<size_t>(8, (200 * 1024) / a.warp_smem);
    if (wpc < 1) return fail(CSIM_ERR_UNSUPPORTED, "roster of %d electors needs %zu B of shared memory per trial", E, a.warp_smem);
    a.warps_per_cta = wpc;
    const size_t smem = a.warp_smem * wpc;
    int grid = 1;
    csim_status s = persistent_grid(c, k_trials_exact64, wpc * 32, smem, (uint64_t)b.n_sets * b.n_sims, wpc, &grid);
    if (s != CSIM_OK) return s;
    CU_TRY(cudaEventRecord(c->ev0, st));
    k_trials_exact64<<<grid, wpc * 32, smem, st>>>(a);
    CU_TRY(cudaEventRecord(c->ev1, st));
    c->ev_valid = true;
    ++c->launches;
    CU_TRY(cudaGetLastError());
    return CSIM_OK;
}

template <int LC, bool MODEL, bool FACT>
static csim_status launch_dense32_t(csim_ctx* c, Dense32Args& a, cudaStream_t st) {
    a.cta_smem = dense32_cta_smem(a.nch, a.L, a.ro.G);
    a.warp_smem = dense32_warp_smem(a.ro.E, a.nch, a.L, a.kmax, MODEL);
    int wpc = 8;
    const size_t budget = (size_t)(220 * 1024) / CSIM_MINB;                        // CSIM_MINB CTAs per SM
    while (wpc > 1 && a.cta_smem + a.warp_smem * wpc > budget) --wpc;
    if (a.cta_smem + a.warp_smem * wpc > 220 * 1024)
        return fail(CSIM_ERR_UNSUPPORTED, "roster of %d electors / %d regions needs %zu B of shared memory", a.ro.E, a.ro.G, a.cta_smem + a.warp_smem);
    a.warps_per_cta = wpc;
    const size_t smem = a.cta_smem + a.warp_smem * wpc;
    int grid = 1;
    void (*kern)(Dense32Args) = k_trials_dense32<LC, MODEL, FACT, false>;
    if (c->probe) {
        kern = k_trials_dense32<LC, MODEL, FACT, true>;
        a.probe_round = c->probe->round; a.probe_prev_vote = c->probe->prev_vote; a.probe_prev_count = c->probe->prev_count;
        a.probe_scores = c->probe->scores; a.probe_prefix = c->probe->prefix; a.probe_cap = c->probe->cap;
        c->probe->nch = a.nch; c->probe->chunk = a.L;
    }
    csim_status s = persistent_grid(c, kern, wpc * 32, smem, (uint64_t)a.b.n_sets * a.b.n_sims, wpc, &grid);
    if (s != CSIM_OK) return s;
    if (FACT && a.b.R > 0) {
        const int CP = a.nch * a.L;
        const size_t tb = sizeof(float) * (size_t)a.b.n_sets * a.b.R * 4 * CP;
        CU_TRY(c->tab.reserve(tb));
        a.tab = (const float*)c->tab.p;
        for (int sr0 = 0; sr0 < a.b.n_sets * a.b.R; sr0 += CSIM_GRID_TILE) {
            dim3 g((CP + 127) / 128, std::min(CSIM_GRID_TILE, a.b.n_sets * a.b.R - sr0));
            k_round_tables32<<<g, 128, 0, st>>>(a.ro, a.b.sets, (const double*)c->betas.p, a.b.R, CP, c->x_mid, (float*)c->tab.p, sr0);
            ++c->launches;
        }
    }
    CU_TRY(cudaEventRecord(c->ev0, st));
    kern<<<grid, wpc * 32, smem, st>>>(a);
    CU_TRY(cudaEventRecord(c->ev1, st));
    c->ev_valid = true;
    ++c->launches;
    CU_TRY(cudaGetLastError());
    return CSIM_OK;
}

template <int LC>
static csim_status launch_dense32_d(csim_ctx* c, Dense32Args& a, bool model, bool fact, cudaStream_t st) {
    if (model) return fact ? launch_dense32_t<LC, true, true>(c, a, st) : launch_dense32_t<LC, true, false>(c, a, st);
    return fact ? launch_dense32_t<LC, false, true>(c, a, st) : launch_dense32_t<LC, false, false>(c, a, st);
}

static csim_status launch_dense32(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const std::vector<double>& hb,
                                  const BatchDev& b, const TrialOut& out, cudaStream_t st) {
    for (const DevSet& s : hs)
        if (s.en_fat || s.en_stop)
            return fail(CSIM_ERR_UNSUPPORTED, "candidate fatigue / stop-candidate are implemented in CSIM_FP64 precision only");
    for (int i = 0; i < ra->n_param_sets; ++i)
        if (ra->params[i].regional_affinity_bonus < 0 || ra->params[i].bandwagon_strength < 0 || ra->params[i].papabile_weight_factor < 0)
            return fail(CSIM_ERR_UNSUPPORTED, "negative bonus / strength / papabile factor: use CSIM_FP64 precision");
    if (c->G > 64) return fail(CSIM_ERR_UNSUPPORTED, "%d distinct regions (> 64): use CSIM_FP64 precision", c->G);
    Dense32Args a;
    a.ro = roster_dev(c); a.b = b; a.out = out; a.nbeta = (const float*)c->nbeta.p; a.wiring = ra->wiring; a.tab = nullptr;
    a.probe_round = 0; a.probe_prev_vote = nullptr; a.probe_prev_count = nullptr; a.probe_scores = nullptr; a.probe_prefix = nullptr; a.probe_cap = 0;
    int kmax = 1;
    for (const DevSet& s : hs) kmax = std::max(kmax, s.k);
    a.kmax = kmax;
    // factorised exponential: needs beta >= 0 and e^{+-beta x'} comfortably inside fp32 range
    bool fact = getenv("CSIM_DENSE_MUFU") == nullptr;
    for (double be : hb) if (!(be >= 0.0) || !(be * c->x_half <= 26.0)) fact = false;   // e^{3 beta x'} must stay finite in fp32
    const bool model = ra->wiring == CSIM_WIRING_MODEL;
    const int E = c->E;
    if (E <= 256) {                       // chunks of 8 candidates, at most 32 chunks
        a.L = 8; a.nch = (E + 7) / 8;
        return launch_dense32_d<8>(c, a, model, fact, st);
    }
    a.L = (((E + 31) / 32) + 3) & ~3;     // at most 32 chunks of L candidates
    a.nch = (E + a.L - 1) / a.L;
    return launch_dense32_d<0>(c, a, model, fact, st);
}

template <class Real>
static csim_status launch_table_t(csim_ctx* c, const BatchDev& b, const TrialOut& out, int kmax, const Real* cdf, const Real* cdfk,
                                  cudaStream_t st) {
    TableArgs<Real> a;
    a.ro = roster_dev(c); a.b = b; a.out = out; a.cdf = cdf; a.cdfk = cdfk; a.kmax = kmax;
    const int E = c->E;
    const size_t Ei = ((size_t)E + 3) & ~(size_t)3;
    a.warp_smem = (Ei * 4 + (((size_t)kmax + 3) & ~(size_t)3) * 4 + Ei + 15) & ~(size_t)15;
    const size_t qtab = (((size_t)b.R * E * sizeof(Real)) + 15) & ~(size_t)15;
    const size_t full = (cdf_round_stride(E) * sizeof(Real) + 15) & ~(size_t)15;
    int wpc = CSIM_TABLE_WPC;
    a.stage_q = (qtab + a.warp_smem * wpc <= 64 * 1024) ? 1 : 0;
    a.stage = (a.stage_q && full + qtab + a.warp_smem * wpc <= 200 * 1024) ? 1 : 0;
    a.cta_smem = (a.stage ? full : 0) + (a.stage_q ? qtab : 0);
    while (wpc > 1 && a.cta_smem + a.warp_smem * wpc > 220 * 1024) --wpc;
    if (a.cta_smem + a.warp_smem * wpc > 220 * 1024)
        return fail(CSIM_ERR_UNSUPPORTED, "table engine: a roster of %d electors does not fit in shared memory", E);
    a.warps_per_cta = wpc;
    const size_t smem = a.cta_smem + a.warp_smem * wpc;
    void (*kern)(TableArgs<Real>) = a.stage ? k_trials_table<Real, true> : k_trials_table<Real, false>;
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpc * 32, smem));
    if (per_sm < 1) return fail(CSIM_ERR_UNSUPPORTED, "table kernel does not fit on an SM (smem=%zu)", smem);
    const uint64_t sb_trials = (uint64_t)wpc * 32;
    const uint64_t n_sb = ((b.n_sims + sb_trials - 1) / sb_trials) * b.n_sets;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_sb, (uint64_t)c->prop.multiProcessorCount * per_sm));
    CU_TRY(cudaEventRecord(c->ev0, st));
    kern<<<grid, wpc * 32, smem, st>>>(a);
    CU_TRY(cudaEventRecord(c->ev1, st));
    c->ev_valid = true;
    ++c->launches;
    CU_TRY(cudaGetLastError());
    return CSIM_OK;
}

static csim_status launch_table(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b,
                                const TrialOut& out, cudaStream_t st) {
    if (ra->wiring != CSIM_WIRING_REF_CLI)
        return fail(CSIM_ERR_UNSUPPORTED, "the table engine exists for the ref_cli wiring only (under `model` the matrix depends on the trial)");
    const int E = c->E, R = b.R;
    int kmax = 1;
    for (const DevSet& s : hs) kmax = std::max(kmax, s.k);
    const size_t rows = (size_t)b.n_sets * R * E;
    const size_t cdf_elems = (size_t)b.n_sets * R * cdf_round_stride(E);
    const size_t wbytes = sizeof(double) * (size_t)b.n_sets * R * E * c->Epad;
    if (wbytes + rows * E * 12 > ((size_t)120 << 30)) return fail(CSIM_ERR_UNSUPPORTED, "CDF tables need more than 120 GB");
    CU_TRY(c->W.reserve(wbytes));
    CU_TRY(c->cdf64.reserve(sizeof(double) * cdf_elems));
    CU_TRY(c->cdfk64.reserve(sizeof(double) * rows * kmax));
    RosterDev ro = roster_dev(c);
    for (int sr0 = 0; sr0 < b.n_sets * R; sr0 += CSIM_GRID_TILE) {       // tiled over the 65535 limit of gridDim.y / .z
        const int nt = std::min(CSIM_GRID_TILE, b.n_sets * R - sr0);
        dim3 g((c->Epad + 127) / 128, E, nt);
        k_base_scores64<<<g, 128, 0, st>>>(ro, b.sets, (const double*)c->betas.p, R, (double*)c->W.p, sr0);
        dim3 g2((E + 63) / 64, nt);
        k_cdf_tables64<<<g2, 64, 0, st>>>(ro, b.sets, (const double*)c->W.p, R, kmax, (double*)c->cdf64.p, (double*)c->cdfk64.p, sr0);
        c->launches += 2;
    }
    if (ra->precision == CSIM_FP64)
        return launch_table_t<double>(c, b, out, kmax, (const double*)c->cdf64.p, (const double*)c->cdfk64.p, st);
    CU_TRY(c->cdf32.reserve(sizeof(float) * cdf_elems));
    CU_TRY(c->cdfk32.reserve(sizeof(float) * rows * kmax));
    if (R > 0) {
        k_f64_to_f32<<<c->prop.multiProcessorCount * 4, 256, 0, st>>>((const double*)c->cdf64.p, (float*)c->cdf32.p, cdf_elems);
        k_f64_to_f32<<<c->prop.multiProcessorCount, 256, 0, st>>>((const double*)c->cdfk64.p, (float*)c->cdfk32.p, rows * kmax);
        c->launches += 2;
    }
    return launch_table_t<float>(c, b, out, kmax, (const float*)c->cdf32.p, (const float*)c->cdfk32.p, st);
}

// ---------------------------------------------------------------------------------------------
// PREFIX engine (prefix.cuh)
// ---------------------------------------------------------------------------------------------
template <bool LC8, bool MODEL, bool STAGE>
static csim_status launch_prefix_t(csim_ctx* c, PrefixArgs& a, cudaStream_t st) {
    void (*kern)(const PrefixArgs) = k_trials_prefix<LC8, MODEL, STAGE, false>;
    if (c->probe) {
        if (STAGE) return fail(CSIM_ERR_STATE, "prefix probe: tables must be read through global memory");
        kern = k_trials_prefix<LC8, MODEL, false, true>;
        a.probe_round = c->probe->round; a.probe_prev_vote = c->probe->prev_vote; a.probe_prev_count = c->probe->prev_count;
        a.probe_scores = c->probe->scores; a.probe_prefix = c->probe->prefix; a.probe_cap = c->probe->cap;
        c->probe->nch = a.nch; c->probe->chunk = 1 << a.Ls;
    }
    const int wpc = a.warps_per_cta;
    const size_t smem = a.cta_smem + a.warp_smem * wpc;
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpc * 32, smem));
    if (per_sm < 1) return fail(CSIM_ERR_UNSUPPORTED, "prefix kernel does not fit on an SM (threads=%d, smem=%zu)", wpc * 32, smem);
    const uint64_t ctas = (uint64_t)c->prop.multiProcessorCount * per_sm;      // persistent: a multiple of the SM count
    // work item = warps_per_cta x tpi trials of one set; ~8 items per CTA keeps the tail short while a CTA
    // that stays inside one parameter set never re-stages its tables
    uint64_t tpi = a.b.n_sims / ((uint64_t)wpc * ctas * 8);
    tpi = std::max<uint64_t>(1, std::min<uint64_t>(tpi, 64));
    a.tpi = (int)tpi;
    const uint64_t per_item = (uint64_t)wpc * tpi;
    const uint64_t n_items = ((a.b.n_sims + per_item - 1) / per_item) * a.b.n_sets;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_items, ctas));
    CU_TRY(cudaEventRecord(c->ev0, st));
    kern<<<grid, wpc * 32, smem, st>>>(a);
    CU_TRY(cudaEventRecord(c->ev1, st));
    c->ev_valid = true;
    ++c->launches;
    CU_TRY(cudaGetLastError());
    return CSIM_OK;
}

// preconditions of the fp32 PREFIX engine (launch_prefix reports the same conditions as errors)
static bool prefix_supported(const csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, int R) {
    for (const DevSet& s : hs) if (s.en_fat || s.en_stop) return false;
    for (int i = 0; i < ra->n_param_sets; ++i)
        if (ra->params[i].regional_affinity_bonus < 0 || ra->params[i].bandwagon_strength < 0 || ra->params[i].papabile_weight_factor < 0) return false;
    if (c->E >= 32768) return false;
    int kmax = 1;
    for (const DevSet& s : hs) kmax = std::max(kmax, s.k);
    PrefixArgs probe;                                        // does it fit at all? (many regions make the bonus table large)
    memset(&probe, 0, sizeof probe);
    return prefix_plan(probe, c->E, c->G, R, std::max(R, 1), kmax, ra->wiring == CSIM_WIRING_MODEL,
                       (size_t)c->prop.sharedMemPerBlockOptin - 64, true);
}

static csim_status launch_prefix(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b,
                                 const TrialOut& out, cudaStream_t st) {
    for (const DevSet& s : hs)
        if (s.en_fat || s.en_stop)
            return fail(CSIM_ERR_UNSUPPORTED, "candidate fatigue / stop-candidate are implemented in CSIM_FP64 precision only");
    for (int i = 0; i < ra->n_param_sets; ++i)
        if (ra->params[i].regional_affinity_bonus < 0 || ra->params[i].bandwagon_strength < 0 || ra->params[i].papabile_weight_factor < 0)
            return fail(CSIM_ERR_UNSUPPORTED, "negative bonus / strength / papabile factor: use CSIM_FP64 precision");
    if (ra->precision != CSIM_FP32) return fail(CSIM_ERR_UNSUPPORTED, "the prefix engine is an fp32 engine (use DENSE or TABLE with CSIM_FP64)");
    const int E = c->E, R = b.R;
    if (E >= 32768) return fail(CSIM_ERR_UNSUPPORTED, "prefix engine: at most 32767 electors");
    const bool model = ra->wiring == CSIM_WIRING_MODEL;
    PrefixArgs a;
    memset(&a, 0, sizeof a);
    a.ro = roster_dev(c); a.b = b; a.out = out; a.nbeta = (const float*)c->nbeta.p;
    int kmax = 1;
    for (const DevSet& s : hs) kmax = std::max(kmax, s.k);
    // only rounds that can be full-field need tables: once a run-off exists (k > 0) every round after
    // max(runoff_threshold_rounds, 1) is a run-off round (simulate.py:183-186)
    int Rt = 1;
    for (const DevSet& s : hs) Rt = std::max(Rt, s.k > 0 ? std::min(R, std::max(s.rr, 1)) : R);
    Rt = std::max(1, std::min(Rt, std::max(R, 1)));
    const bool stage_ok = getenv("CSIM_PREFIX_NOSTAGE") == nullptr && !c->probe;    // a probe reads the same tables through global memory
    const char* ls_env = getenv("CSIM_PREFIX_LS");                  // tuning: force the chunk size (log2)
    if (!prefix_plan(a, E, c->G, R, Rt, kmax, model, (size_t)c->prop.sharedMemPerBlockOptin - 64 /* static: the staging mbarrier */, stage_ok,
                     ls_env ? atoi(ls_env) : 0))
        return fail(CSIM_ERR_UNSUPPORTED, "prefix engine: a roster of %d electors does not fit in shared memory", E);
    const bool stage = a.tab_smem != 0;
    const int Ls = a.Ls;
    const size_t tbytes = a.set_pitch * 4 * (size_t)b.n_sets;
    if (tbytes > ((size_t)64 << 30)) return fail(CSIM_ERR_UNSUPPORTED, "prefix tables need %zu bytes", tbytes);
    CU_TRY(c->ptab.reserve(tbytes));
    a.T = (const float*)c->ptab.p;
    for (int sr0 = 0; R > 0 && sr0 < b.n_sets * Rt; sr0 += CSIM_GRID_TILE) {
        dim3 g((E + 7) / 8, std::min(CSIM_GRID_TILE, b.n_sets * Rt - sr0));
        k_prefix_tables32<<<g, 256, 0, st>>>(a.ro, b.sets, (const double*)c->betas.p, R, Rt, a.Ls, a.nch, a.stride, a.rows, a.e_quad_end,
                                             a.set_pitch, a.col_major, (float*)c->ptab.p, sr0);
        ++c->launches;
    }
    const bool lc8 = Ls == 3;
    if (model) {
        if (stage) return lc8 ? launch_prefix_t<true, true, true>(c, a, st) : launch_prefix_t<false, true, true>(c, a, st);
        return lc8 ? launch_prefix_t<true, true, false>(c, a, st) : launch_prefix_t<false, true, false>(c, a, st);
    }
    if (stage) return lc8 ? launch_prefix_t<true, false, true>(c, a, st) : launch_prefix_t<false, false, true>(c, a, st);
    return lc8 ? launch_prefix_t<true, false, false>(c, a, st) : launch_prefix_t<false, false, false>(c, a, st);
}

// ---------------------------------------------------------------------------------------------
// FULL engine (full32.cuh): fp32, every model term inline, fatigue + stop-candidate included
// ---------------------------------------------------------------------------------------------
static csim_status launch_full32(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b,
                                 const TrialOut& out, cudaStream_t st) {
    if (ra->precision != CSIM_FP32) return fail(CSIM_ERR_UNSUPPORTED, "the full engine is an fp32 engine (CSIM_FP64: use DENSE)");
    const int E = c->E, R = b.R;
    if (E >= 32768) return fail(CSIM_ERR_UNSUPPORTED, "full engine: at most 32767 electors");
    Full32Args a;
    memset(&a, 0, sizeof a);
    int kmax = 1, Fmax = 1;
    for (const DevSet& s : hs) { kmax = std::max(kmax, s.k); Fmax = std::max(Fmax, s.fat_rounds); }
    if (!full32_plan(a, E, c->G, R, kmax, Fmax, (size_t)c->prop.sharedMemPerBlockOptin))
        return fail(CSIM_ERR_UNSUPPORTED, "full engine: a roster of %d electors (fatigue window %d) does not fit in shared memory", E, Fmax);
    a.ro = roster_dev(c); a.b = b; a.out = out; a.nbeta = (const float*)c->nbeta.p;
    const int wpc = a.warps_per_cta;
    const size_t smem = a.cta_smem + a.warp_smem * wpc;
    bool ext = false, neg = false;                            // which terms the batch needs decides the instantiation
    for (const DevSet& s : hs) ext = ext || s.en_fat || s.en_stop;
    for (int i = 0; i < ra->n_param_sets; ++i) {
        const csim_params& q = ra->params[i];
        neg = neg || q.regional_affinity_bonus < 0 || q.bandwagon_strength < 0 || q.papabile_weight_factor < 0 || q.stickiness_factor < -1 ||
              q.fatigue_penalty_factor < 0 || q.stop_candidate_boost_factor < 0;
    }
    void (*kern)(const Full32Args) = ext ? (neg ? k_trials_full32<true, true, false> : k_trials_full32<true, false, false>)
                                         : (neg ? k_trials_full32<false, true, false> : k_trials_full32<false, false, false>);
    if (c->probe) {
        kern = ext ? (neg ? k_trials_full32<true, true, true> : k_trials_full32<true, false, true>)
                   : (neg ? k_trials_full32<false, true, true> : k_trials_full32<false, false, true>);
        a.probe_round = c->probe->round; a.probe_prev_vote = c->probe->prev_vote; a.probe_prev_count = c->probe->prev_count;
        a.probe_fatigued = c->probe->fatigued;
        a.probe_scores = c->probe->scores; a.probe_prefix = c->probe->prefix; a.probe_cap = c->probe->cap;
        c->probe->nch = a.nch; c->probe->chunk = 8 * a.Lb;
    }
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpc * 32, smem));
    if (per_sm < 1) return fail(CSIM_ERR_UNSUPPORTED, "full kernel does not fit on an SM (threads=%d, smem=%zu)", wpc * 32, smem);
    const uint64_t ctas = (uint64_t)c->prop.multiProcessorCount * per_sm;
    uint64_t tpi = a.b.n_sims / ((uint64_t)wpc * ctas * 8);
    tpi = std::max<uint64_t>(1, std::min<uint64_t>(tpi, 64));
    a.tpi = (int)tpi;
    const uint64_t per_item = (uint64_t)wpc * tpi;
    const uint64_t n_items = ((a.b.n_sims + per_item - 1) / per_item) * a.b.n_sets;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_items, ctas));
    CU_TRY(cudaEventRecord(c->ev0, st));
    kern<<<grid, wpc * 32, smem, st>>>(a);
    CU_TRY(cudaEventRecord(c->ev1, st));
    c->ev_valid = true;
    ++c->launches;
    CU_TRY(cudaGetLastError());
    return CSIM_OK;
}

extern "C" csim_status csim_prefix_plan_query(int32_t n_electors, int32_t n_regions, int32_t max_rounds, int32_t table_rounds, int32_t kmax,
                                              int32_t wiring, uint64_t smem_budget, int32_t allow_stage, csim_prefix_plan* out) {
    if (!out) return fail(CSIM_ERR_INVALID, "csim_prefix_plan_query: out is NULL");
    if (n_electors < 1 || n_electors >= 32768 || n_regions < 1 || n_regions > n_electors || max_rounds < 0 || kmax < 0)
        return fail(CSIM_ERR_INVALID, "csim_prefix_plan_query: need 1 <= electors < 32768, 1 <= regions <= electors, rounds >= 0, kmax >= 0");
    PrefixArgs a;
    memset(&a, 0, sizeof a);
    memset(out, 0, sizeof *out);
    const bool model = wiring == CSIM_WIRING_MODEL;
    if (!prefix_plan(a, n_electors, n_regions, max_rounds, std::max(1, std::min(table_rounds, std::max(max_rounds, 1))), std::max(kmax, 1), model,
                     (size_t)smem_budget, allow_stage != 0))
        return fail(CSIM_ERR_UNSUPPORTED, "prefix engine: a roster of %d electors does not fit in %llu bytes of shared memory", n_electors,
                    (unsigned long long)smem_budget);
    out->staged = a.tab_smem != 0; out->chunk = 1 << a.Ls; out->n_chunks = a.nch; out->row_pitch = a.stride; out->rows = a.rows;
    out->table_rounds = a.Rt; out->warps_per_cta = a.warps_per_cta; out->e_quad_end = a.e_quad_end;
    out->table_bytes_per_set = a.set_pitch * 4; out->smem_tables = a.tab_smem; out->smem_cta = a.cta_smem; out->smem_per_warp = a.warp_smem;
    out->smem_total = a.cta_smem + a.warp_smem * (uint64_t)a.warps_per_cta;
    const uint32_t off[12] = {a.o_xs, a.o_pw, a.o_rb, a.o_xg, a.o_nbs, a.w_bwp, a.w_tally, a.w_eff, a.w_votes, a.w_mark, a.cp4, (uint32_t)a.step0_4};
    for (int i = 0; i < 12; ++i) out->offsets[i] = off[i];
    return CSIM_OK;
}

extern "C" int32_t csim_prefix_slot(int32_t elector, int32_t e_quad_end) { return prefix_slot(elector, e_quad_end); }

extern "C" csim_status csim_run(csim_ctx* c, const csim_run_args* ra) {
    if (!c || !ra) return fail(CSIM_ERR_INVALID, "csim_run: NULL argument");
    if (c->E <= 0) return fail(CSIM_ERR_STATE, "csim_run: roster not uploaded");
    if (ra->n_param_sets < 0) return fail(CSIM_ERR_INVALID, "csim_run: n_param_sets < 0");
    if ((uint64_t)ra->n_param_sets * ra->n_sims == 0) return CSIM_OK;      // an empty batch (e.g. a rank whose shard is empty) touches nothing
    if (!ra->params) return fail(CSIM_ERR_INVALID, "csim_run: params is NULL");
    if (!ra->winner || !ra->rounds) return fail(CSIM_ERR_INVALID, "csim_run: winner and rounds outputs are required");
    if (ra->wiring != CSIM_WIRING_REF_CLI && ra->wiring != CSIM_WIRING_MODEL) return fail(CSIM_ERR_INVALID, "csim_run: bad wiring %d", ra->wiring);
    if (ra->precision != CSIM_FP32 && ra->precision != CSIM_FP64) return fail(CSIM_ERR_INVALID, "csim_run: bad precision %d", ra->precision);
    if (ra->engine != CSIM_ENGINE_AUTO && ra->engine != CSIM_ENGINE_DENSE && ra->engine != CSIM_ENGINE_TABLE && ra->engine != CSIM_ENGINE_PREFIX && ra->engine != CSIM_ENGINE_FULL)
        return fail(CSIM_ERR_INVALID, "csim_run: bad engine %d", ra->engine);
    if (ra->uniform_tape && ra->precision != CSIM_FP64) return fail(CSIM_ERR_INVALID, "csim_run: a uniform tape requires CSIM_FP64 precision");
    const int R = ra->params[0].max_rounds;
    if (R < 0) return fail(CSIM_ERR_INVALID, "csim_run: max_rounds < 0");
    for (int i = 0; i < ra->n_param_sets; ++i)
        if (ra->params[i].max_rounds != R) return fail(CSIM_ERR_INVALID, "csim_run: every parameter set of a batch must share max_rounds");
    if ((uint64_t)ra->n_param_sets * (uint64_t)std::max(R, 1) > (1ull << 30)) return fail(CSIM_ERR_INVALID, "csim_run: n_param_sets * max_rounds too large");
    const uint64_t n = (uint64_t)ra->n_param_sets * ra->n_sims;
    DeviceGuard guard;
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)ra->stream;
    CU_TRY(ctx_wait_prev(c, st));          // the previous call's kernels still read the scratch this call overwrites
    const int E = c->E;
    const bool dev = ra->buffers_on_device != 0;

    // per-set constants
    std::vector<DevSet> hs(ra->n_param_sets);
    std::vector<double> hb((size_t)ra->n_param_sets * std::max(R, 1));
    std::vector<float> hnb(hb.size());
    for (int i = 0; i < ra->n_param_sets; ++i) {
        hs[i] = make_set(&ra->params[i], E, ra->wiring);
        if (R > 0) csim_beta_schedule(&ra->params[i], 1, R, &hb[(size_t)i * R]);
        for (int r = 0; r < R; ++r) hnb[(size_t)i * R + r] = (float)(-hb[(size_t)i * R + r] * 1.4426950408889634);
    }
    CU_TRY(c->sets.reserve(sizeof(DevSet) * hs.size()));
    CU_TRY(c->betas.reserve(sizeof(double) * hb.size()));
    CU_TRY(c->nbeta.reserve(sizeof(float) * hnb.size()));
    CU_TRY(cudaMemcpyAsync(c->sets.p, hs.data(), sizeof(DevSet) * hs.size(), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(c->betas.p, hb.data(), sizeof(double) * hb.size(), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(c->nbeta.p, hnb.data(), sizeof(float) * hnb.size(), cudaMemcpyHostToDevice, st));
    // the vectors above are pageable: the async copies have completed their host read on return

    BatchDev b;
    b.sets = (const DevSet*)c->sets.p; b.n_sets = ra->n_param_sets; b.R = R; b.n_sims = ra->n_sims;
    b.sim_id_begin = ra->sim_id_begin; b.seed_lo = (uint32_t)ra->seed; b.seed_hi = (uint32_t)(ra->seed >> 32);
    b.tape = nullptr;
    TrialOut out;
    memset(&out, 0, sizeof out);
    const size_t nh_w = (size_t)ra->n_param_sets * (E + 1), nh_r = (size_t)ra->n_param_sets * (R + 1), nh_t = (size_t)ra->n_param_sets * 2;
    if (dev) {
        b.tape = ra->uniform_tape;
        out.winner = ra->winner; out.rounds = ra->rounds; out.reason = ra->reason; out.final_votes = ra->final_votes;
        out.round_tallies = ra->round_tallies;
        out.winner_hist = (unsigned long long*)ra->winner_hist; out.rounds_hist = (unsigned long long*)ra->rounds_hist;
        out.totals = (unsigned long long*)ra->totals;
    } else {
        if (ra->uniform_tape) {
            const size_t tb = sizeof(double) * n * (size_t)R * E;
            CU_TRY(c->i_tape.reserve(tb));
            CU_TRY(cudaMemcpyAsync(c->i_tape.p, ra->uniform_tape, tb, cudaMemcpyHostToDevice, st));
            b.tape = (const double*)c->i_tape.p;
        }
        CU_TRY(c->o_winner.reserve(4 * n)); out.winner = (int32_t*)c->o_winner.p;
        CU_TRY(c->o_rounds.reserve(4 * n)); out.rounds = (int32_t*)c->o_rounds.p;
        if (ra->reason) { CU_TRY(c->o_reason.reserve(n)); out.reason = (uint8_t*)c->o_reason.p; }
        if (ra->final_votes) { CU_TRY(c->o_final.reserve(4 * n * E)); out.final_votes = (int32_t*)c->o_final.p; }
        if (ra->round_tallies) { CU_TRY(c->o_tallies.reserve(4 * n * (size_t)R * E)); out.round_tallies = (int32_t*)c->o_tallies.p; }
        if (ra->winner_hist) { CU_TRY(c->o_whist.reserve(8 * nh_w)); CU_TRY(cudaMemsetAsync(c->o_whist.p, 0, 8 * nh_w, st)); out.winner_hist = (unsigned long long*)c->o_whist.p; }
        if (ra->rounds_hist) { CU_TRY(c->o_rhist.reserve(8 * nh_r)); CU_TRY(cudaMemsetAsync(c->o_rhist.p, 0, 8 * nh_r, st)); out.rounds_hist = (unsigned long long*)c->o_rhist.p; }
        if (ra->totals) { CU_TRY(c->o_totals.reserve(8 * nh_t)); CU_TRY(cudaMemsetAsync(c->o_totals.p, 0, 8 * nh_t, st)); out.totals = (unsigned long long*)c->o_totals.p; }
    }

    // engine choice (CSIM_ENGINE_AUTO)
    //   model wiring, fp32: the PREFIX engine (trial-invariant chunk-prefix tables + per-trial corrections) whenever its
    //     preconditions hold; with candidate fatigue / stop-candidate live the FULL engine (every term inline); else DENSE
    //     (which reports what it cannot do);
    //   ref_cli wiring: the TABLE engine when the batch fills the GPU with its 512-trial super-batches (or fp64); rosters
    //     whose E x C rows do not fit in shared memory (E > 256) go to PREFIX, and so do small fp32 batches (BASELINE
    //     configs 1-2: 100 / 1000 trials finish in 0.06-0.08 ms on its one-warp-per-trial grid; DENSE if PREFIX cannot run).
    int engine = ra->engine;
    if (engine == CSIM_ENGINE_AUTO) {
        const bool pfx_ok = ra->precision == CSIM_FP32 && prefix_supported(c, ra, hs, R);
        bool extras = false;                                     // candidate fatigue / stop-candidate live (model wiring only)
        for (const DevSet& s : hs) extras = extras || s.en_fat || s.en_stop;
        if (ra->wiring == CSIM_WIRING_MODEL) {
            // (FULL takes everything PREFIX refuses: fatigue / stop-candidate, negative factors; DENSE in fp32 is explicit-only)
            engine = pfx_ok ? CSIM_ENGINE_PREFIX : (ra->precision == CSIM_FP32 ? CSIM_ENGINE_FULL : CSIM_ENGINE_DENSE);
        } else {
            const uint64_t job_sims = ra->auto_batch_sims ? ra->auto_batch_sims : ra->n_sims;      // the whole job's size when this is a shard
            const uint64_t n_sb = ((job_sims + (uint64_t)CSIM_TABLE_WPC * 32 - 1) / ((uint64_t)CSIM_TABLE_WPC * 32)) * ra->n_param_sets;
            if (ra->precision == CSIM_FP64) engine = CSIM_ENGINE_TABLE;
            else if (E > 256) engine = pfx_ok ? CSIM_ENGINE_PREFIX : CSIM_ENGINE_TABLE;
            else engine = n_sb >= (uint64_t)c->prop.multiProcessorCount / 2 ? CSIM_ENGINE_TABLE : (pfx_ok ? CSIM_ENGINE_PREFIX : CSIM_ENGINE_FULL);
        }
    }
    csim_status s = engine == CSIM_ENGINE_FULL ? launch_full32(c, ra, hs, b, out, st)
                    : engine == CSIM_ENGINE_PREFIX ? launch_prefix(c, ra, hs, b, out, st)
                    : engine == CSIM_ENGINE_TABLE ? launch_table(c, ra, hs, b, out, st)
                    : (ra->precision == CSIM_FP64 ? launch_exact64(c, ra, hs, b, out, st) : launch_dense32(c, ra, hs, hb, b, out, st));
    int ran_engine = engine, ran_precision = (engine == CSIM_ENGINE_DENSE || engine == CSIM_ENGINE_TABLE) ? ra->precision : CSIM_FP32;
    int fell = 0;
    if (s == CSIM_ERR_UNSUPPORTED && ra->engine == CSIM_ENGINE_AUTO) {
        // AUTO does not refuse a shape one of the engines can run.  What its first choice cannot hold or express goes, in fp32, to
        // the FULL engine (every term inline, no table-size limits), then to the fp32 dense engine (FULL with 1024 electors x 60
        // regions of bonus rows does not fit in shared memory), and what that cannot express either (fatigue / stop-candidate,
        // negative factors, > 64 regions) to the fp64 exact engine, whose tables live in global memory.  In that last case the
        // results are of CSIM_FP64 quality — and cost — whatever precision was asked for: csim_last_run_info reports it and the
        // Python host warns.  An explicitly named engine still reports UNSUPPORTED.
        fell = 1;
        if (ra->precision == CSIM_FP32 && engine != CSIM_ENGINE_FULL) {
            s = launch_full32(c, ra, hs, b, out, st); ran_engine = CSIM_ENGINE_FULL; ran_precision = CSIM_FP32;
        }
        if (s == CSIM_ERR_UNSUPPORTED && ra->precision == CSIM_FP32 && engine != CSIM_ENGINE_DENSE) {
            s = launch_dense32(c, ra, hs, hb, b, out, st); ran_engine = CSIM_ENGINE_DENSE; ran_precision = CSIM_FP32;
        }
        if (s == CSIM_ERR_UNSUPPORTED && !(engine == CSIM_ENGINE_DENSE && ra->precision == CSIM_FP64)) {
            s = launch_exact64(c, ra, hs, b, out, st); ran_engine = CSIM_ENGINE_DENSE; ran_precision = CSIM_FP64;
        }
    }
    if (s != CSIM_OK) return s;
    c->last_engine = ran_engine; c->last_precision = ran_precision; c->last_fell = fell;
    CU_TRY(ctx_mark_busy(c, st));

    if (!dev) {
        CU_TRY(cudaMemcpyAsync(ra->winner, out.winner, 4 * n, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(ra->rounds, out.rounds, 4 * n, cudaMemcpyDeviceToHost, st));
        if (ra->reason) CU_TRY(cudaMemcpyAsync(ra->reason, out.reason, n, cudaMemcpyDeviceToHost, st));
        if (ra->final_votes) CU_TRY(cudaMemcpyAsync(ra->final_votes, out.final_votes, 4 * n * E, cudaMemcpyDeviceToHost, st));
        if (ra->round_tallies) CU_TRY(cudaMemcpyAsync(ra->round_tallies, out.round_tallies, 4 * n * (size_t)R * E, cudaMemcpyDeviceToHost, st));
        std::vector<int64_t> tw, tr, tt;
        if (ra->winner_hist) { tw.resize(nh_w); CU_TRY(cudaMemcpyAsync(tw.data(), out.winner_hist, 8 * nh_w, cudaMemcpyDeviceToHost, st)); }
        if (ra->rounds_hist) { tr.resize(nh_r); CU_TRY(cudaMemcpyAsync(tr.data(), out.rounds_hist, 8 * nh_r, cudaMemcpyDeviceToHost, st)); }
        if (ra->totals) { tt.resize(nh_t); CU_TRY(cudaMemcpyAsync(tt.data(), out.totals, 8 * nh_t, cudaMemcpyDeviceToHost, st)); }
        CU_TRY(cudaStreamSynchronize(st));
        for (size_t i = 0; i < tw.size(); ++i) ra->winner_hist[i] += tw[i];
        for (size_t i = 0; i < tr.size(); ++i) ra->rounds_hist[i] += tr[i];
        for (size_t i = 0; i < tt.size(); ++i) ra->totals[i] += tt[i];
        float ms = -1.f;
        if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->last_ms = ms;
    }
    return CSIM_OK;
}

// Page-locked host memory for result arrays: device-to-host copies into it run at full PCIe speed and do not stage
// through the driver's bounce buffers.  The Python host keeps a small pool of these (engine.py: _PinnedPool).
extern "C" csim_status csim_host_alloc(uint64_t bytes, void** out) {
    if (!out) return fail(CSIM_ERR_INVALID, "csim_host_alloc: out is NULL");
    *out = nullptr;
    if (bytes == 0) return CSIM_OK;
    CU_TRY(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return CSIM_OK;
}
extern "C" csim_status csim_host_free(void* p) {
    if (p) CU_TRY(cudaFreeHost(p));
    return CSIM_OK;
}

// ---------------------------------------------------------------------------------------------
// csim_engine_probs: the parity probe (header).  Runs ONE trial of the named engine through its PROBE instantiation.
// ---------------------------------------------------------------------------------------------
extern "C" csim_status csim_engine_probs(csim_ctx* c, const csim_params* p, int32_t wiring, int32_t engine, int32_t round,
                                         const int32_t* prev_vote, const int32_t* prev_count, const uint8_t* fatigued,
                                         float* scores_out, float* prefix_out, int32_t prefix_capacity, int32_t* n_chunks_out, int32_t* chunk_out) {
    if (!c || !p || !scores_out || !prefix_out || !n_chunks_out || !chunk_out) return fail(CSIM_ERR_INVALID, "csim_engine_probs: NULL argument");
    if (c->E <= 0) return fail(CSIM_ERR_STATE, "csim_engine_probs: roster not uploaded");
    if (round < 1) return fail(CSIM_ERR_INVALID, "csim_engine_probs: round must be >= 1");
    if (engine != CSIM_ENGINE_DENSE && engine != CSIM_ENGINE_TABLE && engine != CSIM_ENGINE_PREFIX && engine != CSIM_ENGINE_FULL)
        return fail(CSIM_ERR_INVALID, "csim_engine_probs: name one engine (not AUTO)");
    if (wiring != CSIM_WIRING_REF_CLI && wiring != CSIM_WIRING_MODEL) return fail(CSIM_ERR_INVALID, "csim_engine_probs: bad wiring %d", wiring);
    if (engine == CSIM_ENGINE_TABLE && (prev_vote || prev_count || wiring != CSIM_WIRING_REF_CLI))
        return fail(CSIM_ERR_INVALID, "csim_engine_probs: the table engine exists for the ref_cli wiring only (no previous-round inputs)");
    const int E = c->E;
    DeviceGuard guard;
    CU_TRY(cudaSetDevice(c->device));
    if (c->busy_valid) CU_TRY(cudaEventSynchronize(c->ev_busy));
    csim_params q = *p;
    q.max_rounds = round;
    q.runoff_candidate_count = 0;                    // no run-off: round `round` is a full-field round
    int32_t w1 = -1, r1 = 0;
    csim_run_args ra;
    memset(&ra, 0, sizeof ra);
    ra.params = &q; ra.n_param_sets = 1; ra.wiring = wiring; ra.engine = engine; ra.precision = CSIM_FP32;
    ra.sim_id_begin = 0; ra.n_sims = 1; ra.seed = 1; ra.winner = &w1; ra.rounds = &r1;
    if (engine == CSIM_ENGINE_TABLE) {
        if (prefix_capacity < E) return fail(CSIM_ERR_INVALID, "csim_engine_probs: prefix_capacity %d < E = %d (the table engine's rows are whole CDFs)", prefix_capacity, E);
        csim_status s = csim_run(c, &ra);            // builds cdf32 for rounds 1..round exactly as a batch does
        if (s != CSIM_OK) return s;
        std::vector<float> cdf((size_t)E * E);
        CU_TRY(cudaMemcpy(cdf.data(), (const float*)c->cdf32.p + (size_t)(round - 1) * cdf_round_stride(E), sizeof(float) * (size_t)E * E, cudaMemcpyDeviceToHost));
        for (int e = 0; e < E; ++e)
            for (int j = 0; j < E; ++j) {
                const float v = cdf[(size_t)e * E + j];
                prefix_out[(size_t)e * prefix_capacity + j] = v;
                scores_out[(size_t)e * E + j] = j == 0 ? v : v - cdf[(size_t)e * E + j - 1];
            }
        *n_chunks_out = E; *chunk_out = 1;
        return CSIM_OK;
    }
    if (prefix_capacity < (E + 7) / 8) return fail(CSIM_ERR_INVALID, "csim_engine_probs: prefix_capacity %d < ceil(E / 8) = %d", prefix_capacity, (E + 7) / 8);
    ProbeReq req;
    req.round = round; req.cap = prefix_capacity;
    if (prev_vote) { CU_TRY(c->p_prev.reserve(4 * E)); CU_TRY(cudaMemcpy(c->p_prev.p, prev_vote, 4 * E, cudaMemcpyHostToDevice)); req.prev_vote = (const int32_t*)c->p_prev.p; }
    if (prev_count) { CU_TRY(c->p_cnt.reserve(4 * E)); CU_TRY(cudaMemcpy(c->p_cnt.p, prev_count, 4 * E, cudaMemcpyHostToDevice)); req.prev_count = (const int32_t*)c->p_cnt.p; }
    if (fatigued) { CU_TRY(c->p_fat.reserve(E)); CU_TRY(cudaMemcpy(c->p_fat.p, fatigued, E, cudaMemcpyHostToDevice)); req.fatigued = (const uint8_t*)c->p_fat.p; }
    CU_TRY(c->p_out.reserve(sizeof(float) * (size_t)E * E));
    CU_TRY(c->p_pref.reserve(sizeof(float) * (size_t)E * prefix_capacity));
    CU_TRY(cudaMemset(c->p_out.p, 0, sizeof(float) * (size_t)E * E));
    CU_TRY(cudaMemset(c->p_pref.p, 0, sizeof(float) * (size_t)E * prefix_capacity));
    req.scores = (float*)c->p_out.p; req.prefix = (float*)c->p_pref.p;
    c->probe = &req;
    csim_status s = csim_run(c, &ra);
    c->probe = nullptr;
    if (s != CSIM_OK) return s;
    if (req.nch > prefix_capacity) return fail(CSIM_ERR_INVALID, "csim_engine_probs: prefix_capacity %d < %d chunks", prefix_capacity, req.nch);
    CU_TRY(cudaMemcpy(scores_out, c->p_out.p, sizeof(float) * (size_t)E * E, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(prefix_out, c->p_pref.p, sizeof(float) * (size_t)E * prefix_capacity, cudaMemcpyDeviceToHost));
    *n_chunks_out = req.nch; *chunk_out = req.chunk;
    return CSIM_OK;
}

extern "C" csim_status csim_last_run_info(csim_ctx* c, int32_t* engine, int32_t* precision, int32_t* fell_through) {
    if (!c) return fail(CSIM_ERR_INVALID, "csim_last_run_info: ctx is NULL");
    if (c->last_engine == CSIM_ENGINE_AUTO) return fail(CSIM_ERR_STATE, "csim_last_run_info: no csim_run has completed on this context");
    if (engine) *engine = c->last_engine;
    if (precision) *precision = c->last_precision;
    if (fell_through) *fell_through = c->last_fell;
    return CSIM_OK;
}

extern "C" csim_status csim_stats(csim_ctx* c, int64_t* kernel_launches, float* last_trial_kernel_ms) {
    if (!c) return fail(CSIM_ERR_INVALID, "csim_stats: ctx is NULL");
    if (kernel_launches) *kernel_launches = c->launches;
    if (last_trial_kernel_ms) {
        float ms = c->last_ms;
        if (c->ev_valid && cudaEventQuery(c->ev1) == cudaSuccess) {
            float m2 = -1.f;
            if (cudaEventElapsedTime(&m2, c->ev0, c->ev1) == cudaSuccess) ms = c->last_ms = m2;
        }
        *last_trial_kernel_ms = ms;
    }
    return CSIM_OK;
}

extern "C" csim_status csim_microbench(csim_ctx* c, int32_t kind, double* ginstr_per_s) {
    if (!c || !ginstr_per_s) return fail(CSIM_ERR_INVALID, "csim_microbench: NULL argument");
    if (kind < 0 || kind > 3) return fail(CSIM_ERR_INVALID, "csim_microbench: kind %d", kind);
    DeviceGuard guard;
    CU_TRY(cudaSetDevice(c->device));
    float* sink = nullptr;
    CU_TRY(cudaMalloc(&sink, 4));
    const int iters = 1 << 15, threads = 256, grid = c->prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (kind == 0) k_microbench<0><<<grid, threads>>>(sink, iters, 1.0000001f, 1e-9f);
        else if (kind == 1) k_microbench<1><<<grid, threads>>>(sink, iters, 0.f, 0.f);
        else if (kind == 2) k_microbench<2><<<grid, threads>>>(sink, iters, 0.37f, -1.3f);
        else k_microbench<3><<<grid, threads>>>(sink, iters, 0.999f, 0.998f);
        cudaEventRecord(e1);
        ++c->launches;
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaFree(sink); return fail(CSIM_ERR_CUDA, "microbench: %s", cudaGetErrorString(e)); }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        // thread-instructions per loop iteration (kind 2: 8 pairs x {FADD,FMUL,MUFU,FFMA}; kind 3: 8 pairs = 4 x {2 FMUL2, 2 FMNMX, FADD2})
        const double per_iter = kind == 2 ? 32.0 : (kind == 3 ? 20.0 : 8.0);
        const double g = (double)grid * threads * iters * per_iter / (ms * 1e-3) * 1e-9;
        if (rep > 0) best = std::max(best, g);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    *ginstr_per_s = best;
    return CSIM_OK;
}

// the multi-GPU entry points (NCCL bound at run time, csim_group_*)
#include "multigpu.cuh"
