// cabi.cu — implementation of include/conclave_b200.h (the C ABI) on top of the kernels.
//
// Host-side responsibilities: validate arguments, derive the per-set constants exactly as the
// reference derives them in fp64 (beta schedule model.py:562-580, vote threshold
// simulate.py:175), size grids as multiples of the SM count, and move buffers when the caller
// hands host memory.  There is no CPU implementation of the hot path in this file.
#include "internal.cuh"
#include "exact64.cuh"      // k_base_scores64 / k_prob_matrix64 of csim_prob_matrix
#include "prob32.cuh"
#include "table.cuh"        // cdf_round_stride (csim_engine_probs reads the table engine's fp32 CDF rows back)

static thread_local std::string g_err;

csim_status fail(csim_status st, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return st;
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
    cudaEventCreateWithFlags(&c->ev_stage, cudaEventDisableTiming);
    *out = c;
    return CSIM_OK;
}

extern "C" csim_status csim_destroy(csim_ctx* c) {
    if (!c) return CSIM_OK;
    DeviceGuard guard;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();             // nothing of this context may still be in flight when its buffers go
    for (DevBuf* b : {&c->x64, &c->x32, &c->region, &c->pap, &c->sets, &c->betas, &c->nbeta, &c->W, &c->tab, &c->cdf64, &c->cdfk64, &c->cdf32, &c->cdfk32, &c->ptab, &c->accm, &c->accb, &c->work, &c->o_winner, &c->o_rounds,
                      &c->o_reason, &c->o_final, &c->o_tallies, &c->o_whist, &c->o_rhist, &c->o_totals, &c->i_tape,
                      &c->p_prev, &c->p_cnt, &c->p_fat, &c->p_shares, &c->p_bw, &c->p_out, &c->p_cand, &c->p_pref})
        b->release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_busy) cudaEventDestroy(c->ev_busy);
    if (c->ev_stage) cudaEventDestroy(c->ev_stage);
    if (c->h_stage) cudaFreeHost(c->h_stage);
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
    {   // through page-locked staging (see csim_ctx::h_stage): wait only for the PREVIOUS call's copies out of it, never for its kernels
        const size_t b_sets = (sizeof(DevSet) * hs.size() + 15) & ~(size_t)15, b_betas = (sizeof(double) * hb.size() + 15) & ~(size_t)15;
        const size_t b_nb = sizeof(float) * hnb.size(), need = b_sets + b_betas + b_nb;
        if (c->stage_valid) CU_TRY(cudaEventSynchronize(c->ev_stage));
        if (need > c->h_stage_cap) {
            if (c->h_stage) cudaFreeHost(c->h_stage);
            c->h_stage = nullptr; c->h_stage_cap = 0;
            CU_TRY(cudaHostAlloc(&c->h_stage, need * 2, cudaHostAllocDefault));
            c->h_stage_cap = need * 2;
        }
        char* hp = (char*)c->h_stage;
        memcpy(hp, hs.data(), sizeof(DevSet) * hs.size());
        memcpy(hp + b_sets, hb.data(), sizeof(double) * hb.size());
        memcpy(hp + b_sets + b_betas, hnb.data(), b_nb);
        CU_TRY(cudaMemcpyAsync(c->sets.p, hp, sizeof(DevSet) * hs.size(), cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(c->betas.p, hp + b_sets, sizeof(double) * hb.size(), cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(c->nbeta.p, hp + b_sets + b_betas, b_nb, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaEventRecord(c->ev_stage, st));
        c->stage_valid = true;
    }

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
        std::vector<float> cdf(cdf_round_stride(E));
        CU_TRY(cudaMemcpy(cdf.data(), (const float*)c->cdf32.p + (size_t)(round - 1) * cdf_round_stride(E), sizeof(float) * cdf_round_stride(E), cudaMemcpyDeviceToHost));
        const int eqe = table_e_quad_end(E);
        for (int e = 0; e < E; ++e) {
            const float* row = cdf.data() + (size_t)table_slot(e, eqe) * E;     // rows are stored in lane-visiting order
            for (int j = 0; j < E; ++j) {
                prefix_out[(size_t)e * prefix_capacity + j] = row[j];
                scores_out[(size_t)e * E + j] = j == 0 ? row[j] : row[j] - row[j - 1];
            }
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
        // thread-instructions per loop iteration (kind 2: 8 pairs x {FADD,FMUL,MUFU,FFMA}; kind 3: 8 pairs = 4 x {FMUL2, 2 FMNMX, FADD2})
        const double per_iter = kind == 2 ? 32.0 : (kind == 3 ? 16.0 : 8.0);
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
