// eng_dense.cu — the fp32 DENSE engine: launcher of k_trials_dense32 (dense32.cuh), the roofline kernel
#include "internal.cuh"
#include "dense32.cuh"

template <int LC, bool MODEL, bool FACT>
static csim_status launch_dense32_t(csim_ctx* c, Dense32Args& a, cudaStream_t st) {
    const int E = a.ro.E, G = a.ro.G, R = a.b.R;
    // Where the round tables live.  FACT: the tables of all R rounds of a set CTA-shared (staged once per set) when that leaves
    // at least as many warps per CTA as per-warp tables would (a warp then stages the round it plays, as in round 1 of this
    // repo); the MUFU form has ONE table per set (ideology, papabile weight) and always shares it.
    const size_t budget = (size_t)(220 * 1024) / CSIM_MINB;                        // CSIM_MINB CTAs per SM
    auto fit = [&](size_t cta, size_t warp) { int w = 8; while (w > 0 && cta + warp * w > budget) --w; return w; };
    const int n_tab_shared = FACT ? R : 1;                                        // (R == 0: no round is ever played, nothing to stage)
    // rounds beyond a.n_tab_full are run-off rounds in every set of the batch: they only read the first kmax candidates
    const int n_full = (FACT && LC == 8) ? std::min(a.n_tab_full, n_tab_shared) : n_tab_shared;
    const int nk = (a.kmax + 7) / 8;
    const size_t ctaA = dense32_cta_smem(a.nch, a.L, G, n_full, n_tab_shared - n_full, nk), warpA = dense32_warp_smem(E, a.nch, a.L, a.kmax, MODEL, false);
    const size_t ctaB = dense32_cta_smem(a.nch, a.L, G, 0), warpB = dense32_warp_smem(E, a.nch, a.L, a.kmax, MODEL, true);
    const int wA = fit(ctaA, warpA), wB = FACT ? fit(ctaB, warpB) : 0;
    const bool shared_tabs = getenv("CSIM_DENSE_WARPTAB") == nullptr ? (wA >= wB && wA > 0) : !(FACT && wB > 0);
    a.cta_tabs = shared_tabs ? 1 : 0;
    a.n_tab = shared_tabs ? n_tab_shared : 0;
    a.n_tab_full = shared_tabs ? n_full : 0;
    a.nk_chunks = nk;
    a.cta_smem = shared_tabs ? ctaA : ctaB;
    a.warp_smem = shared_tabs ? warpA : warpB;
    int wpc = shared_tabs ? wA : wB;
    if (wpc < 1) {                                                               // not at CSIM_MINB CTAs per SM: one CTA, as many warps as fit
        wpc = 8;
        while (wpc > 1 && a.cta_smem + a.warp_smem * wpc > 220 * 1024) --wpc;
        if (a.cta_smem + a.warp_smem * wpc > 220 * 1024) {
            if (shared_tabs && FACT) {                                           // the shared tables themselves are too large: per-warp tables
                a.cta_tabs = 0; a.n_tab = 0; a.n_tab_full = 0; a.cta_smem = ctaB; a.warp_smem = warpB;
                wpc = 8;
                while (wpc > 1 && a.cta_smem + a.warp_smem * wpc > 220 * 1024) --wpc;
            }
            if (a.cta_smem + a.warp_smem * wpc > 220 * 1024)
                return fail(CSIM_ERR_UNSUPPORTED, "roster of %d electors / %d regions needs %zu B of shared memory", E, G, a.cta_smem + a.warp_smem);
        }
    }
    a.warps_per_cta = wpc;
    const size_t smem = a.cta_smem + a.warp_smem * wpc;
    void (*kern)(Dense32Args) = k_trials_dense32<LC, MODEL, FACT, false>;
    if (c->probe) {
        kern = k_trials_dense32<LC, MODEL, FACT, true>;
        a.probe_round = c->probe->round; a.probe_prev_vote = c->probe->prev_vote; a.probe_prev_count = c->probe->prev_count;
        a.probe_scores = c->probe->scores; a.probe_prefix = c->probe->prefix; a.probe_cap = c->probe->cap;
        c->probe->nch = a.nch; c->probe->chunk = a.L;
    }
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpc * 32, smem));
    if (per_sm < 1) return fail(CSIM_ERR_UNSUPPORTED, "dense kernel does not fit on an SM (threads=%d, smem=%zu)", wpc * 32, smem);
    const uint64_t ctas = (uint64_t)c->prop.multiProcessorCount * per_sm;      // persistent: a multiple of the SM count
    // trials are claimed dynamically, `chunk` at a time per warp (common.cuh: WorkQueue); no more CTAs than there are warps' worth of trials
    const uint64_t want = (a.b.n_sims * (uint64_t)std::max(a.b.n_sets, 1) + wpc - 1) / wpc;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, ctas));
    csim_status qs = work_queue(c, a.b.n_sets, a.b.n_sims, (uint64_t)grid * wpc, st, &a.wq);
    if (qs != CSIM_OK) return qs;
    if (FACT && R > 0) {
        const int CP = a.nch * a.L;
        const size_t tb = sizeof(float) * (size_t)a.b.n_sets * R * 4 * CP;
        if (tb / sizeof(float) > 0xffffffffull) return fail(CSIM_ERR_UNSUPPORTED, "dense engine: %d sets x %d rounds of round tables exceed 2^32 entries", a.b.n_sets, R);
        CU_TRY(c->tab.reserve(tb));
        a.tab = (const float*)c->tab.p;
        for (int sr0 = 0; sr0 < a.b.n_sets * R; sr0 += CSIM_GRID_TILE) {
            dim3 g((CP + 127) / 128, std::min(CSIM_GRID_TILE, a.b.n_sets * R - sr0));
            k_round_tables32<<<g, 128, 0, st>>>(a.ro, a.b.sets, (const double*)c->betas.p, R, CP, c->x_mid, (float*)c->tab.p, sr0);
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

csim_status launch_dense32(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const std::vector<double>& hb,
                                  const BatchDev& b, const TrialOut& out, cudaStream_t st) {
    for (const DevSet& s : hs)
        if (s.en_fat || s.en_stop)
            return fail(CSIM_ERR_UNSUPPORTED, "candidate fatigue / stop-candidate are implemented in CSIM_FP64 precision only");
    for (int i = 0; i < ra->n_param_sets; ++i)
        if (ra->params[i].regional_affinity_bonus < 0 || ra->params[i].bandwagon_strength < 0 || ra->params[i].papabile_weight_factor < 0)
            return fail(CSIM_ERR_UNSUPPORTED, "negative bonus / strength / papabile factor: use CSIM_FP64 precision");
    Dense32Args a;
    a.ro = roster_dev(c); a.b = b; a.out = out; a.nbeta = (const float*)c->nbeta.p; a.wiring = ra->wiring; a.tab = nullptr;
    a.probe_round = 0; a.probe_prev_vote = nullptr; a.probe_prev_count = nullptr; a.probe_scores = nullptr; a.probe_prefix = nullptr; a.probe_cap = 0;
    int kmax = 1;
    for (const DevSet& s : hs) kmax = std::max(kmax, s.k);
    a.kmax = kmax;
    // the rounds that can be full-field in SOME set of the batch: once a run-off exists (k > 0) every round after
    // max(runoff_threshold_rounds, 1) is a run-off round (simulate.py:183-186)
    a.n_tab_full = 1;
    for (const DevSet& s : hs) a.n_tab_full = std::max(a.n_tab_full, s.k > 0 ? std::min(b.R, std::max(s.rr, 1)) : b.R);
    a.nk_chunks = 1;
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

