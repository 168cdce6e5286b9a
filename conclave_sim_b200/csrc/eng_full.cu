// eng_full.cu — the FULL engine: launcher of k_trials_full32 (full32.cuh)
#include "internal.cuh"
#include "full32.cuh"
#ifndef CSIM_EXT_NE
#define CSIM_EXT_NE 1                  // electors of a lane evaluated together by the EXT instantiation: 1 (the plain dense kernel shares the
#endif                                 // candidate operands between 2: here the kernel is latency-bound and 20 warps of 96 registers with half
#define CSIM_NE CSIM_EXT_NE            // the prefix columns beat 16 warps of 128; this unit holds no other instantiation of the dense kernel)
#include "dense32.cuh"

// ---------------------------------------------------------------------------------------------
// The FULL engine's fast path: the EXT instantiation of the dense kernel (dense32.cuh) — the factorised two-elector pair loop with
// candidate fatigue folded into a per-warp copy of the round table and stop-candidate as a byte-masked second accumulator.  It takes
// what that formulation can express: E <= 255 (chunks of 8, tallies in bytes), every factor non-negative, a factorisable beta
// schedule, tables that fit in shared memory.  Everything else stays on k_trials_full32.  `*taken` says whether it ran.
// ---------------------------------------------------------------------------------------------
// Host-side plan of the fast path (pure arithmetic): layout, warps per CTA, whether the round tables of a set are CTA-shared.
//   n_tab_full = rounds that can be full-field in some set of the batch; min_warps = fewest warps per CTA worth launching
// Returns false when the shape is not the fast path's (E outside 2 .. 255, no round) or fewer than min_warps warps fit.
static bool ext_plan(Dense32Args& a, int E, int G, int R, int n_tab_full, int kmax, int Fmax, int min_warps) {
    if (E > 255 || E < 2 || R < 1) return false;
    a.ro.E = E; a.ro.G = G;
    a.kmax = kmax; a.Fmax = Fmax < 1 ? 1 : Fmax;
    a.L = 8; a.nch = (E + 7) / 8;
    a.epitch = (E + 1) & ~1;
    const int nk = (kmax + 7) / 8;
    const int n_full = std::max(1, std::min(n_tab_full, R));
    const size_t budget = 220 * 1024;
    // CTA-shared round tables of all R rounds (whole tables for the full-field rounds, one short table per run-off round) when that
    // costs no warp; else every warp stages the round it plays (its own table is there anyway, the phi-folded copy: ~50 instructions
    // per round, while a warp more or less is ~5 % of the kernel's speed)
    const size_t warp = dense32_warp_smem(E, a.nch, 8, kmax, true, true) + dense32_ext_warp_smem(E, a.nch, G, a.Fmax);
    const size_t ext_cta = dense32_ext_cta_smem(E, a.nch, G, a.epitch);
    const size_t ctaA = (dense32_cta_smem(a.nch, 8, G, n_full, R - n_full, nk) + 15) & ~(size_t)15;
    const size_t ctaB = (dense32_cta_smem(a.nch, 8, G, 0) + 15) & ~(size_t)15;
    int max_wpc = CSIM_EXT_THREADS / 32;
    if (const char* ev = getenv("CSIM_EXT_WPC")) max_wpc = std::max(2, std::min(max_wpc, atoi(ev)));     // tuning aid
    auto fit = [&](size_t cta) { int w = max_wpc; while (w > 0 && cta + ext_cta + warp * w > budget) --w; return w; };
    const int wA = fit(ctaA), wB = fit(ctaB);
    const bool shared_tabs = wA >= wB && wA > 0;
    const size_t base_cta = shared_tabs ? ctaA : ctaB;
    a.cta_tabs = shared_tabs ? 1 : 0; a.n_tab = shared_tabs ? R : 0; a.n_tab_full = shared_tabs ? n_full : 0; a.nk_chunks = nk;
    const int wpc = shared_tabs ? wA : wB;
    // many regions (bonus rows per region, CTA-shared and per warp): below ~10 warps per SM this kernel is no faster than
    // k_trials_full32 with its 28 (12 / 8 warps measured 17 / 21 ms per 250 k trials against 21), so the caller asks for 10 when the
    // general kernel fits and for 2 when it does not
    if (wpc < min_warps) return false;
    a.o_ext = base_cta; a.cta_smem = base_cta + ext_cta; a.warp_smem = warp;
    a.warps_per_cta = wpc;
    return true;
}

static csim_status launch_dense_ext(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b, const TrialOut& out,
                                    cudaStream_t st, bool any_stop, int accw, int min_warps, bool* taken) {
    *taken = false;
    const int E = c->E, G = c->G, R = b.R;
    if (getenv("CSIM_FULL_LEGACY") || R < 1) return CSIM_OK;
    for (int i = 0; i < ra->n_param_sets; ++i) {
        const csim_params& q = ra->params[i];
        std::vector<double> hb((size_t)R);
        csim_beta_schedule(&q, 1, R, hb.data());
        for (double be : hb) if (!(be >= 0.0) || !(be * c->x_half <= 26.0)) return CSIM_OK;      // not factorisable (dense32.cuh header)
    }
    Dense32Args a;
    memset(&a, 0, sizeof a);
    int kmax = 1, Fmax = 1, n_tab_full = 1;
    for (const DevSet& s : hs) {
        kmax = std::max(kmax, s.k); Fmax = std::max(Fmax, s.fat_rounds);
        n_tab_full = std::max(n_tab_full, s.k > 0 ? std::min(R, std::max(s.rr, 1)) : R);
    }
    if (!ext_plan(a, E, G, R, n_tab_full, kmax, Fmax, min_warps)) return CSIM_OK;
    a.ro = roster_dev(c); a.b = b; a.out = out; a.nbeta = (const float*)c->nbeta.p; a.wiring = CSIM_WIRING_MODEL;
    a.accw = accw; a.accmask = (const uint32_t*)c->accm.p;
    const int wpc = a.warps_per_cta;
    const size_t smem = a.cta_smem + a.warp_smem * wpc;
    void (*kern)(Dense32Args) = k_trials_dense32<8, true, true, false, true>;
    if (c->probe) {
        kern = k_trials_dense32<8, true, true, true, true>;
        a.probe_round = c->probe->round; a.probe_prev_vote = c->probe->prev_vote; a.probe_prev_count = c->probe->prev_count;
        a.probe_fatigued = c->probe->fatigued;
        a.probe_scores = c->probe->scores; a.probe_prefix = c->probe->prefix; a.probe_cap = c->probe->cap;
        c->probe->nch = a.nch; c->probe->chunk = 8;
    }
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpc * 32, smem));
    if (per_sm < 1) return CSIM_OK;
    const int CP = a.nch * 8;
    const size_t tb = sizeof(float) * (size_t)b.n_sets * R * 4 * CP;
    if (tb / sizeof(float) > 0xffffffffull) return CSIM_OK;
    CU_TRY(c->tab.reserve(tb));
    a.tab = (const float*)c->tab.p;
    for (int sr0 = 0; sr0 < b.n_sets * R; sr0 += CSIM_GRID_TILE) {
        dim3 g((CP + 127) / 128, std::min(CSIM_GRID_TILE, b.n_sets * R - sr0));
        k_round_tables32<<<g, 128, 0, st>>>(a.ro, b.sets, (const double*)c->betas.p, R, CP, c->x_mid, (float*)c->tab.p, sr0);
        ++c->launches;
    }
    if (any_stop) {                                           // the acceptable bits again, as bytes per (chunk, elector slot)
        const size_t per_set = (size_t)a.nch * a.epitch;
        CU_TRY(c->accb.reserve(sizeof(uint2) * per_set * b.n_sets));
        CU_TRY(cudaMemsetAsync(c->accb.p, 0, sizeof(uint2) * per_set * b.n_sets, st));
        a.accb = (const uint2*)c->accb.p;
        for (int s0 = 0; s0 < b.n_sets; s0 += CSIM_GRID_TILE) {
            dim3 g((E * a.nch + 127) / 128, std::min(CSIM_GRID_TILE, b.n_sets - s0));
            k_ext_accbytes<<<g, 128, 0, st>>>(a.ro, b.sets + s0, a.nch, a.epitch, (uint2*)c->accb.p + (size_t)s0 * per_set);
            ++c->launches;
        }
    }
    const uint64_t ctas = (uint64_t)c->prop.multiProcessorCount * per_sm;
    const uint64_t want = (a.b.n_sims * (uint64_t)std::max(a.b.n_sets, 1) + wpc - 1) / wpc;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, ctas));
    csim_status qs = work_queue(c, a.b.n_sets, a.b.n_sims, (uint64_t)grid * wpc, st, &a.wq);
    if (qs != CSIM_OK) return qs;
    CU_TRY(cudaEventRecord(c->ev0, st));
    kern<<<grid, wpc * 32, smem, st>>>(a);
    CU_TRY(cudaEventRecord(c->ev1, st));
    c->ev_valid = true;
    ++c->launches;
    CU_TRY(cudaGetLastError());
    *taken = true;
    return CSIM_OK;
}

// ---------------------------------------------------------------------------------------------
// FULL engine (full32.cuh): fp32, every model term inline, fatigue + stop-candidate included
// ---------------------------------------------------------------------------------------------
csim_status launch_full32(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b,
                                 const TrialOut& out, cudaStream_t st) {
    if (ra->precision != CSIM_FP32) return fail(CSIM_ERR_UNSUPPORTED, "the full engine is an fp32 engine (CSIM_FP64: use DENSE)");
    const int E = c->E, R = b.R;
    if (E >= 32768) return fail(CSIM_ERR_UNSUPPORTED, "full engine: at most 32767 electors");
    Full32Args a;
    memset(&a, 0, sizeof a);
    int kmax = 1, Fmax = 1;
    for (const DevSet& s : hs) { kmax = std::max(kmax, s.k); Fmax = std::max(Fmax, s.fat_rounds); }
    const bool general_fits = full32_plan(a, E, c->G, R, kmax, Fmax, (size_t)c->prop.sharedMemPerBlockOptin);
    a.ro = roster_dev(c); a.b = b; a.out = out; a.nbeta = (const float*)c->nbeta.p;
    a.accw = (E + 31) / 32; a.accmask = nullptr;
    bool any_stop = false;
    for (const DevSet& s : hs) any_stop = any_stop || s.en_stop;
    if (any_stop) {                                           // acceptable(e, c) bits, fp64-exact, once per batch
        CU_TRY(c->accm.reserve(sizeof(uint32_t) * (size_t)b.n_sets * E * a.accw));
        a.accmask = (const uint32_t*)c->accm.p;
        for (int s0 = 0; s0 < b.n_sets; s0 += CSIM_GRID_TILE) {
            dim3 g((E * a.accw + 127) / 128, std::min(CSIM_GRID_TILE, b.n_sets - s0));
            k_full_accmask<<<g, 128, 0, st>>>(a.ro, b.sets + s0, a.accw, (uint32_t*)c->accm.p + (size_t)s0 * E * a.accw);
            ++c->launches;
        }
    }
    const int wpc = a.warps_per_cta;
    const size_t smem = a.cta_smem + a.warp_smem * wpc;
    bool ext = false, neg = false;                            // which terms the batch needs decides the instantiation
    for (const DevSet& s : hs) ext = ext || s.en_fat || s.en_stop;
    for (int i = 0; i < ra->n_param_sets; ++i) {
        const csim_params& q = ra->params[i];
        neg = neg || q.regional_affinity_bonus < 0 || q.bandwagon_strength < 0 || q.papabile_weight_factor < 0 || q.stickiness_factor < -1 ||
              q.fatigue_penalty_factor < 0 || q.stop_candidate_boost_factor < 0;
    }
    if (ext && !neg) {                                        // fatigue / stop-candidate with non-negative factors: the dense EXT kernel
        bool taken = false;
        csim_status es = launch_dense_ext(c, ra, hs, b, out, st, any_stop, a.accw, general_fits ? 10 : 2, &taken);
        if (es != CSIM_OK || taken) return es;
    }
    if (!general_fits)
        return fail(CSIM_ERR_UNSUPPORTED, "full engine: a roster of %d electors (fatigue window %d) does not fit in shared memory", E, Fmax);
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
    // trials are claimed dynamically, `chunk` at a time per warp (common.cuh: WorkQueue)
    const uint64_t want = (a.b.n_sims * (uint64_t)std::max(a.b.n_sets, 1) + wpc - 1) / wpc;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, ctas));
    csim_status qs = work_queue(c, a.b.n_sets, a.b.n_sims, (uint64_t)grid * wpc, st, &a.wq);
    if (qs != CSIM_OK) return qs;
    CU_TRY(cudaEventRecord(c->ev0, st));
    kern<<<grid, wpc * 32, smem, st>>>(a);
    CU_TRY(cudaEventRecord(c->ev1, st));
    c->ev_valid = true;
    ++c->launches;
    CU_TRY(cudaGetLastError());
    return CSIM_OK;
}


// ---------------------------------------------------------------------------------------------
// csim_full_plan_query (header): which kernel the FULL engine would run for a batch with fatigue / stop-candidate live, and its budget
// ---------------------------------------------------------------------------------------------
extern "C" csim_status csim_full_plan_query(int32_t n_electors, int32_t n_regions, int32_t max_rounds, int32_t full_field_rounds, int32_t kmax,
                                            int32_t fatigue_window, int32_t fast_path_eligible, csim_full_plan* out) {
    if (!out) return fail(CSIM_ERR_INVALID, "csim_full_plan_query: out is NULL");
    if (n_electors < 1 || n_electors >= 32768 || n_regions < 1 || n_regions > n_electors || max_rounds < 0 || kmax < 0)
        return fail(CSIM_ERR_INVALID, "csim_full_plan_query: need 1 <= electors < 32768, 1 <= regions <= electors, rounds >= 0, kmax >= 0");
    memset(out, 0, sizeof *out);
    Full32Args g;
    memset(&g, 0, sizeof g);
    const bool general = full32_plan(g, n_electors, n_regions, max_rounds, std::max(kmax, 1), std::max(fatigue_window, 1), (size_t)232448);
    out->general_fits = general ? 1 : 0;
    Dense32Args a;
    memset(&a, 0, sizeof a);
    const bool fast = fast_path_eligible && !getenv("CSIM_FULL_LEGACY") &&
                      ext_plan(a, n_electors, n_regions, max_rounds, full_field_rounds, std::max(kmax, 1), std::max(fatigue_window, 1), general ? 10 : 2);
    if (fast) {
        out->fast_path = 1; out->warps_per_cta = a.warps_per_cta; out->cta_tabs = a.cta_tabs; out->n_chunks = a.nch; out->acc_pitch = a.epitch;
        out->smem_cta = a.cta_smem; out->smem_per_warp = a.warp_smem; out->smem_total = a.cta_smem + a.warp_smem * (uint64_t)a.warps_per_cta;
    } else if (general) {
        out->warps_per_cta = g.warps_per_cta; out->n_chunks = g.nch;
        out->smem_cta = g.cta_smem; out->smem_per_warp = g.warp_smem; out->smem_total = g.cta_smem + g.warp_smem * (uint64_t)g.warps_per_cta;
    } else {
        return fail(CSIM_ERR_UNSUPPORTED, "full engine: a roster of %d electors / %d regions fits neither of its kernels", n_electors, n_regions);
    }
    return CSIM_OK;
}
