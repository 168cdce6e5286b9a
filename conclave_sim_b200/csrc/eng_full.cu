// eng_full.cu — the FULL engine: launcher of k_trials_full32 (full32.cuh)
#include "internal.cuh"
#include "full32.cuh"

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
    if (!full32_plan(a, E, c->G, R, kmax, Fmax, (size_t)c->prop.sharedMemPerBlockOptin))
        return fail(CSIM_ERR_UNSUPPORTED, "full engine: a roster of %d electors (fatigue window %d) does not fit in shared memory", E, Fmax);
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

