// eng_prefix.cu — the PREFIX engine: chunk-prefix table builder, launch plan + launcher of k_trials_prefix (prefix.cuh)
#include "internal.cuh"
#include "prefix.cuh"

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

// preconditions of the fp32 PREFIX engine (launch_prefix reports the same conditions as errors)
bool prefix_supported(const csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, int R) {
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

csim_status launch_prefix(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b,
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

