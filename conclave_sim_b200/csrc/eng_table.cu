// eng_table.cu — the TABLE engine: CDF table builders + launcher of k_trials_table (table.cuh)
#include "internal.cuh"
#include "table.cuh"

template <class Real>
static csim_status launch_table_t(csim_ctx* c, const BatchDev& b, const TrialOut& out, int kmax, const Real* cdf, const Real* cdfk,
                                  cudaStream_t st) {
    TableArgs<Real> a;
    a.ro = roster_dev(c); a.b = b; a.out = out; a.cdf = cdf; a.cdfk = cdfk; a.kmax = kmax;
    const int E = c->E;
    const size_t Ei = ((size_t)E + 3) & ~(size_t)3;
    // fp32: the tail electors' Philox words of a round are computed once for the warp's 32 trials (table.cuh: tailbuf)
    const int n_tail = E - table_e_quad_end(E);
    a.tail_words = (sizeof(Real) == 4 && n_tail > 0 && n_tail <= 16 && getenv("CSIM_TABLE_NOTAIL") == nullptr) ? 4 * ((n_tail + 3) / 4) : 0;
    auto warp_bytes = [&](int tail_words) { return (Ei * 4 + (((size_t)kmax + 3) & ~(size_t)3) * 4 + ((Ei + 15) & ~(size_t)15) + (size_t)tail_words * 32 * 4 + 15) & ~(size_t)15; };
    {   // the tail buffer must not cost the second resident CTA (2 x 16 warps per SM is what hides the search latency)
        const size_t fixed = ((cdf_round_stride(E) * sizeof(Real) + 15) & ~(size_t)15) + ((((size_t)b.R * table_rows(E) * sizeof(Real)) + 15) & ~(size_t)15);
        const size_t half = ((size_t)c->prop.sharedMemPerMultiprocessor - 2 * 1024) / 2;
        if (a.tail_words && fixed + warp_bytes(0) * CSIM_TABLE_WPC <= half && fixed + warp_bytes(a.tail_words) * CSIM_TABLE_WPC > half) a.tail_words = 0;
    }
    a.warp_smem = warp_bytes(a.tail_words);
    const size_t qtab = (((size_t)b.R * table_rows(E) * sizeof(Real)) + 15) & ~(size_t)15;
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
    // Super-batches are claimed dynamically (one global counter).  Big ones of wpc x 32 trials amortise the per-round staging; the
    // last ~one-big-batch-per-CTA's worth of the job is cut into small ones (16 trials per warp) so that the CTAs finish within
    // one small batch of each other instead of one big one (1.25 M trials on 296 CTAs: 8.25 big batches per CTA).
    const uint64_t ctas = (uint64_t)c->prop.multiProcessorCount * per_sm;
    int tpw = 32, tpw_tail = 16;
    if (const char* e = getenv("CSIM_TABLE_TPW")) tpw = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("CSIM_TABLE_TPW_TAIL")) tpw_tail = std::max(1, std::min(32, atoi(e)));
    tpw_tail = std::min(tpw_tail, tpw);
    a.tpw = tpw; a.tpw_tail = tpw_tail;
    const uint64_t big_trials = (uint64_t)wpc * tpw, small_trials = (uint64_t)wpc * tpw_tail;
    uint64_t tail_sims = std::min<uint64_t>(b.n_sims, ctas * big_trials);
    if (const char* e = getenv("CSIM_TABLE_TAIL_BATCHES")) tail_sims = std::min<uint64_t>(b.n_sims, (uint64_t)atoll(e) * big_trials);
    a.n_big_last = (b.n_sims - tail_sims) / big_trials;
    const uint64_t sb_per_set = (b.n_sims + big_trials - 1) / big_trials;
    const uint64_t n_sb = sb_per_set * (uint64_t)(b.n_sets - 1) + a.n_big_last + (b.n_sims - a.n_big_last * big_trials + small_trials - 1) / small_trials;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_sb, ctas));
    csim_status qs = work_queue(c, 1, n_sb, (uint64_t)grid, st, &a.wq);
    if (qs != CSIM_OK) return qs;
    CU_TRY(cudaEventRecord(c->ev0, st));
    kern<<<grid, wpc * 32, smem, st>>>(a);
    CU_TRY(cudaEventRecord(c->ev1, st));
    c->ev_valid = true;
    ++c->launches;
    CU_TRY(cudaGetLastError());
    return CSIM_OK;
}

csim_status launch_table(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b,
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

