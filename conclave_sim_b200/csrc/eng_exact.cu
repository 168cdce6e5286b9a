// eng_exact.cu — the fp64 exact DENSE engine: launcher of k_trials_exact64 (exact64.cuh)
#include "internal.cuh"
#include "exact64.cuh"

csim_status launch_exact64(csim_ctx* c, const csim_run_args* ra, const std::vector<DevSet>& hs, const BatchDev& b,
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
    s = work_queue(c, 1, (uint64_t)b.n_sets * b.n_sims, (uint64_t)grid * wpc, st, &a.wq);
    if (s != CSIM_OK) return s;
    CU_TRY(cudaEventRecord(c->ev0, st));
    k_trials_exact64<<<grid, wpc * 32, smem, st>>>(a);
    CU_TRY(cudaEventRecord(c->ev1, st));
    c->ev_valid = true;
    ++c->launches;
    CU_TRY(cudaGetLastError());
    return CSIM_OK;
}

