// dense32.cuh — the fp32 dense engine: every trial re-evaluates all E x C scores each round.
//
// This is the kernel the roofline is quoted on (SURVEY §8d "canonical dense per-trial
// formulation").  One warp owns one trial and walks it from round 1 to its end, then takes the
// next trial (static stride over trials: results never depend on the schedule).  No block-level
// barrier exists anywhere; only __syncwarp.
//
// A lane owns the electors of whole Philox quads (4q..4q+3, one Philox call -> four 24-bit
// uniforms) plus at most a few leftover singles.  For one elector and one full-field round:
//
//   phase 1  for every candidate c:  v = ex2(|x_e - x_c| * nb_r) ; chunk_sum[c / L] += v * pw_c
//            4 issue slots per (elector, candidate) pair: FADD, FMUL(|.|), MUFU.EX2, FFMA; the
//            candidate operands come from shared memory as warp-uniform (broadcast) 128-bit loads.
//            The terms that do not need the exponential are added per CHUNK, not per pair:
//            regional bonus  = bonus * #{c in chunk : region_c == region_e}   (roster table)
//            bandwagon       = sum of bw_c over the chunk                       (per trial-round)
//            self vote       = - s_ee  in e's chunk            (model.py:418-424)
//            stickiness      = + st * s_e,prev  in prev's chunk (model.py:357-367)
//   phase 2  S = sum of chunk sums, t = u * S; the chunk whose prefix crosses t is found on the
//            register-resident chunk sums, then ONLY that chunk's L candidates are re-evaluated
//            with every term inline, in roster order, to find the first c with cum > t
//            (== searchsorted(cumsum(p)/sum, u, 'right'), simulate.py:153).
//
// Run-off rounds (simulate.py:144-154,183-186) only need the first k columns of the row
// (P[e, :k] renormalised), so they cost k pair evaluations per elector.
#pragma once
#include "common.cuh"

struct Dense32Args {
    RosterDev ro;
    BatchDev b;
    TrialOut out;
    const float* nbeta;     // [sets][R]  -beta_r * log2(e)
    int wiring;
    int L;                  // candidates per chunk (multiple of 4); NCH * L >= C
    int kmax;
    int warps_per_cta;
    size_t cta_smem;        // bytes of CTA-shared tables
    size_t warp_smem;       // bytes per warp
};

template <int NCH>
__host__ __device__ inline size_t dense32_cta_smem(int L, int G) {
    const size_t CP = (size_t)NCH * L;
    return ((CP * 4 /*xs*/ + CP * 4 /*reg*/ + (size_t)NCH * G * 4 /*regcnt*/) + 15) & ~(size_t)15;
}
template <int NCH>
__host__ __device__ inline size_t dense32_warp_smem(int E, int L, int kmax, bool model) {
    const size_t CP = (size_t)NCH * L;
    size_t n = CP * 4 /*pw*/ + (size_t)E * 4 /*tally*/ + (size_t)kmax * 4 /*eff*/ + (((size_t)E + 3) & ~(size_t)3) /*mark*/;
    if (model) n += (size_t)E * 4 * 2 /*prev_tally, votes*/ + CP * 4 /*bw*/ + (size_t)NCH * 4 /*bwchunk*/;
    return (n + 15) & ~(size_t)15;
}

// every term of one score, inline (phase 2 and run-off rounds)
template <bool MODEL>
__device__ __forceinline__ float full_score32(int e, int c, float xe, int ge, float nb, float bonus, float st1p, int prev,
                                              bool use_bw, const float* xs, const int32_t* reg, const float* pw, const float* bw) {
    float v = ex2_approx(fabsf(xe - xs[c]) * nb) * pw[c];
    if (reg[c] == ge) v += bonus;
    if (MODEL) {
        if (use_bw) v += bw[c];
        if (c == prev) v *= st1p;
    }
    return c == e ? 0.0f : v;
}

template <int NCH, int LC, bool MODEL>
__global__ void __launch_bounds__(256) k_trials_dense32(Dense32Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int E = a.ro.E, C = E, R = a.b.R, G = a.ro.G;
    const int L = LC > 0 ? LC : a.L;
    const int CP = NCH * L;
    // ---- CTA-shared roster tables ------------------------------------------------------------
    float* xs = (float*)smem_raw;
    int32_t* reg = (int32_t*)(xs + CP);
    float* regcnt = (float*)(reg + CP);
    for (int c = threadIdx.x; c < CP; c += blockDim.x) {
        xs[c] = c < C ? a.ro.x32[c] : 0.0f;
        reg[c] = c < C ? a.ro.region[c] : -1;
    }
    for (int i = threadIdx.x; i < NCH * G; i += blockDim.x) {
        const int j = i / G, g = i - j * G;
        int n = 0;
        for (int c = j * L; c < (j + 1) * L && c < C; ++c) n += (a.ro.region[c] == g);
        regcnt[i] = (float)n;
    }
    // ---- per-warp state ----------------------------------------------------------------------
    unsigned char* base = smem_raw + a.cta_smem + (size_t)wib * a.warp_smem;
    float* pw = (float*)base;                   base += sizeof(float) * CP;
    int32_t* tally = (int32_t*)base;            base += sizeof(int32_t) * E;
    int32_t* eff = (int32_t*)base;              base += sizeof(int32_t) * a.kmax;
    uint8_t* mark = (uint8_t*)base;             base += ((size_t)E + 3) & ~(size_t)3;
    int32_t* prev_tally = nullptr; int32_t* votes = nullptr; float* bw = nullptr; float* bwchunk = nullptr;
    if (MODEL) {
        prev_tally = (int32_t*)base;            base += sizeof(int32_t) * E;
        votes = (int32_t*)base;                 base += sizeof(int32_t) * E;
        bw = (float*)base;                      base += sizeof(float) * CP;
        bwchunk = (float*)base;
    }
    __syncthreads();   // the only block barrier: roster tables are read-only from here on

    const uint64_t n_trials = (uint64_t)a.b.n_sets * a.b.n_sims;
    const uint64_t warp_global = (uint64_t)blockIdx.x * a.warps_per_cta + wib;
    const uint64_t n_warps = (uint64_t)gridDim.x * a.warps_per_cta;
    const int m_quads = E >> 7;                              // whole quads per lane
    const int n_main = 4 * m_quads;
    const int e_single0 = 128 * m_quads + lane;
    const int n_single = e_single0 < E ? (E - e_single0 + 31) >> 5 : 0;
    const int n_slots = n_main + n_single;

    int cur_set = -1;
    float bonus = 0.f, st1p = 1.f, bw_strength = 0.f;
    int need = 0, k = 0, rr = 0;
    bool has_sticky = false, has_bw = false;
    unsigned long long acc_rounds = 0, acc_trials = 0;

    for (uint64_t t = warp_global; t < n_trials; t += n_warps) {
        const int set = (int)(t / a.b.n_sims);
        const uint64_t sim = a.b.sim_id_begin + (t % a.b.n_sims);
        if (set != cur_set) {
            if (cur_set >= 0 && lane == 0 && a.out.totals) {
                atomicAdd(&a.out.totals[2 * cur_set], acc_rounds); atomicAdd(&a.out.totals[2 * cur_set + 1], acc_trials);
            }
            acc_rounds = acc_trials = 0;
            cur_set = set;
            const DevSet* S = &a.b.sets[set];
            bonus = S->f_bonus; st1p = S->f_st1p; bw_strength = S->f_bandwagon;
            need = S->need; k = S->k; rr = S->rr;
            has_sticky = MODEL && S->has_sticky; has_bw = MODEL && S->has_bw;
            const float pwf = S->f_pwf;
            __syncwarp();
            for (int c = lane; c < CP; c += 32) pw[c] = c < C ? (a.ro.pap[c] ? pwf : 1.0f) : 0.0f;
            if (MODEL) for (int c = lane; c < CP; c += 32) bw[c] = 0.0f;
            __syncwarp();
        }
        const float* nbeta = a.nbeta + (size_t)set * R;
        bool have_eff = false, have_prev = false;
        int winner = -1, rounds = 0, reason = CSIM_REASON_MAX_ROUNDS;

        for (int r = 1; r <= R; ++r) {
            rounds = r;
            const float nb = __ldg(&nbeta[r - 1]);
            bool use_bw = false;
            if (MODEL) {
                if (has_bw && have_prev) {                                   // model.py:326-354
                    int mx, arg;
                    warp_argmax_first(prev_tally, C, lane, mx, arg);
                    if (mx > 0) {
                        use_bw = true;
                        const float inv = 1.0f / (float)mx;
                        for (int c = lane; c < C; c += 32) bw[c] = (float)prev_tally[c] * inv * bw_strength;
                        __syncwarp();
                        if (lane < NCH) {
                            float s = 0.f;
                            for (int c = lane * L; c < (lane + 1) * L; ++c) s += bw[c];
                            bwchunk[lane] = s;
                        }
                    }
                }
            }
            for (int c = lane; c < C; c += 32) tally[c] = 0;
            __syncwarp();

            uint32_t w[4] = {0, 0, 0, 0};
            for (int n = 0; n < n_slots; ++n) {
                int e;
                if (n < n_main) {
                    const int q = lane + 32 * (n >> 2);
                    e = 4 * q + (n & 3);
                    if ((n & 3) == 0)
                        philox4x32_10((uint32_t)sim, (uint32_t)(sim >> 32), (uint32_t)r, (uint32_t)q, a.b.seed_lo, a.b.seed_hi, w);
                } else {
                    e = e_single0 + 32 * (n - n_main);
                    philox4x32_10((uint32_t)sim, (uint32_t)(sim >> 32), (uint32_t)r, (uint32_t)(e >> 2), a.b.seed_lo, a.b.seed_hi, w);
                }
                const int sel = e & 3;
                const uint32_t wa = sel == 0 ? w[0] : (sel == 1 ? w[1] : (sel == 2 ? w[2] : w[3]));
                const float u = u24_from_word(wa);
                const float xe = xs[e];
                const int ge = reg[e];
                const int prev = (MODEL && have_prev && has_sticky) ? votes[e] : -1;
                int vote;
                if (!have_eff) {
                    // ---------------- phase 1: chunk sums of the exponential part ----------------
                    float cs[NCH];
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        float acc0 = 0.f, acc1 = 0.f;
                        const float4* x4 = reinterpret_cast<const float4*>(xs + j * L);
                        const float4* p4 = reinterpret_cast<const float4*>(pw + j * L);
                        if (LC > 0) {
#pragma unroll
                            for (int i = 0; i < (LC > 0 ? LC / 4 : 1); ++i) {
                                const float4 xv = x4[i], pv = p4[i];
                                acc0 = fmaf(ex2_approx(fabsf(xe - xv.x) * nb), pv.x, acc0);
                                acc1 = fmaf(ex2_approx(fabsf(xe - xv.y) * nb), pv.y, acc1);
                                acc0 = fmaf(ex2_approx(fabsf(xe - xv.z) * nb), pv.z, acc0);
                                acc1 = fmaf(ex2_approx(fabsf(xe - xv.w) * nb), pv.w, acc1);
                            }
                        } else {
                            for (int i = 0; i < (L >> 2); ++i) {
                                const float4 xv = x4[i], pv = p4[i];
                                acc0 = fmaf(ex2_approx(fabsf(xe - xv.x) * nb), pv.x, acc0);
                                acc1 = fmaf(ex2_approx(fabsf(xe - xv.y) * nb), pv.y, acc1);
                                acc0 = fmaf(ex2_approx(fabsf(xe - xv.z) * nb), pv.z, acc0);
                                acc1 = fmaf(ex2_approx(fabsf(xe - xv.w) * nb), pv.w, acc1);
                            }
                        }
                        float s = acc0 + acc1;
                        s = fmaf(bonus, regcnt[j * G + ge], s);              // regional bonus of the chunk
                        if (MODEL && use_bw) s += bwchunk[j];                // bandwagon of the chunk
                        cs[j] = s;
                    }
                    // self vote and stickiness corrections (chunk je / jp)
                    const int je = e / L;
                    float see = pw[e] + bonus;
                    if (MODEL && use_bw) see += bw[e];
                    int jp = -1; float sadd = 0.f;
                    if (MODEL && prev >= 0 && prev != e) {
                        float sp = ex2_approx(fabsf(xe - xs[prev]) * nb) * pw[prev];
                        if (reg[prev] == ge) sp += bonus;
                        if (use_bw) sp += bw[prev];
                        sadd = sp * (st1p - 1.0f);
                        jp = prev / L;
                    }
                    float S = 0.f;
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        if (j == je) cs[j] -= see;
                        if (MODEL && j == jp) cs[j] += sadd;
                        cs[j] = fmaxf(cs[j], 0.0f);
                        S += cs[j];
                    }
                    if (!(S > 0.0f)) {
                        // all scores vanished: the reference makes the row uniform INCLUDING self (model.py:447-449)
                        const int j = (int)(u * (float)C);
                        vote = j < C ? j : C - 1;
                    } else {
                        // ---------------- phase 2: locate the chunk, re-scan it -------------------
                        const float tgt = u * S;
                        float prefix = 0.f, below = 0.f;
                        int jstar = 0;
#pragma unroll
                        for (int j = 0; j < NCH; ++j) {
                            prefix += cs[j];
                            const bool le = prefix <= tgt;
                            below = le ? prefix : below;
                            jstar += le ? 1 : 0;
                        }
                        const int last_chunk = (C - 1) / L;
                        if (jstar > last_chunk) { jstar = last_chunk; }
                        float cum = below;
                        int lastpos = -1;
                        vote = -1;
                        const int c0 = jstar * L;
#pragma unroll 4
                        for (int i = 0; i < L; ++i) {
                            const int c = c0 + i;
                            const float v = c < C ? full_score32<MODEL>(e, c, xe, ge, nb, bonus, st1p, prev, use_bw, xs, reg, pw, bw) : 0.0f;
                            cum += v;
                            if (v > 0.0f) lastpos = c;
                            if (vote < 0 && v > 0.0f && cum > tgt) vote = c;
                        }
                        if (vote < 0) vote = lastpos;                  // rounding at the chunk's upper edge
                        if (vote < 0) {                                // chunk empty after rounding: nearest positive candidate
                            for (int c = c0 - 1; c >= 0 && vote < 0; --c)
                                if (full_score32<MODEL>(e, c, xe, ge, nb, bonus, st1p, prev, use_bw, xs, reg, pw, bw) > 0.0f) vote = c;
                            for (int c = c0 + L; c < C && vote < 0; ++c)
                                if (full_score32<MODEL>(e, c, xe, ge, nb, bonus, st1p, prev, use_bw, xs, reg, pw, bw) > 0.0f) vote = c;
                            if (vote < 0) vote = (e == 0 && C > 1) ? 1 : 0;
                        }
                    }
                } else {
                    // ---------------- run-off round: first k columns, renormalised ---------------
                    float sum = 0.f;
                    for (int j = 0; j < k; ++j)
                        sum += full_score32<MODEL>(e, j, xe, ge, nb, bonus, st1p, prev, use_bw, xs, reg, pw, bw);
                    int jsel;
                    if (!(sum > 0.0f)) {
                        jsel = (int)(u * (float)k);                    // simulate.py:151 uniform fallback
                    } else {
                        const float tgt = u * sum;
                        float cum = 0.f; jsel = -1; int lastpos = 0;
                        for (int j = 0; j < k; ++j) {
                            const float v = full_score32<MODEL>(e, j, xe, ge, nb, bonus, st1p, prev, use_bw, xs, reg, pw, bw);
                            cum += v;
                            if (v > 0.0f) lastpos = j;
                            if (jsel < 0 && v > 0.0f && cum > tgt) jsel = j;
                        }
                        if (jsel < 0) jsel = lastpos;
                    }
                    if (jsel >= k) jsel = k - 1;
                    vote = eff[jsel];
                }
                atomicAdd(&tally[vote], 1);
                if (MODEL) {
                    // prev votes are read by this lane only for its own electors, so in-place update is safe
                    votes[e] = vote;
                }
            }
            __syncwarp();
            // ---------------- close the round -----------------------------------------------------
            if (a.out.round_tallies)
                for (int c = lane; c < C; c += 32)
                    a.out.round_tallies[(t * (uint64_t)R + (uint64_t)(r - 1)) * (uint64_t)C + c] = tally[c];
            if (MODEL) for (int c = lane; c < C; c += 32) prev_tally[c] = tally[c];
            have_prev = true;
            int mx, arg;
            warp_argmax_first(tally, C, lane, mx, arg);
            if (mx >= need) { winner = arg; reason = CSIM_REASON_WINNER; break; }        // simulate.py:174-180
            if (r >= rr && k > 0) { warp_top_k(tally, C, k, lane, eff, mark); have_eff = true; }   // simulate.py:183-186
            __syncwarp();
        }
        // ---------------- trial outputs -----------------------------------------------------------
        if (a.out.final_votes) for (int c = lane; c < C; c += 32) a.out.final_votes[t * (uint64_t)C + c] = tally[c];
        if (a.out.round_tallies)
            for (size_t i = (size_t)rounds * C + lane; i < (size_t)R * C; i += 32) a.out.round_tallies[t * (uint64_t)R * C + i] = -1;
        if (lane == 0) {
            a.out.winner[t] = winner; a.out.rounds[t] = rounds;
            if (a.out.reason) a.out.reason[t] = (uint8_t)reason;
            if (a.out.winner_hist) atomicAdd(&a.out.winner_hist[(size_t)set * (C + 1) + (winner >= 0 ? winner : C)], 1ULL);
            if (a.out.rounds_hist && winner >= 0) atomicAdd(&a.out.rounds_hist[(size_t)set * (R + 1) + rounds], 1ULL);
        }
        acc_rounds += (unsigned long long)rounds; acc_trials += 1;
        __syncwarp();
    }
    if (cur_set >= 0 && lane == 0 && a.out.totals) {
        atomicAdd(&a.out.totals[2 * cur_set], acc_rounds); atomicAdd(&a.out.totals[2 * cur_set + 1], acc_trials);
    }
}
