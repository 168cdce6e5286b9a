// full32.cuh — fp32 engine for the `model` wiring with EVERY term of the model live: bandwagon, stickiness,
// candidate fatigue (model.py:372-379, set chosen as simulate.py:84-118) and stop-candidate (model.py:384-415).
//
// Fatigue scales whole columns by a per-trial factor and stop-candidate scales, per elector, the candidates within
// an ideological distance — neither is a sum of per-candidate constants nor a point correction, so the
// trial-invariant tables of prefix.cuh do not apply; this kernel evaluates every score of a row inline (the
// "canonical dense per-trial formulation" of SURVEY §8d, MUFU form).  It is the GENERAL kernel of the FULL engine: batches with
// fatigue / stop-candidate live, at most 255 electors, non-negative factors and a factorisable beta schedule run on the EXT
// instantiation of the dense kernel instead (dense32.cuh, 1.6x faster; eng_full.cu decides).  One warp owns one trial; a lane owns
// whole electors:
//
//   pass 1   the row's scores block by block (8 candidates, operands as 128-bit loads), running prefix per chunk
//            into the lane's column of shared memory;
//   pass 2   S = last prefix, target = u S, binary search over the <= 32 chunk prefixes, re-scan of the chunk the
//            target falls into with the SAME score routine (== searchsorted(cumsum(p)/sum, u, 'right'), simulate.py:153).
//
// The fatigue set and the stop-candidate trigger are decided in fp64 exactly as the exact engine decides them
// (exact64.cuh), so both engines fatigue / boost the same candidates; only the scores themselves are fp32.
// Both clamps of the model (max(s, 0) after stickiness and at the end, model.py:369,427) are applied, so negative
// bonus / bandwagon / papabile / penalty / boost factors — which the table-based engines refuse — are taken too: this is
// the fp32 engine of last resort.
#pragma once
#include "common.cuh"
#include "prefix.cuh"      // prefix_slot: the same dealing of electors to lanes

#ifndef CSIM_FULL32_THREADS
#define CSIM_FULL32_THREADS 896        // 28 warps per CTA, 1 CTA per SM (72 registers): 512 / 768 / 896 / 1024 threads = 29.6 / 26.6 / 25.5 / 25.9 ms
#endif

struct Full32Args {
    RosterDev ro;
    BatchDev b;
    TrialOut out;
    const float* nbeta;     // [sets][R]  -beta_r * log2(e)
    const uint32_t* accmask;// [sets][E][accw]: bit c of row e <=> candidate c is ACCEPTABLE to elector e for that set's stop-candidate
                            // distance: !(|x_e - x_c| > D) decided in fp64 exactly as model.py:390 (k_full_accmask, once per batch);
                            // null when no set has stop-candidate enabled.  The hot loop tests bits instead of fp64 distances.
    int accw;               // words per row
    int nblk;               // blocks of 8 candidates, 8 nblk >= C
    int Lb;                 // blocks per chunk
    int nch;                // chunks, nch * Lb >= nblk, nch <= 32
    int e_quad_end, rows;   // lane dealing (prefix_slot)
    int kmax, Fmax;
    int warps_per_cta;
    WorkQueue wq;           // dynamic trial scheduling (common.cuh)
    size_t cta_smem, warp_smem;
    // byte offsets (host-computed, full32_layout)
    uint32_t o_xs, o_pw, o_rb, o_xg, o_nbs;                                                  // CTA part
    uint32_t w_bw, w_phi, w_col, w_tally, w_eff, w_votes, w_hist, w_avg, w_fat, w_mark, w_thr; // warp part
    // PROBE instantiation only (csim_engine_probs): see Dense32Args; probe_fatigued = the fatigue mask to apply (or null)
    int probe_round;
    const int32_t* probe_prev_vote;
    const int32_t* probe_prev_count;
    const uint8_t* probe_fatigued;
    float* probe_scores;
    float* probe_prefix;
    int probe_cap;
};

__host__ inline bool full32_plan(Full32Args& a, int E, int G, int R, int kmax, int Fmax, size_t budget) {
    a.ro.E = E; a.ro.G = G; a.kmax = kmax; a.Fmax = Fmax < 1 ? 1 : Fmax;
    a.nblk = (E + 7) >> 3;
    a.Lb = (a.nblk + 31) / 32;
    a.nch = (a.nblk + a.Lb - 1) / a.Lb;
    const int n_left = E & 127;
    a.e_quad_end = (E & ~127) + (n_left > 64 ? 128 : 0);
    a.rows = E > a.e_quad_end ? E : a.e_quad_end;
    const uint32_t CP = 8u * (uint32_t)a.nblk, Ei = ((uint32_t)E + 7u) & ~7u, rows8 = ((uint32_t)a.rows + 7u) & ~7u;
    uint32_t o = 0;
    a.o_xs = o; o += 4u * CP;
    a.o_pw = o; o += 4u * CP;
    a.o_rb = o; o += 4u * CP * (uint32_t)G;
    a.o_xg = o; o += 8u * rows8;
    a.o_nbs = o; o += 4u * (uint32_t)(R > 0 ? R : 1);
    a.cta_smem = ((size_t)o + 15) & ~(size_t)15;
    uint32_t w = 0;
    a.w_bw = w; w += 4u * CP;
    a.w_phi = w; w += 4u * CP;
    a.w_col = w; w += 4u * 32u * (uint32_t)a.nch;
    a.w_avg = w; w += 8u * Ei;
    a.w_tally = w; w += 4u * Ei;
    a.w_hist = w; w += 4u * Ei * (uint32_t)a.Fmax;
    a.w_eff = w; w += 4u * (((uint32_t)kmax + 3u) & ~3u);
    a.w_thr = w; w += 4u * 32u;
    a.w_votes = w; w += 2u * rows8;
    a.w_fat = w; w += Ei;
    a.w_mark = w; w += Ei;
    a.warp_smem = ((size_t)w + 15) & ~(size_t)15;
    if (a.cta_smem + a.warp_smem > budget) return false;
    const size_t fit = (budget - a.cta_smem) / a.warp_smem;
    const size_t mx = CSIM_FULL32_THREADS / 32;
    a.warps_per_cta = (int)(fit < mx ? fit : mx);
    return true;
}

// acceptable(e, c) for every set with stop-candidate enabled (model.py:384-415): grid = (ceil(E * W / 128), sets), block = 128
static __global__ void k_full_accmask(RosterDev ro, const DevSet* sets, int W, uint32_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, set = blockIdx.y;
    const int E = ro.E;
    if (i >= E * W) return;
    const int e = i / W, wd = i - e * W;
    const double xe = ro.x64[e], D = sets[set].stop_D;
    uint32_t m = 0;
    for (int b = 0; b < 32; ++b) {
        const int c = 32 * wd + b;
        if (c < E && !(fabs(xe - ro.x64[c]) > D)) m |= 1u << b;
    }
    out[((size_t)set * E + e) * W + wd] = m;
}

struct Full32Ctx {          // what one score needs (per warp / per round)
    const float* xs; const float* pw; const float* rb; const float* bw; const float* phi;
    float nb, st1p, boost;
    int C;
};

// The 8 scores of block `blk` of elector e's row, every term inline in the order of model.py:311-427.
//   RAW  pass 1: stickiness and the self vote are NOT applied per candidate (two compares per pair saved); the caller fixes the
//        running prefix at the chunks of prev and of e with the exact differences (full32_score1 with / without them)
//   EXT  candidate fatigue / stop-candidate live for this batch;  NEG  some factor is negative: the model's clamps matter
template <bool RAW, bool EXT, bool NEG>
__device__ __forceinline__ void full32_score8(const Full32Ctx& k, int blk, int e, float xe, const uint32_t* arow, const float* rbrow, int prev, bool trig,
                                              float (&v)[8]) {
    const int cb = blk << 3;
    const float4 xa = *reinterpret_cast<const float4*>(k.xs + cb), xb = *reinterpret_cast<const float4*>(k.xs + cb + 4);
    const float4 pa = *reinterpret_cast<const float4*>(k.pw + cb), pb = *reinterpret_cast<const float4*>(k.pw + cb + 4);
    const float4 ra = *reinterpret_cast<const float4*>(rbrow + cb), rb = *reinterpret_cast<const float4*>(rbrow + cb + 4);
    const float4 ba = *reinterpret_cast<const float4*>(k.bw + cb), bb = *reinterpret_cast<const float4*>(k.bw + cb + 4);
    const float x8[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
    const float p8[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
    const float r8[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
    const float b8[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
    const int is = e - cb, ip = prev - cb;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float s = fmaf(ex2_approx(fabsf(xe - x8[i]) * k.nb), p8[i], r8[i]) + b8[i];     // exp, papabile, regional, bandwagon
        if (!RAW && ip == i) s *= k.st1p;                                                 // stickiness
        if (NEG) s = fmaxf(s, 0.0f);                                                      // clamp (model.py:369)
        v[i] = s;
    }
    if (EXT) {
        const float4 fa = *reinterpret_cast<const float4*>(k.phi + cb), fb = *reinterpret_cast<const float4*>(k.phi + cb + 4);
        const float f8[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= f8[i];                                        // fatigue
        if (trig) {                                                                       // stop-candidate: boost the acceptable ones
            const uint32_t m8 = __ldg(arow + (cb >> 5)) >> (cb & 31);          // this block's 8 "acceptable" bits (model.py:390, k_full_accmask)
#pragma unroll
            for (int i = 0; i < 8; ++i) if ((m8 >> i) & 1u) v[i] *= k.boost;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (!RAW && is == i) v[i] = 0.0f;                                                 // no self vote
        else if (NEG) v[i] = fmaxf(v[i], 0.0f);                                           // final clamp (model.py:427)
    }
}

// one score; `sticky` / `self` say whether those two point rules are applied (pass 1 needs the difference they make)
template <bool EXT, bool NEG>
__device__ __forceinline__ float full32_score1(const Full32Ctx& k, int c, int e, float xe, const uint32_t* arow, const float* rbrow, int prev, bool trig,
                                               bool sticky = true, bool self = true) {
    float s = fmaf(ex2_approx(fabsf(xe - k.xs[c]) * k.nb), k.pw[c], rbrow[c]) + k.bw[c];
    if (sticky && c == prev) s *= k.st1p;
    if (NEG) s = fmaxf(s, 0.0f);
    if (EXT) {
        s *= k.phi[c];
        if (trig && ((__ldg(arow + (c >> 5)) >> (c & 31)) & 1u)) s *= k.boost;
    }
    if (self && c == e) return 0.0f;
    return NEG ? fmaxf(s, 0.0f) : s;
}

template <bool EXT, bool NEG, bool PROBE = false>
__global__ void __launch_bounds__(CSIM_FULL32_THREADS, 1) k_trials_full32(const __grid_constant__ Full32Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    asm volatile("" : "+r"(lane));       // opaque: kept in a register instead of being re-derived from %tid inside the loops
    const int E = a.ro.E, C = E, R = a.b.R, G = a.ro.G;
    const int nblk = a.nblk, CP = nblk << 3, Lb = a.Lb, nch = a.nch;
    float* xs = (float*)(smem_raw + a.o_xs);
    float* pwv = (float*)(smem_raw + a.o_pw);
    float* rbv = (float*)(smem_raw + a.o_rb);
    int2* xg = (int2*)(smem_raw + a.o_xg);
    float* nbs = (float*)(smem_raw + a.o_nbs);
    for (int i = threadIdx.x; i < CP; i += blockDim.x) { xs[i] = i < C ? a.ro.x32[i] : 0.0f; pwv[i] = 0.0f; }
    for (int i = threadIdx.x; i < CP * G; i += blockDim.x) rbv[i] = 0.0f;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        xg[prefix_slot(c, a.e_quad_end)] = make_int2(__float_as_int(a.ro.x32[c]), a.ro.region[c]);
    unsigned char* base = smem_raw + a.cta_smem + (size_t)wib * a.warp_smem;
    float* bw = (float*)(base + a.w_bw);
    float* phi = (float*)(base + a.w_phi);
    float* col = (float*)(base + a.w_col) + lane;                 // this lane's column: col[32 j]
    double* avg = (double*)(base + a.w_avg);
    int32_t* tally = (int32_t*)(base + a.w_tally);
    int32_t* hist = (int32_t*)(base + a.w_hist);
    int32_t* eff = (int32_t*)(base + a.w_eff);
    int32_t* thr = (int32_t*)(base + a.w_thr);
    uint16_t* votes = (uint16_t*)(base + a.w_votes);
    uint8_t* fat = base + a.w_fat;
    uint8_t* mark = base + a.w_mark;

    Full32Ctx kx;
    kx.xs = xs; kx.pw = pwv; kx.rb = rbv; kx.bw = bw; kx.phi = phi;
    kx.nb = 0.f; kx.st1p = 1.f; kx.boost = 1.f; kx.C = C;
    int step0 = 1;
    while (step0 * 2 <= nch) step0 *= 2;

    auto for_each_elector = [&](int r, uint32_t s_lo, uint32_t s_hi, auto&& fn) {
        deal_electors_once(lane, E, a.e_quad_end, r, s_lo, s_hi, a.b.seed_lo, a.b.seed_hi, fn);
    };

    int cur_set = -1;
    unsigned long long acc_rounds = 0, acc_trials = 0;
    DevSet S;
    memset(&S, 0, sizeof S);

    __shared__ int s_skip[2];
    for (int si = 0; si < a.b.n_sets; ++si) {
        const int set = cta_set_of_step(si, a.b.n_sets);         // uniform over the CTA
        if (threadIdx.x == 0) s_skip[si & 1] = set_exhausted(a.wq, set, a.b.n_sims) ? 1 : 0;
        __syncthreads();
        if (s_skip[si & 1]) continue;
        {
            if (cur_set >= 0 && lane == 0 && a.out.totals && acc_trials) {
                atomicAdd(&a.out.totals[2 * cur_set], acc_rounds); atomicAdd(&a.out.totals[2 * cur_set + 1], acc_trials);
            }
            acc_rounds = acc_trials = 0;
            cur_set = set;
            S = a.b.sets[set];
            kx.st1p = S.f_st1p; kx.boost = S.f_stop_boost;
            __syncthreads();                                    // readers of the previous set's tables are done
            for (int r = threadIdx.x; r < R; r += blockDim.x) nbs[r] = a.nbeta[(size_t)set * R + r];
            for (int c = threadIdx.x; c < C; c += blockDim.x) {
                pwv[c] = a.ro.pap[c] ? S.f_pwf : 1.0f;
                rbv[a.ro.region[c] * CP + c] = S.f_bonus;
            }
            __syncthreads();
        }
        const int k = S.k, need = S.need, rr = S.rr, F = S.fat_rounds;
        for (uint64_t i0 = warp_claim_trials(a.wq, set, lane); i0 < a.b.n_sims; i0 = warp_claim_trials(a.wq, set, lane))
        for (int ti = 0; ti < a.wq.chunk; ++ti) {
            const uint64_t i = i0 + (uint64_t)ti;
            if (i >= a.b.n_sims) break;
            const uint64_t t = (uint64_t)set * a.b.n_sims + i;
            const uint64_t sim = a.b.sim_id_begin + i;
            const uint32_t s_lo = (uint32_t)sim, s_hi = (uint32_t)(sim >> 32);
            bool have_eff = false, have_prev = false;
            int winner = -1, rounds = 0, reason = CSIM_REASON_MAX_ROUNDS, n_hist = 0, prev_mx = 0;
            for (int c = lane; c < CP; c += 32) { bw[c] = 0.0f; phi[c] = c < C ? 1.0f : 0.0f; }
            if (R == 0) for (int c = lane; c < C; c += 32) tally[c] = 0;
            __syncwarp();
            if (PROBE && (a.probe_prev_vote || a.probe_prev_count)) {          // start from the given previous-round state
                int m = 0;
                for (int c = lane; c < C; c += 32) {
                    const int v = a.probe_prev_count ? a.probe_prev_count[c] : 0;
                    tally[c] = v; hist[c] = v; m = max(m, v);
                    votes[prefix_slot(c, a.e_quad_end)] = (uint16_t)(a.probe_prev_vote ? a.probe_prev_vote[c] : 0xFFFF);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(CSIM_FULL, m, off));
                prev_mx = m; have_prev = true; n_hist = 1;
                __syncwarp();
            }

            for (int r = PROBE ? a.probe_round : 1; r <= R; ++r) {
                rounds = r;
                kx.nb = nbs[r - 1];
                // ---- per-round state from round r-1 (tally still holds it) ---------------------------------
                // fatigue set, simulate.py:84-118, decided in fp64 exactly as exact64.cuh decides it
                bool use_fat = false;
                if (S.en_fat) {
                    for (int c = lane; c < C; c += 32) { fat[c] = 0; mark[c] = 0; }
                    __syncwarp();
                    if (r > F) {
                        const int nh = n_hist < F ? n_hist : F;
                        if (S.fat_top_n > 0 && n_hist > 0 && nh > 0) {
                            for (int c = lane; c < C; c += 32) {
                                double acc = __ddiv_rn((double)hist[(size_t)((n_hist - nh) % a.Fmax) * C + c], (double)E);
                                for (int q = 1; q < nh; ++q)
                                    acc = __dadd_rn(acc, __ddiv_rn((double)hist[(size_t)((n_hist - nh + q) % a.Fmax) * C + c], (double)E));
                                avg[c] = __ddiv_rn(acc, (double)nh);
                            }
                            __syncwarp();
                            const int n_imm = S.fat_top_n < C ? S.fat_top_n : C;
                            for (int j = 0; j < n_imm; ++j) {                  // top-n by mean share; ties -> highest index
                                double bv = -1.0; int ba = -1;
                                for (int c = lane; c < C; c += 32)
                                    if (!mark[c] && (ba < 0 || avg[c] >= bv)) { bv = avg[c]; ba = c; }
#pragma unroll
                                for (int off = 16; off > 0; off >>= 1) {
                                    const double ov = __shfl_xor_sync(CSIM_FULL, bv, off);
                                    const int oa = __shfl_xor_sync(CSIM_FULL, ba, off);
                                    if (oa >= 0 && (ba < 0 || ov > bv || (ov == bv && oa > ba))) { bv = ov; ba = oa; }
                                }
                                if (lane == 0 && ba >= 0) mark[ba] = 1;
                                __syncwarp();
                            }
                        }
                        if (n_hist >= F) {
                            for (int c = lane; c < C; c += 32) {
                                if (mark[c]) continue;
                                bool low = true;
                                for (int q = 0; q < F && low; ++q)                  // share < threshold, as an integer test (DevSet::fat_cnt)
                                    low = hist[(size_t)((n_hist - F + q) % a.Fmax) * C + c] < S.fat_cnt;
                                fat[c] = low ? 1 : 0;
                            }
                            use_fat = true;
                        }
                        __syncwarp();
                    }
                    for (int c = lane; c < C; c += 32) phi[c] = (use_fat && fat[c]) ? S.f_fat_pen : 1.0f;
                    if (PROBE && a.probe_fatigued) for (int c = lane; c < C; c += 32) phi[c] = a.probe_fatigued[c] ? S.f_fat_pen : 1.0f;
                }
                // stop-candidate threats, model.py:384-400: candidates whose last share exceeds the threshold
                int n_thr = 0;
                // round 1 has no previous shares: the reference tests 0.0 > threshold against every candidate (simulate.py:73,
                // model.py:397), so with a NEGATIVE threat share every candidate is a threat already then
                const bool all_thr = S.en_stop && !have_prev && 0.0 > S.stop_T;
                if (all_thr) n_thr = 33;                                       // more than the list holds: the per-candidate branch below
                if (S.en_stop && have_prev) {
                    for (int c0 = 0; c0 < C; c0 += 32) {
                        const int c = c0 + lane;
                        const bool is_thr = c < C && tally[c] >= S.stop_cnt;       // share > threshold (DevSet::stop_cnt)
                        const unsigned m = __ballot_sync(CSIM_FULL, is_thr);
                        if (is_thr) { const int pos = n_thr + __popc(m & ((1u << lane) - 1u)); if (pos < 32) thr[pos] = c; }
                        n_thr += __popc(m);
                    }
                }
                // bandwagon, model.py:326-354
                if (S.has_bw && have_prev && prev_mx > 0) {
                    const float inv = 1.0f / (float)prev_mx;
                    for (int c = lane; c < C; c += 32) bw[c] = (float)tally[c] * inv * S.f_bandwagon;
                }
                __syncwarp();
                for (int c = lane; c < C; c += 32) tally[c] = 0;
                __syncwarp();
                const bool sticky_live = have_prev && S.has_sticky;

                for_each_elector(r, s_lo, s_hi, [&](int e, float u, int slot) {
                    const int2 g2 = xg[slot];
                    const float xe = __int_as_float(g2.x);
                    const float* rbrow = rbv + g2.y * CP;
                    int prev = sticky_live ? (int)votes[slot] : -1;
                    if (PROBE && prev == 0xFFFF) prev = -1;              // probe input: this elector had no previous vote
                    const uint32_t* arow = a.accmask + ((size_t)set * E + e) * a.accw;      // dereferenced only when stop-candidate is live
                    bool trig = false;
                    if (n_thr > 0) {
                        if (n_thr <= 32) {
                            for (int q = 0; q < n_thr && !trig; ++q) { const int c = thr[q]; trig = !((__ldg(arow + (c >> 5)) >> (c & 31)) & 1u); }   // an UNacceptable threat
                        } else {                                     // more threats than the list holds: test every candidate
                            for (int q = 0; q < C && !trig; ++q)
                                trig = !((__ldg(arow + (q >> 5)) >> (q & 31)) & 1u) && (all_thr || hist[(size_t)((n_hist - 1) % a.Fmax) * C + q] >= S.stop_cnt);
                        }
                    }
                    int vote;
                    if (!have_eff) {
                        // ---- pass 1: the whole row, running prefix per chunk into this lane's column ----------
                        // the point rules enter the prefix at the chunks of e and of prev, as exact differences
                        const int je = (e >> 3) / Lb;
                        const float d_self = -full32_score1<EXT, NEG>(kx, e, e, xe, arow, rbrow, -1, trig, false, false);
                        int jp = -1;
                        float d_prev = 0.0f;
                        if (prev >= 0 && prev != e) {
                            jp = (prev >> 3) / Lb;
                            d_prev = full32_score1<EXT, NEG>(kx, prev, e, xe, arow, rbrow, prev, trig, true, false) -
                                     full32_score1<EXT, NEG>(kx, prev, e, xe, arow, rbrow, prev, trig, false, false);
                        }
                        float run = 0.0f;
                        for (int j = 0; j < nch; ++j) {
                            for (int b = 0; b < Lb; ++b) {
                                const int blk = j * Lb + b;
                                if (blk >= nblk) break;
                                float v[8];
                                full32_score8<true, EXT, NEG>(kx, blk, e, xe, arow, rbrow, prev, trig, v);
                                run += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
                            }
                            if (j == je) run += d_self;
                            if (j == jp) run += d_prev;
                            col[32 * j] = run;
                        }
                        if (PROBE) {                                  // csim_engine_probs: the numbers this draw is made from
                            for (int j = 0; j < nch; ++j) a.probe_prefix[(size_t)e * a.probe_cap + j] = col[32 * j];
                            for (int blk = 0; blk < nblk; ++blk) {
                                float v[8];
                                full32_score8<false, EXT, NEG>(kx, blk, e, xe, arow, rbrow, prev, trig, v);
#pragma unroll
                                for (int q = 0; q < 8; ++q) if (8 * blk + q < C) a.probe_scores[(size_t)e * C + 8 * blk + q] = v[q];
                            }
                        }
                        const float Sm = run;
                        if (!(Sm > 0.0f)) {                           // model.py:447-449: uniform row INCLUDING self
                            const int j = (int)(u * (float)C);
                            vote = j < C ? j : C - 1;
                        } else {
                            // ---- pass 2: chunk search + re-scan ---------------------------------------------
                            const float tgt = u * Sm;
                            int pos = 0;               // bound-free search (table.cuh): windows [0, P) / [nch - P, nch), P = step0
                            float cum = 0.0f;
                            if (nch > step0) {
                                const float f = col[32 * (nch - step0 - 1)];
                                const bool le = f <= tgt;
                                pos = le ? nch - step0 : 0;
                                cum = le ? f : 0.0f;
                            }
                            for (int st = step0 >> 1; st > 0; st >>= 1) {
                                const float f = col[32 * (pos + st - 1)];
                                const bool le = f <= tgt;
                                pos = le ? pos + st : pos;
                                cum = le ? f : cum;
                            }
                            int cnt = 0;
                            const int blk0 = pos * Lb;
                            for (int b = 0; b < Lb; ++b) {
                                const int blk = blk0 + b;
                                if (blk >= nblk) break;
                                float v[8];
                                full32_score8<false, EXT, NEG>(kx, blk, e, xe, arow, rbrow, prev, trig, v);
#pragma unroll
                                for (int q = 0; q < 8; ++q) { cum += v[q]; cnt += (cum <= tgt) ? 1 : 0; }
                                if (cnt < 8 * (b + 1)) break;
                            }
                            vote = (blk0 << 3) + cnt;
                            const int maxc = min(((blk0 + Lb) << 3) - 1, C - 1);
                            vote = vote > maxc ? maxc : vote;
                            if (vote == e) vote = e > 0 ? e - 1 : (C > 1 ? 1 : 0);
                        }
                    } else {
                        // ---- run-off round: first k columns, renormalised (simulate.py:144-154) --------------
                        float sum = 0.f;
                        for (int j = 0; j < k; ++j) sum += full32_score1<EXT, NEG>(kx, j, e, xe, arow, rbrow, prev, trig);
                        int jsel;
                        if (!(sum > 0.0f)) {
                            jsel = (int)(u * (float)k);
                        } else {
                            const float tgt = u * sum;
                            float cum = 0.f; jsel = -1; int lastpos = 0;
                            for (int j = 0; j < k; ++j) {
                                const float v = full32_score1<EXT, NEG>(kx, j, e, xe, arow, rbrow, prev, trig);
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
                    votes[slot] = (uint16_t)vote;                    // read back only by this lane: in place is safe
                });
                __syncwarp();
                if (PROBE) break;                                    // one round is all a probe plays
                // ---- close the round ---------------------------------------------------------------------
                if (a.out.round_tallies)
                    for (int c = lane; c < C; c += 32)
                        a.out.round_tallies[(t * (uint64_t)R + (uint64_t)(r - 1)) * (uint64_t)C + c] = tally[c];
                for (int c = lane; c < C; c += 32) hist[(size_t)(n_hist % a.Fmax) * C + c] = tally[c];
                have_prev = true; ++n_hist;
                __syncwarp();
                int mx, arg;
                warp_argmax_first(tally, C, lane, mx, arg);
                prev_mx = mx;
                if (mx >= need) { winner = arg; reason = CSIM_REASON_WINNER; break; }       // simulate.py:174-180
                if (r >= rr && k > 0) { warp_top_k(tally, C, k, lane, eff, mark); have_eff = true; }   // simulate.py:183-186
                __syncwarp();
            }
            write_trial_outputs(a.out, t, set, C, R, lane, tally, winner, rounds, reason);
            acc_rounds += (unsigned long long)rounds; acc_trials += 1;
            __syncwarp();
        }
    }
    if (cur_set >= 0 && lane == 0 && a.out.totals && acc_trials) {
        atomicAdd(&a.out.totals[2 * cur_set], acc_rounds); atomicAdd(&a.out.totals[2 * cur_set + 1], acc_trials);
    }
}
