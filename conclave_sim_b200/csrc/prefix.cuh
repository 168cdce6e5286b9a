// prefix.cuh — the PREFIX engine (CSIM_ENGINE_PREFIX): trial-invariant chunk-prefix tables + an
// in-chunk re-scan.  Both wirings, fp32.
//
// Every term of a score that depends on the TRIAL is a sum of per-candidate constants or a point
// correction (model.py:311-424):
//
//   s[e,c] = exp(-beta_r |x_e - x_c|) pw_c + bonus [g_e == g_c]        "base": roster, parameters, round
//          + bw_c                                                       bandwagon: the same for every elector
//          then  s[e,prev_e] *= 1 + st   and   s[e,e] = 0               two point corrections
//
// so the prefix of row e over the first j chunks of L candidates is
//
//   F_e(j) = T_r[e][j]  +  ( BWP[j] - [j >= chunk(e)] bw_e )  +  [j >= chunk(prev_e)] st (base[e,prev_e] + bw_prev)
//
// with T_r[e][j] = sum of base[e,c] over c < (j+1) L, c != e — built ONCE per (parameter set, round) in
// fp64 and rounded once (k_prefix_tables32) — and BWP the per-trial-round prefix of bw over chunks
// (17 numbers at E = 134).  A draw is then
//   1. S = F_e(last), target = u S, a binary search over the <= 32 chunk prefixes (<= 5 probes), and
//   2. a re-scan of the ONE chunk the target falls into, every term inline in roster order, to find the
//      first c with cum > target  (== searchsorted(cumsum(p)/sum, u, 'right'), simulate.py:153).
// That is the dense fp32 engine (dense32.cuh) with its phase 1 — the E x C exponentials every trial
// re-evaluates — replaced by a table lookup: O(log C + L) work per elector-round instead of O(C).  Under
// the ref_cli wiring the bandwagon / stickiness terms vanish and F_e(j) is one load.
//
// The tables of ALL rounds of one parameter set are small (E x 17 floats per round, 182 KB for 20 rounds
// at E = 134) and live in shared memory (1 CTA per SM, cp.async-staged once per set), so — unlike the
// full-resolution CDF tables of table.cuh, which only fit one round at a time — trials stay independent:
// one warp owns one trial from round 1 to its end, then takes the next.  When they do not fit (cfg-5:
// 1024 electors x 50 rounds = 6.7 MB) the probes read them through L2 instead (STAGE = false).
//
// Table rows are stored in the order the lanes visit electors (prefix_slot) with an odd row pitch, so the
// 32 lanes of a warp probing the same chunk column hit 32 different banks.
#pragma once
#include "common.cuh"

#ifndef CSIM_PREFIX_THREADS
#define CSIM_PREFIX_THREADS 768        // 24 warps per CTA, 1 CTA per SM (<= 85 registers)
#endif

struct PrefixArgs {
    RosterDev ro;
    BatchDev b;
    TrialOut out;
    const float* nbeta;     // [sets][R]  -beta_r * log2(e)
    const float* T;         // [sets][R][rows][stride]  chunk-prefix tables
    int Ls;                 // log2(candidates per chunk); L = 1 << Ls >= 8
    int nch;                // chunks, nch * L >= C, nch <= 32
    int stride;             // row pitch in floats (odd, >= nch)
    int rows;               // rows per round
    int e_quad_end;         // electors below this are dealt as whole Philox quads per lane, the rest one per lane
    int kmax;
    int warps_per_cta;
    int tpi;                // trials per warp per work item
    size_t set_pitch;       // floats between the tables of consecutive sets (16-byte multiple)
    size_t tab_smem;        // bytes of the staged tables (0 when STAGE = false)
    size_t cta_smem;        // staged tables + roster tables
    size_t warp_smem;       // per-warp state
};

// Row of elector e inside one round's table: quad-dealt electors (e = 4 (lane + 32 qi) + i) sit at
// 128 qi + 32 i + lane, so one pass of the warp (fixed i) touches 32 consecutive rows.
__host__ __device__ inline int prefix_slot(int e, int e_quad_end) {
    return e < e_quad_end ? ((e >> 7) << 7) + ((e & 3) << 5) + ((e >> 2) & 31) : e;
}
__host__ __device__ inline size_t prefix_roster_smem(int CP, int G, int R) {
    const size_t nblk = (size_t)CP >> 3;
    return ((size_t)CP * 8 /*xs, reg*/ + nblk * G * 4 /*regmask*/ + nblk * 4 /*papmask*/ + (size_t)R * 4 /*nbeta*/ + 63) & ~(size_t)15;
}
__host__ __device__ inline size_t prefix_warp_smem(int E, int CP, int kmax, bool model) {
    const size_t Ei = ((size_t)E + 7) & ~(size_t)7;
    size_t n = Ei * 4 /*tally*/ + (((size_t)kmax + 3) & ~(size_t)3) * 4 /*eff*/ + Ei /*mark*/;
    if (model) n += (size_t)CP * 4 /*bw*/ + 32 * 4 /*bwp*/ + Ei * 2 /*votes*/;
    return (n + 15) & ~(size_t)15;
}

// K0: chunk-prefix tables.  One warp per (set, round, elector), lane j sums chunk j in fp64 (the order of the
// terms is model.py:311-324: exp, papabile factor, regional bonus), an inclusive warp scan makes the prefixes.
// grid = (ceil(E / 8), sets * R), block = 256.
__global__ void k_prefix_tables32(RosterDev ro, const DevSet* sets, const double* betas, int R, int Ls, int nch, int stride,
                                  int rows, int e_quad_end, size_t set_pitch, float* T) {
    const int lane = threadIdx.x & 31, e = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int sr = blockIdx.y;                       // set * R + (round - 1)
    const int C = ro.E;
    if (e >= C) return;
    const DevSet* S = &sets[sr / R];
    const double beta = betas[sr], xe = ro.x64[e], pwf = S->pwf, bonus = S->bonus;
    const int ge = ro.region[e];
    double s = 0.0;
    if (lane < nch) {
        const int c1 = min(C, (lane + 1) << Ls);
        for (int c = lane << Ls; c < c1; ++c) {
            if (c == e) continue;                                            // model.py:418-424
            double v = exp(-beta * fabs(xe - ro.x64[c]));
            if (ro.pap[c]) v *= pwf;
            if (ro.region[c] == ge) v += bonus;
            s += fmax(v, 0.0);
        }
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double o = __shfl_up_sync(CSIM_FULL, s, off);
        if (lane >= off) s += o;
    }
    const double last = __shfl_sync(CSIM_FULL, s, nch - 1);
    float* row = T + (size_t)(sr / R) * set_pitch + ((size_t)(sr % R) * rows + prefix_slot(e, e_quad_end)) * stride;
    if (lane < stride) row[lane] = (float)(lane < nch ? s : last);
    if (lane == 0 && stride > 32) row[32] = (float)last;
}

struct PrefixCtx {              // per-warp / per-round constants of a draw
    const float* xs; const int32_t* reg; const uint32_t* regmask; const uint32_t* papmask;
    const float* bw; const float* bwp;
    float nb, bonus, st1p, pwf;
    int C, G, Ls, nch, step0;
};

__device__ __forceinline__ float prefix_pw(const PrefixCtx& k, int c) {
    return ((k.papmask[c >> 3] >> (c & 7)) & 1u) ? k.pwf : 1.0f;
}

// every term of one score, inline (run-off rounds)
template <bool MODEL>
__device__ __forceinline__ float prefix_full_score(const PrefixCtx& k, int e, int c, float xe, int ge, int prev) {
    float v = ex2_approx(fabsf(xe - k.xs[c]) * k.nb) * prefix_pw(k, c);
    if (k.reg[c] == ge) v += k.bonus;
    if (MODEL) {
        v += k.bw[c];
        if (c == prev) v *= k.st1p;
    }
    return c == e ? 0.0f : v;
}

// One vote of a full-field round.  Trow = this elector's row of the round's chunk-prefix table.
template <bool LC8, bool MODEL>
__device__ __forceinline__ int prefix_draw(const PrefixCtx& k, const float* __restrict__ Trow, int e, float u, int prev) {
    const int C = k.C, Ls = LC8 ? 3 : k.Ls, nch = k.nch;
    const float xe = k.xs[e];
    const int ge = k.reg[e];
    const int je = e >> Ls;
    float selfbw = 0.f, sadd = 0.f;
    int jp = 0x7fffffff;
    if (MODEL) {
        selfbw = k.bw[e];                                   // bwp counts e's own bandwagon term; its score is 0 (model.py:418-424)
        if (prev >= 0 && prev != e) {                       // stickiness, model.py:357-367
            float b = ex2_approx(fabsf(xe - k.xs[prev]) * k.nb) * prefix_pw(k, prev);
            if (k.reg[prev] == ge) b += k.bonus;
            b += k.bw[prev];
            sadd = b * (k.st1p - 1.0f);
            jp = prev >> Ls;
        }
    }
    auto F = [&](int j) -> float {                          // prefix of the row over chunks 0..j
        float v = Trow[j];
        if (MODEL) {
            v += k.bwp[j];
            v -= (j >= je) ? selfbw : 0.0f;
            v += (j >= jp) ? sadd : 0.0f;
        }
        return v;
    };
    const float S = F(nch - 1);
    if (!(S > 0.0f)) {            // all scores vanished: the reference makes the row uniform INCLUDING self (model.py:447-449)
        const int j = (int)(u * (float)C);
        return j < C ? j : C - 1;
    }
    const float tgt = u * S;
    int pos = 0;                  // number of chunks whose prefix is <= tgt
    for (int step = k.step0; step > 0; step >>= 1) {
        const int np = pos + step;
        if (np <= nch && F(np - 1) <= tgt) pos = np;
    }
    const int jstar = pos > nch - 1 ? nch - 1 : pos;
    float cum = jstar > 0 ? F(jstar - 1) : 0.0f;
    const int c0 = jstar << Ls;
    int cnt = 0;
    const int nb8 = LC8 ? 1 : (1 << (Ls - 3));
    const float nb = k.nb;
    for (int bi = 0; bi < nb8; ++bi) {                      // blocks of 8 candidates of chunk jstar, in roster order
        const int blk = (c0 >> 3) + bi, cb = blk << 3;
        const float4 xa = *reinterpret_cast<const float4*>(k.xs + cb), xb = *reinterpret_cast<const float4*>(k.xs + cb + 4);
        const uint32_t rm = k.regmask[blk * k.G + ge];       // bit i: candidate cb+i is of e's region
        const uint32_t pz = k.papmask[blk];                  // bits 0-7: papabile; bits 8-15: index >= C (padding)
        const uint32_t zm = (pz >> 8) | ((blk == (e >> 3)) ? (1u << (e & 7)) : 0u);   // bit i: score is zero (self / padding)
        uint32_t sm = 0;
        float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
        if (MODEL) {
            if (prev >= 0 && (prev >> 3) == blk) sm = 1u << (prev & 7);
            ba = *reinterpret_cast<const float4*>(k.bw + cb); bb = *reinterpret_cast<const float4*>(k.bw + cb + 4);
        }
#define CSIM_PRESCAN(i, xv, bv)                                                          \
        {                                                                                \
            const float m = ex2_approx(fabsf(xe - (xv)) * nb);                           \
            float v = fmaf(m, ((pz >> (i)) & 1u) ? k.pwf : 1.0f, ((rm >> (i)) & 1u) ? k.bonus : 0.0f); \
            if (MODEL) { v += (bv); if ((sm >> (i)) & 1u) v *= k.st1p; }                 \
            if ((zm >> (i)) & 1u) v = 0.0f;                                              \
            cum += v;                                                                    \
            cnt += (cum <= tgt) ? 1 : 0;                                                 \
        }
        CSIM_PRESCAN(0, xa.x, ba.x) CSIM_PRESCAN(1, xa.y, ba.y) CSIM_PRESCAN(2, xa.z, ba.z) CSIM_PRESCAN(3, xa.w, ba.w)
        CSIM_PRESCAN(4, xb.x, bb.x) CSIM_PRESCAN(5, xb.y, bb.y) CSIM_PRESCAN(6, xb.z, bb.z) CSIM_PRESCAN(7, xb.w, bb.w)
#undef CSIM_PRESCAN
        if (cnt < 8 * (bi + 1)) break;                       // the crossing is inside this block
    }
    int vote = c0 + cnt;                                     // == first c with cum > tgt
    const int maxc = min(c0 + (1 << Ls) - 1, C - 1);
    vote = vote > maxc ? maxc : vote;                        // rounding at the chunk's upper edge
    if (vote == e) vote = e > 0 ? e - 1 : (C > 1 ? 1 : 0);
    return vote;
}

template <bool LC8, bool MODEL, bool STAGE>
__global__ void __launch_bounds__(CSIM_PREFIX_THREADS, 1) k_trials_prefix(PrefixArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int E = a.ro.E, C = E, R = a.b.R, G = a.ro.G;
    const int Ls = LC8 ? 3 : a.Ls, nch = a.nch, CP = nch << Ls, nblk = CP >> 3;
    const int Ei = (E + 7) & ~7;
    const int wpc = a.warps_per_cta;
    // ---- CTA-shared: staged tables of the current set, then roster tables --------------------------
    float* Ts = (float*)smem_raw;
    unsigned char* p = smem_raw + a.tab_smem;
    float* xs = (float*)p;                      p += sizeof(float) * CP;
    int32_t* reg = (int32_t*)p;                 p += sizeof(int32_t) * CP;
    uint32_t* regmask = (uint32_t*)p;           p += sizeof(uint32_t) * nblk * G;   // [blk][G]: bit i <-> candidate 8 blk + i is of region g
    uint32_t* papmask = (uint32_t*)p;           p += sizeof(uint32_t) * nblk;
    float* nbs = (float*)p;                                                         // [R] -beta_r log2 e of the current set
    for (int c = threadIdx.x; c < CP; c += blockDim.x) {
        xs[c] = c < C ? a.ro.x32[c] : 0.0f;
        reg[c] = c < C ? a.ro.region[c] : -1;
    }
    for (int i = threadIdx.x; i < nblk * G; i += blockDim.x) {
        const int blk = i / G, g = i - blk * G;
        uint32_t m = 0;
        for (int j = 0; j < 8; ++j) { const int c = 8 * blk + j; if (c < C && a.ro.region[c] == g) m |= 1u << j; }
        regmask[i] = m;
    }
    for (int blk = threadIdx.x; blk < nblk; blk += blockDim.x) {
        uint32_t m = 0;
        for (int j = 0; j < 8; ++j) { const int c = 8 * blk + j; if (c >= C) m |= 0x100u << j; else if (a.ro.pap[c]) m |= 1u << j; }
        papmask[blk] = m;
    }
    // ---- per-warp state ----------------------------------------------------------------------------
    unsigned char* base = smem_raw + a.cta_smem + (size_t)wib * a.warp_smem;
    float* bw = nullptr; float* bwp = nullptr;
    if (MODEL) {
        bw = (float*)base;                      base += sizeof(float) * CP;
        bwp = (float*)base;                     base += sizeof(float) * 32;
    }
    int32_t* tally = (int32_t*)base;            base += sizeof(int32_t) * Ei;
    int32_t* eff = (int32_t*)base;              base += sizeof(int32_t) * ((a.kmax + 3) & ~3);
    uint16_t* votes = nullptr;
    if (MODEL) { votes = (uint16_t*)base;       base += sizeof(uint16_t) * Ei; }
    uint8_t* mark = (uint8_t*)base;

    const int n_qit = a.e_quad_end >> 7;                     // quad iterations (128 electors each)
    const bool top2_ok = C >= 2 && C < 32768;
    const size_t round_pitch = (size_t)a.rows * a.stride;

    PrefixCtx kx;
    kx.xs = xs; kx.reg = reg; kx.regmask = regmask; kx.papmask = papmask; kx.bw = bw; kx.bwp = bwp;
    kx.C = C; kx.G = G; kx.Ls = Ls; kx.nch = nch; kx.nb = 0.f; kx.bonus = 0.f; kx.st1p = 1.f; kx.pwf = 1.f;
    kx.step0 = 1;
    while (kx.step0 * 2 <= nch) kx.step0 *= 2;

    // the electors of one round, each lane with its uniform: whole Philox quads per lane, then the tail one per lane
    auto for_each_elector = [&](int r, uint32_t s_lo, uint32_t s_hi, auto&& fn) {
        for (int qi = 0; qi < n_qit; ++qi) {
            const int q = lane + 32 * qi;
            if (4 * q < E) {
                uint32_t w[4];
                philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)q, a.b.seed_lo, a.b.seed_hi, w);
#pragma unroll 1
                for (int i = 0; i < 4; ++i) {
                    const int e = 4 * q + i;
                    const uint32_t wa = i == 0 ? w[0] : (i == 1 ? w[1] : (i == 2 ? w[2] : w[3]));
                    if (e < E) fn(e, u24_from_word(wa), (qi << 7) + (i << 5) + lane);
                }
            }
        }
        for (int e = a.e_quad_end + lane; e < E; e += 32) {
            uint32_t w[4];
            philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)(e >> 2), a.b.seed_lo, a.b.seed_hi, w);
            const int sel = e & 3;
            const uint32_t wa = sel == 0 ? w[0] : (sel == 1 ? w[1] : (sel == 2 ? w[2] : w[3]));
            fn(e, u24_from_word(wa), e);
        }
    };

    const uint64_t per_item = (uint64_t)wpc * a.tpi;
    const uint64_t items_per_set = (a.b.n_sims + per_item - 1) / per_item;
    const uint64_t n_items = items_per_set * a.b.n_sets;
    int cur_set = -1;
    float bw_strength = 0.f;
    int need = 0, k = 0, rr = 0;
    bool has_sticky = false, has_bw = false;
    unsigned long long acc_rounds = 0, acc_trials = 0;
    const float* Tset = a.T;

    for (uint64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int set = (int)(it / items_per_set);
        const uint64_t i0 = (it % items_per_set) * per_item;
        if (set != cur_set) {                                   // uniform over the CTA: every warp walks the same items
            if (cur_set >= 0 && lane == 0 && a.out.totals && acc_trials) {
                atomicAdd(&a.out.totals[2 * cur_set], acc_rounds); atomicAdd(&a.out.totals[2 * cur_set + 1], acc_trials);
            }
            acc_rounds = acc_trials = 0;
            cur_set = set;
            const DevSet* S = &a.b.sets[set];
            kx.bonus = S->f_bonus; kx.st1p = S->f_st1p; kx.pwf = S->f_pwf; bw_strength = S->f_bandwagon;
            need = S->need; k = S->k; rr = S->rr;
            has_sticky = MODEL && S->has_sticky; has_bw = MODEL && S->has_bw;
            Tset = a.T + (size_t)set * a.set_pitch;
            __syncthreads();                                    // readers of the previous set's tables are done
            if (STAGE) {
                // asynchronous 16-byte copies global -> shared (LDGSTS), all of a thread's chunks in flight at once
                const uint32_t n16 = (uint32_t)(a.tab_smem / 16);
                const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(Ts);
                const char* gbase = reinterpret_cast<const char*>(Tset);
                for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + i * 16), "l"(gbase + (size_t)i * 16) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            for (int r = threadIdx.x; r < R; r += blockDim.x) nbs[r] = a.nbeta[(size_t)set * R + r];
            __syncthreads();
        }
        for (int ti = 0; ti < a.tpi; ++ti) {
            const uint64_t i = i0 + (uint64_t)ti * wpc + wib;
            if (i >= a.b.n_sims) break;
            const uint64_t t = (uint64_t)set * a.b.n_sims + i;
            const uint64_t sim = a.b.sim_id_begin + i;
            const uint32_t s_lo = (uint32_t)sim, s_hi = (uint32_t)(sim >> 32);
            const bool fast2 = (k == 2) && top2_ok;              // two-candidate run-off: O(1) closing
            bool have_eff = false, have_prev = false;
            int winner = -1, rounds = 0, reason = CSIM_REASON_MAX_ROUNDS;
            int ef0 = 0, ef1 = 0, prev_mx = 0;
            if (R == 0) { for (int c = lane; c < C; c += 32) tally[c] = 0; }
            if (MODEL) {
                for (int c = lane; c < CP; c += 32) bw[c] = 0.0f;
                bwp[lane] = 0.0f;
            }
            __syncwarp();

            for (int r = 1; r <= R; ++r) {
                rounds = r;
                kx.nb = nbs[r - 1];
                const float* Tr = (STAGE ? Ts : Tset) + (size_t)(r - 1) * round_pitch;
                if (MODEL && has_bw && have_prev && prev_mx > 0) {               // bandwagon, model.py:326-354
                    const float inv = 1.0f / (float)prev_mx;
                    for (int c = lane; c < C; c += 32) bw[c] = (float)tally[c] * inv * bw_strength;   // tally = previous round's
                    __syncwarp();
                    float s = 0.f;                                               // bandwagon of chunk `lane`, then prefix over chunks
                    if (lane < nch) {
                        const float4* b4 = reinterpret_cast<const float4*>(bw + (lane << Ls));
                        for (int i = 0; i < (1 << (Ls - 2)); ++i) { const float4 v = b4[i]; s += (v.x + v.y) + (v.z + v.w); }
                    }
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const float o = __shfl_up_sync(CSIM_FULL, s, off);
                        if (lane >= off) s += o;
                    }
                    bwp[lane] = s;
                }
                __syncwarp();
                for (int c = lane; c < C; c += 32) tally[c] = 0;
                __syncwarp();
                const bool sticky_live = MODEL && have_prev && has_sticky;

                int cnt0 = 0;                                        // run-off (fast2): this lane's votes for ef0
                if (!have_eff) {
                    // =============== full-field round ==============================================
                    for_each_elector(r, s_lo, s_hi, [&](int e, float u, int slot) {
                        const int prev = sticky_live ? (int)votes[e] : -1;
                        const int vote = prefix_draw<LC8, MODEL>(kx, Tr + (size_t)slot * a.stride, e, u, prev);
                        atomicAdd(&tally[vote], 1);
                        if (MODEL) votes[e] = (uint16_t)vote;        // read back only by this lane: in place is safe
                    });
                } else {
                    // =============== run-off round: first k columns (simulate.py:144-154) ===========
                    for_each_elector(r, s_lo, s_hi, [&](int e, float u, int slot) {
                        const float xe = xs[e];
                        const int ge = reg[e];
                        const int prev = sticky_live ? (int)votes[e] : -1;
                        int vote;
                        if (fast2) {
                            const float v0 = prefix_full_score<MODEL>(kx, e, 0, xe, ge, prev);
                            const float v1 = prefix_full_score<MODEL>(kx, e, 1, xe, ge, prev);
                            const float sum = v0 + v1;
                            int j;
                            if (!(sum > 0.0f)) j = (u * 2.0f >= 1.0f) ? 1 : 0;       // simulate.py:151 uniform fallback
                            else j = (v0 > u * sum || !(v1 > 0.0f)) ? 0 : 1;
                            cnt0 += (j == 0);
                            vote = j == 0 ? ef0 : ef1;
                        } else {
                            float sum = 0.f;
                            for (int j = 0; j < k; ++j) sum += prefix_full_score<MODEL>(kx, e, j, xe, ge, prev);
                            int jsel;
                            if (!(sum > 0.0f)) {
                                jsel = (int)(u * (float)k);
                            } else {
                                const float tgt = u * sum;
                                float cum = 0.f; jsel = -1; int lastpos = 0;
                                for (int j = 0; j < k; ++j) {
                                    const float v = prefix_full_score<MODEL>(kx, e, j, xe, ge, prev);
                                    cum += v;
                                    if (v > 0.0f) lastpos = j;
                                    if (jsel < 0 && v > 0.0f && cum > tgt) jsel = j;
                                }
                                if (jsel < 0) jsel = lastpos;
                            }
                            if (jsel >= k) jsel = k - 1;
                            vote = eff[jsel];
                            atomicAdd(&tally[vote], 1);
                        }
                        if (MODEL) votes[e] = (uint16_t)vote;
                    });
                }
                int n0 = 0, n1 = 0;
                if (have_eff && fast2) {
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) cnt0 += __shfl_xor_sync(CSIM_FULL, cnt0, off);
                    n0 = cnt0; n1 = E - cnt0;
                    if (lane == 0) { tally[ef0] = n0; tally[ef1] = n1; }
                }
                __syncwarp();
                // ---------------- close the round -------------------------------------------------
                if (a.out.round_tallies)
                    for (int c = lane; c < C; c += 32)
                        a.out.round_tallies[(t * (uint64_t)R + (uint64_t)(r - 1)) * (uint64_t)C + c] = tally[c];
                have_prev = true;
                if (fast2) {
                    int mx, arg, second;
                    if (have_eff && n0 > 0 && n1 > 0) {
                        const bool first0 = n0 > n1 || (n0 == n1 && ef0 < ef1);
                        mx = first0 ? n0 : n1; arg = first0 ? ef0 : ef1; second = first0 ? ef1 : ef0;
                    } else {
                        uint32_t k1, k2;
                        warp_top2(tally, C, lane, k1, k2);
                        mx = (int)(k1 >> 16); arg = 0xFFFF - (int)(k1 & 0xFFFFu); second = 0xFFFF - (int)(k2 & 0xFFFFu);
                    }
                    prev_mx = mx;
                    if (mx >= need) { winner = arg; reason = CSIM_REASON_WINNER; break; }        // simulate.py:174-180
                    if (r >= rr) { ef0 = arg; ef1 = second; have_eff = true; }                   // simulate.py:183-186
                } else {
                    int mx, arg;
                    warp_argmax_first(tally, C, lane, mx, arg);
                    prev_mx = mx;
                    if (mx >= need) { winner = arg; reason = CSIM_REASON_WINNER; break; }
                    if (r >= rr && k > 0) { warp_top_k(tally, C, k, lane, eff, mark); have_eff = true; }
                }
                __syncwarp();
            }
            // ---------------- trial outputs -------------------------------------------------------
            if (a.out.final_votes) for (int c = lane; c < C; c += 32) a.out.final_votes[t * (uint64_t)C + c] = tally[c];
            if (a.out.round_tallies)
                for (size_t i2 = (size_t)rounds * C + lane; i2 < (size_t)R * C; i2 += 32) a.out.round_tallies[t * (uint64_t)R * C + i2] = -1;
            if (lane == 0) {
                a.out.winner[t] = winner; a.out.rounds[t] = rounds;
                if (a.out.reason) a.out.reason[t] = (uint8_t)reason;
                if (a.out.winner_hist) atomicAdd(&a.out.winner_hist[(size_t)set * (C + 1) + (winner >= 0 ? winner : C)], 1ULL);
                if (a.out.rounds_hist && winner >= 0) atomicAdd(&a.out.rounds_hist[(size_t)set * (R + 1) + rounds], 1ULL);
            }
            acc_rounds += (unsigned long long)rounds; acc_trials += 1;
            __syncwarp();
        }
    }
    if (cur_set >= 0 && lane == 0 && a.out.totals && acc_trials) {
        atomicAdd(&a.out.totals[2 * cur_set], acc_rounds); atomicAdd(&a.out.totals[2 * cur_set + 1], acc_trials);
    }
}
