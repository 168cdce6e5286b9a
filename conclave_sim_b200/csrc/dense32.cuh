// dense32.cuh — the fp32 dense engine: every trial re-evaluates all E x C scores each round.
//
// This is the kernel the roofline is quoted on (SURVEY §8d "canonical dense per-trial
// formulation").  One warp owns one trial and walks it from round 1 to its end, then takes the
// next trial (static stride over trials: results never depend on the schedule).  No block-level
// barrier exists after the roster tables are staged; only __syncwarp.
//
// Full-field round, for every elector e (candidates are cut into chunks of L, L = 8 up to E = 256):
//
//   phase 1   chunk_sum[j] = sum over c in chunk j of  exp(-beta_r |x_e - x_c|) * pw_c
//             evaluated for NE (= 2) electors of a lane at once, so every candidate operand fetched
//             from shared memory (a warp-uniform 128-bit load) feeds NE x 4 pair evaluations.  The terms that need no exponential are added per CHUNK:
//               self vote       = - s_ee  in e's chunk            (model.py:418-424)
//               stickiness      = + st * s_e,prev  in prev's chunk (model.py:357-367)
//             and the running prefix over chunks goes to a per-lane column of shared memory.
//             The terms that are sums of per-candidate constants are not accumulated at all: they are
//             read from prefix tables when a prefix is probed:
//               regional bonus  = bonus * #{c <= chunk j : region_c == region_e}   (roster table RBp)
//               bandwagon       = sum of bw_c over chunks <= j                     (per trial-round BWp)
//   phase 2   t = u * S; a 5-probe binary search over the prefix column finds the chunk that t falls
//             into; ONLY that chunk's L candidates are re-evaluated with every term inline, in
//             roster order, to find the first c with cum > t
//             (== searchsorted(cumsum(p)/sum, u, 'right'), simulate.py:153).
//   The E mod 128 electors that do not fill a quad per lane are split P lanes per elector (P = 4 at
//   E = 135), each lane summing every P-th chunk, so the leftover costs ~1/4 of a full pass.
//
// Factorised evaluation (template flag FACT, the default whenever beta_r >= 0 and
// beta_r * half-range(x) <= 40): with x' = x - x_mid,
//     exp(-beta |x_e - x_c|) = min( e^{-beta x'_e} e^{+beta x'_c} ,  e^{+beta x'_e} e^{-beta x'_c} )
//                            = A_e * min( BP_c , R_e * BQ_c ) / pw_c     with A = e^{-beta x'}, R = e^{+2 beta x'},
//                                                                        BP = e^{+beta x'} pw, BQ = e^{-beta x'} pw
// so a pair costs ONE multiply, a min and an add instead of FADD + FMUL + MUFU.EX2 + FFMA; the multiply and
// the add run as sm_100 packed-fp32 instructions (FMUL2 / FADD2, two pairs per issue slot): 2 issue slots
// per pair.  The factor A_e multiplies the chunk prefix once, when it is probed.  The per-round vectors
// come from k_round_tables32 (fp64 exp, rounded once).  FACT = false keeps the MUFU.EX2 form (any roster,
// any beta).
//
// Run-off rounds (simulate.py:144-154,183-186) only need the first k columns of the row
// (P[e, :k] renormalised), so they cost k pair evaluations per elector; with k = 2 the tally is a
// warp-reduced count and the round closes in O(1).
#pragma once
#include "common.cuh"

struct Dense32Args {
    RosterDev ro;
    BatchDev b;
    TrialOut out;
    const float* nbeta;     // [sets][R]  -beta_r * log2(e)
    const float* tab;       // [sets][R][4][nch*L]  A, A', BP, BP' (FACT only)
    int wiring;
    int L;                  // candidates per chunk (multiple of 4)
    int nch;                // number of chunks, nch * L >= C, nch <= 32
    int kmax;
    int warps_per_cta;
    size_t cta_smem;        // bytes of CTA-shared tables
    size_t warp_smem;       // bytes per warp
    // PROBE instantiation only (csim_engine_probs): one trial is started AT round probe_round from the given previous-round
    // state, and every draw of that round also dumps what it computes with: the chunk prefixes its search probes and the
    // re-scan scores of every chunk
    int probe_round;
    const int32_t* probe_prev_vote;     // [E] or null
    const int32_t* probe_prev_count;    // [E] or null
    float* probe_scores;                // [E][C]
    float* probe_prefix;                // [E][probe_cap]
    int probe_cap;
};

#ifndef CSIM_NE
#define CSIM_NE 2          // electors of a lane evaluated together in phase 1 (1, 2 or 4)
#endif
#ifndef CSIM_MINB
#define CSIM_MINB 3        // resident CTAs per SM the register allocation is bounded for
#endif
#define DENSE32_PREF_COLS (CSIM_NE > 1 ? CSIM_NE : 1)
__host__ __device__ inline size_t dense32_cta_smem(int nch, int L, int G) {
    const size_t CP = (size_t)nch * L;
    return ((CP * 4 /*xs*/ + CP * 4 /*reg*/ + (size_t)nch * G * 4 * 2 /*RBp, regmask*/) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t dense32_warp_smem(int E, int nch, int L, int kmax, bool model) {
    const size_t CP = (size_t)nch * L, Ei = ((size_t)E + 3) & ~(size_t)3, n4 = ((size_t)nch + 3) & ~(size_t)3;
    size_t n = CP * 4 * 2 /*pw | BP, BQ*/ + (size_t)DENSE32_PREF_COLS * nch * 32 * 4 /*prefix columns*/ + Ei * 4 /*tally*/ +
               (((size_t)kmax + 3) & ~(size_t)3) * 4 /*eff*/ + Ei /*mark*/;
    if (model) n += CP * 4 /*bw*/ + n4 * 4 /*bwp*/ + Ei * 4 * 2 /*prev_tally, votes*/;
    return (n + 15) & ~(size_t)15;
}

// sm_100 packed fp32 (two lanes of a 64-bit register pair per instruction)
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
    return ((unsigned long long)__float_as_uint(hi) << 32) | (unsigned long long)__float_as_uint(lo);
}
__device__ __forceinline__ float f2_lo(unsigned long long v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float f2_hi(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }

// K0 (FACT): per (set, round) vectors of the factorised exponential, fp64 exp rounded once to fp32.
// grid = (ceil(CP/128), sets*R), block = 128.
static __global__ void k_round_tables32(RosterDev ro, const DevSet* sets, const double* betas, int R, int CP, double xm, float* tab, int sr0) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int sr = sr0 + blockIdx.y;
    if (c >= CP) return;
    float A = 0.f, Ap = 0.f, BP = 0.f, BPp = 0.f;
    if (c < ro.E) {
        const double beta = betas[sr], xp = ro.x64[c] - xm;
        const double ea = exp(-beta * xp), eb = exp(beta * xp);
        const double pw = ro.pap[c] ? sets[sr / R].pwf : 1.0;
        A = (float)ea; Ap = (float)eb; BP = (float)(eb * pw); BPp = (float)(ea * pw);
    }
    float* t = tab + (size_t)sr * 4 * CP;
    t[c] = A; t[CP + c] = Ap; t[2 * CP + c] = BP; t[3 * CP + c] = BPp;
}

struct LaneCtx32 {              // per-warp / per-round constants of the draw
    const float* xs; const int32_t* reg; const float* rbp; const uint32_t* regmask;
    const float* pw; const float* bq; const float* bw; const float* bwp;   // FACT: pw = BP, bq = BQ;  bwp = prefix of bw over chunks
    float nb, bonus, st1p;
    int C, G, L, nch, step0;
    bool use_bw;
    float* dbg_scores; float* dbg_prefix; int dbg_cap;      // PROBE only
};

// One elector's scalars.  FACT: score(e,c) = A * min(BP_c, R * BQ_c);  else A = 1 and x is the ideology.
struct Elector32 { float x, A, Ainv, R; int g, prev; };

// the exponential part of one score in "m" units (score = A * m)
template <bool FACT>
__device__ __forceinline__ float pair_m32(const LaneCtx32& k, int c, const Elector32& el) {
    return FACT ? fminf(k.pw[c], el.R * k.bq[c]) : ex2_approx(fabsf(el.x - k.xs[c]) * k.nb) * k.pw[c];
}

// every term of one score, inline (run-off rounds and the generic re-scan)
template <bool MODEL, bool FACT>
__device__ __forceinline__ float full_score32(const LaneCtx32& k, int e, int c, const Elector32& el) {
    float v = el.A * pair_m32<FACT>(k, c, el);
    if (k.reg[c] == el.g) v += k.bonus;
    if (MODEL) {
        if (k.use_bw) v += k.bw[c];
        if (c == el.prev) v *= k.st1p;
    }
    return c == e ? 0.0f : v;
}

// Phase 1 of one chunk for NE electors of a lane: acc (packed pair of partial sums, m units) += chunk j.
// The candidate operands are loaded once (warp-uniform 128-bit loads) and reused by all NE electors.
template <int NE, int LC, bool FACT>
__device__ __forceinline__ void chunk_acc(const LaneCtx32& k, int j, const float (&xe)[NE], const unsigned long long (&r2)[NE],
                                          unsigned long long (&acc)[NE]) {
    const int L = LC > 0 ? LC : k.L;
    if (FACT) {
        const ulonglong2* b4 = reinterpret_cast<const ulonglong2*>(k.pw) + j * (L >> 2);
        const ulonglong2* c4 = reinterpret_cast<const ulonglong2*>(k.bq) + j * (L >> 2);
#pragma unroll 2
        for (int i = 0; i < (L >> 2); ++i) {
            const ulonglong2 b = b4[i], c = c4[i];
#pragma unroll
            for (int t = 0; t < NE; ++t) {
                unsigned long long t2 = f2_mul(r2[t], c.x);
                acc[t] = f2_add(acc[t], f2_pack(fminf(f2_lo(b.x), f2_lo(t2)), fminf(f2_hi(b.x), f2_hi(t2))));
                t2 = f2_mul(r2[t], c.y);
                acc[t] = f2_add(acc[t], f2_pack(fminf(f2_lo(b.y), f2_lo(t2)), fminf(f2_hi(b.y), f2_hi(t2))));
            }
        }
    } else {
        const float4* x4 = reinterpret_cast<const float4*>(k.xs + j * L);
        const float4* p4 = reinterpret_cast<const float4*>(k.pw + j * L);
        const float nb = k.nb;
#pragma unroll 2
        for (int i = 0; i < (L >> 2); ++i) {
            const float4 xv = x4[i], pv = p4[i];
#pragma unroll
            for (int t = 0; t < NE; ++t) {
                float a0 = f2_lo(acc[t]), a1 = f2_hi(acc[t]);
                a0 = fmaf(ex2_approx(fabsf(xe[t] - xv.x) * nb), pv.x, a0);
                a1 = fmaf(ex2_approx(fabsf(xe[t] - xv.y) * nb), pv.y, a1);
                a0 = fmaf(ex2_approx(fabsf(xe[t] - xv.z) * nb), pv.z, a0);
                a1 = fmaf(ex2_approx(fabsf(xe[t] - xv.w) * nb), pv.w, a1);
                acc[t] = f2_pack(a0, a1);
            }
        }
    }
}

// Per-elector corrections to the exponential-part prefix, in m units: the self score (removed at e's
// chunk, model.py:418-424) and stickiness (added at prev's chunk, model.py:357-367).
template <bool MODEL, bool FACT>
__device__ __forceinline__ void corrections32(const LaneCtx32& k, int e, const Elector32& el, float& see_m, float& sadd_m, int& jp) {
    float flat = k.bonus;                                  // the self pair is of e's own region
    if (MODEL && k.use_bw) flat += k.bw[e];
    see_m = pair_m32<FACT>(k, e, el) + flat * el.Ainv;     // same expression as phase 1, so it cancels
    jp = -1; sadd_m = 0.f;
    if (MODEL && el.prev >= 0 && el.prev != e) {
        float fl = (k.reg[el.prev] == el.g) ? k.bonus : 0.f;
        if (k.use_bw) fl += k.bw[el.prev];
        sadd_m = (pair_m32<FACT>(k, el.prev, el) + fl * el.Ainv) * (k.st1p - 1.0f);
        jp = el.prev / k.L;
    }
}

// prefix of the scores over chunks 0..j:  A * (exponential part) + regional bonus + bandwagon
template <bool MODEL>
__device__ __forceinline__ float probe32(const LaneCtx32& k, const float* pref, int pstride, const float* rbrow, float A, int j) {
    float v = fmaf(k.bonus, rbrow[j], A * pref[j * pstride]);
    if (MODEL && k.use_bw) v += k.bwp[j];
    return v;
}

// The 8 scores of chunk jstar of elector e's row with every term inline (LC == 8): what the re-scan of finish_draw sums.
template <bool MODEL, bool FACT>
__device__ __forceinline__ void rescan8_scores(const LaneCtx32& k, int jstar, int e, const Elector32& el, float (&v)[8]) {
    const int c0 = jstar * 8;
    // vectorised re-scan.  FACT: (xa, xb) hold BQ, (pa, pb) hold BP;  else x and pw
    const float* src_x = FACT ? k.bq : k.xs;
    const float4 xa = *reinterpret_cast<const float4*>(src_x + c0), xb = *reinterpret_cast<const float4*>(src_x + c0 + 4);
    const float4 pa = *reinterpret_cast<const float4*>(k.pw + c0), pb = *reinterpret_cast<const float4*>(k.pw + c0 + 4);
    const uint32_t rm = k.regmask[jstar * k.G + el.g];                   // bit i: candidate c0+i is of e's region
    const uint32_t dm = (jstar == (e >> 3)) ? (1u << (e & 7)) : 0u;      // bit i: candidate c0+i is e itself
    uint32_t pm = 0;
    float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
    if (MODEL) {
        if (el.prev >= 0 && (el.prev >> 3) == jstar) pm = 1u << (el.prev & 7);
        if (k.use_bw) { ba = *reinterpret_cast<const float4*>(k.bw + c0); bb = *reinterpret_cast<const float4*>(k.bw + c0 + 4); }
    }
    const float nb = k.nb;
#define CSIM_RESCAN(i, xv, pv, bv)                                              \
    {                                                                       \
        const float m = FACT ? fminf((pv), el.R * (xv)) : ex2_approx(fabsf(el.x - (xv)) * nb) * (pv); \
        float s = fmaf(el.A, m, ((rm >> (i)) & 1u) ? k.bonus : 0.0f);       \
        if (MODEL) { s += (bv); if ((pm >> (i)) & 1u) s *= k.st1p; }        \
        if ((dm >> (i)) & 1u) s = 0.0f;                                     \
        v[i] = s;                                                           \
    }
    CSIM_RESCAN(0, xa.x, pa.x, ba.x) CSIM_RESCAN(1, xa.y, pa.y, ba.y) CSIM_RESCAN(2, xa.z, pa.z, ba.z) CSIM_RESCAN(3, xa.w, pa.w, ba.w)
    CSIM_RESCAN(4, xb.x, pb.x, bb.x) CSIM_RESCAN(5, xb.y, pb.y, bb.y) CSIM_RESCAN(6, xb.z, pb.z, bb.z) CSIM_RESCAN(7, xb.w, pb.w, bb.w)
#undef CSIM_RESCAN
}

// Phase 2: given the exponential-part prefix over chunks (pref[j * pstride], m units), draw the vote.
template <int LC, bool MODEL, bool FACT, bool PROBE = false>
__device__ __forceinline__ int finish_draw(const LaneCtx32& k, const float* pref, int pstride, int e, float u, const Elector32& el) {
    const int C = k.C, L = LC > 0 ? LC : k.L, nch = k.nch;
    const float* rbrow = k.rbp + el.g * nch;
    if (PROBE) {            // csim_engine_probs: the numbers this draw is made from
        for (int j = 0; j < nch; ++j) k.dbg_prefix[(size_t)e * k.dbg_cap + j] = probe32<MODEL>(k, pref, pstride, rbrow, el.A, j);
        for (int js = 0; js < nch; ++js) {
            if (LC == 8) {
                float v[8];
                rescan8_scores<MODEL, FACT>(k, js, e, el, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) if (js * 8 + i < C) k.dbg_scores[(size_t)e * C + js * 8 + i] = v[i];
            } else {
                for (int i = 0; i < L; ++i) if (js * L + i < C) k.dbg_scores[(size_t)e * C + js * L + i] = full_score32<MODEL, FACT>(k, e, js * L + i, el);
            }
        }
    }
    const float S = probe32<MODEL>(k, pref, pstride, rbrow, el.A, nch - 1);
    if (!(S > 0.0f)) {            // all scores vanished: the reference makes the row uniform INCLUDING self (model.py:447-449)
        const int j = (int)(u * (float)C);
        return j < C ? j : C - 1;
    }
    const float tgt = u * S;
    // Bound-free search for the number of chunks whose prefix is <= tgt, in [0, nch - 1] (see prefix.cuh / table.cuh): with
    // P = step0 = the largest power of two <= nch, one first probe picks between the windows [0, P) and [nch - P, nch); the
    // halving steps stay inside the window, and the last probe taken IS the prefix below the chosen chunk.
    int pos = 0;
    float below = 0.0f;
    if (nch > k.step0) {
        const float f = probe32<MODEL>(k, pref, pstride, rbrow, el.A, nch - k.step0 - 1);
        if (f <= tgt) { pos = nch - k.step0; below = f; }
    }
    for (int step = k.step0 >> 1; step > 0; step >>= 1) {
        const float f = probe32<MODEL>(k, pref, pstride, rbrow, el.A, pos + step - 1);
        if (f <= tgt) { pos += step; below = f; }
    }
    const int jstar = pos;
    const int c0 = jstar * L;
    float cum = below;
    int cnt = 0;
    if (LC == 8) {
        float v[8];
        rescan8_scores<MODEL, FACT>(k, jstar, e, el, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) { cum += v[i]; cnt += (cum <= tgt) ? 1 : 0; }
    } else {
        for (int i = 0; i < L; ++i) {
            const int c = c0 + i;
            const float v = c < C ? full_score32<MODEL, FACT>(k, e, c, el) : 0.0f;
            cum += v;
            cnt += (cum <= tgt) ? 1 : 0;
        }
    }
    int vote = c0 + cnt;                                 // == first c with cum > tgt (searchsorted 'right', simulate.py:153)
    const int maxc = min(c0 + L - 1, C - 1);
    vote = vote > maxc ? maxc : vote;                    // rounding at the chunk's upper edge
    if (vote == e) vote = e > 0 ? e - 1 : (C > 1 ? 1 : 0);
    return vote;
}

template <int LC, bool MODEL, bool FACT, bool PROBE = false>
__global__ void __launch_bounds__(256, CSIM_MINB) k_trials_dense32(Dense32Args a) {
    constexpr int NE = CSIM_NE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int E = a.ro.E, C = E, R = a.b.R, G = a.ro.G;
    const int L = LC > 0 ? LC : a.L;
    const int nch = a.nch;
    const int CP = nch * L;
    const int Ei = (E + 3) & ~3;
    // ---- CTA-shared roster tables ------------------------------------------------------------
    float* xs = (float*)smem_raw;
    int32_t* reg = (int32_t*)(xs + CP);
    float* rbp = (float*)(reg + CP);                         // [G][nch]: #{c in chunks 0..j : region_c == g}
    uint32_t* regmask = (uint32_t*)(rbp + nch * G);          // [nch][G]: bit i <-> candidate j*L+i is of region g
    for (int c = threadIdx.x; c < CP; c += blockDim.x) {
        xs[c] = c < C ? a.ro.x32[c] : 0.0f;
        reg[c] = c < C ? a.ro.region[c] : -1;
    }
    for (int i = threadIdx.x; i < nch * G; i += blockDim.x) {
        const int j = i / G, g = i - j * G;
        int n = 0; uint32_t m = 0;
        for (int c = 0; c < (j + 1) * L && c < C; ++c)
            if (a.ro.region[c] == g) { ++n; if (c >= j * L && c - j * L < 32) m |= 1u << (c - j * L); }
        rbp[g * nch + j] = (float)n;
        regmask[i] = m;
    }
    // ---- per-warp state ----------------------------------------------------------------------
    unsigned char* base = smem_raw + a.cta_smem + (size_t)wib * a.warp_smem;
    // 16-byte aligned float arrays first (they are read with 128-bit loads), then the int arrays
    float* pw = (float*)base;                   base += sizeof(float) * CP;      // FACT: BP of the current round
    float* bq = (float*)base;                   base += sizeof(float) * CP;      // FACT: BQ
    float* pref = (float*)base;                 base += sizeof(float) * DENSE32_PREF_COLS * nch * 32;   // prefix columns [t][j][lane]
    float* bw = nullptr; float* bwp = nullptr;
    if (MODEL) {
        bw = (float*)base;                      base += sizeof(float) * CP;
        bwp = (float*)base;                     base += sizeof(float) * ((nch + 3) & ~3);
    }
    int32_t* tally = (int32_t*)base;            base += sizeof(int32_t) * Ei;
    int32_t* prev_tally = nullptr; int32_t* votes = nullptr;
    if (MODEL) {
        prev_tally = (int32_t*)base;            base += sizeof(int32_t) * Ei;
        votes = (int32_t*)base;                 base += sizeof(int32_t) * Ei;
    }
    int32_t* eff = (int32_t*)base;              base += sizeof(int32_t) * ((a.kmax + 3) & ~3);
    uint8_t* mark = (uint8_t*)base;
    __syncthreads();   // the only block barrier: roster tables are read-only from here on

    const uint64_t n_trials = (uint64_t)a.b.n_sets * a.b.n_sims;
    const uint64_t warp_global = (uint64_t)blockIdx.x * a.warps_per_cta + wib;
    const uint64_t n_warps = (uint64_t)gridDim.x * a.warps_per_cta;
    const int m_quads = E >> 7;                              // whole Philox quads per lane
    const int n_main = 4 * m_quads;
    const int e_left0 = 128 * m_quads;                       // first elector that does not fill a quad per lane
    const int n_left = E - e_left0;
    const int e_single0 = e_left0 + lane;
    const int n_single = e_single0 < E ? (E - e_single0 + 31) >> 5 : 0;
    const int n_slots = n_main + n_single;
    const bool top2_ok = C >= 2 && C < 32768;

    LaneCtx32 kx;
    kx.xs = xs; kx.reg = reg; kx.rbp = rbp; kx.regmask = regmask; kx.pw = pw; kx.bq = bq; kx.bw = bw; kx.bwp = bwp;
    kx.C = C; kx.G = G; kx.L = L; kx.nch = nch; kx.use_bw = false; kx.nb = 0.f; kx.bonus = 0.f; kx.st1p = 1.f;
    kx.step0 = 1;
    while (kx.step0 * 2 <= nch) kx.step0 *= 2;
    kx.dbg_scores = a.probe_scores; kx.dbg_prefix = a.probe_prefix; kx.dbg_cap = a.probe_cap;

    int cur_set = -1;
    float bw_strength = 0.f;
    int need = 0, k = 0, rr = 0;
    bool has_sticky = false, has_bw = false;
    unsigned long long acc_rounds = 0, acc_trials = 0;
    const float* tabA = nullptr;

    // one elector's scalars for this round
    auto load_elector = [&](int e, bool sticky_live) {
        Elector32 el;
        if (FACT) { el.A = __ldg(tabA + e); el.Ainv = __ldg(tabA + CP + e); el.R = el.Ainv * el.Ainv; el.x = 0.f; }
        else { el.A = 1.f; el.Ainv = 1.f; el.R = 1.f; el.x = xs[e]; }
        el.g = reg[e];
        el.prev = (MODEL && sticky_live) ? votes[e] : -1;
        return el;
    };

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
            kx.bonus = S->f_bonus; kx.st1p = S->f_st1p; bw_strength = S->f_bandwagon;
            need = S->need; k = S->k; rr = S->rr;
            has_sticky = MODEL && S->has_sticky; has_bw = MODEL && S->has_bw;
            const float pwf = S->f_pwf;
            __syncwarp();
            if (!FACT) for (int c = lane; c < CP; c += 32) pw[c] = c < C ? (a.ro.pap[c] ? pwf : 1.0f) : 0.0f;
            if (MODEL) for (int c = lane; c < CP; c += 32) bw[c] = 0.0f;
            __syncwarp();
        }
        const float* nbeta = a.nbeta + (size_t)set * R;
        const uint32_t s_lo = (uint32_t)sim, s_hi = (uint32_t)(sim >> 32);
        const bool fast2 = (k == 2) && top2_ok;                  // two-candidate run-off: O(1) closing
        bool have_eff = false, have_prev = false;
        int winner = -1, rounds = 0, reason = CSIM_REASON_MAX_ROUNDS;
        int ef0 = 0, ef1 = 0;                                    // run-off pair when fast2
        if (R == 0) { for (int c = lane; c < C; c += 32) tally[c] = 0; __syncwarp(); }
        if (PROBE && MODEL && (a.probe_prev_vote || a.probe_prev_count)) {      // start from the given previous-round state
            for (int c = lane; c < C; c += 32) {
                prev_tally[c] = a.probe_prev_count ? a.probe_prev_count[c] : 0;
                votes[c] = a.probe_prev_vote ? a.probe_prev_vote[c] : -1;
            }
            have_prev = true;
            __syncwarp();
        }

        for (int r = PROBE ? a.probe_round : 1; r <= R; ++r) {
            rounds = r;
            kx.nb = __ldg(&nbeta[r - 1]);
            kx.use_bw = false;
            if (FACT) {                                          // stage this round's BP / BQ (trial-invariant vectors)
                tabA = a.tab + ((size_t)set * R + (r - 1)) * 4 * CP;
                __syncwarp();
                for (int c = lane; c < CP; c += 32) { pw[c] = __ldg(tabA + 2 * CP + c); bq[c] = __ldg(tabA + 3 * CP + c); }
            }
            if (MODEL) {
                if (has_bw && have_prev) {                                   // model.py:326-354
                    int mx, arg;
                    warp_argmax_first(prev_tally, C, lane, mx, arg);
                    if (mx > 0) {
                        kx.use_bw = true;
                        const float inv = 1.0f / (float)mx;
                        for (int c = lane; c < C; c += 32) bw[c] = (float)prev_tally[c] * inv * bw_strength;
                        __syncwarp();
                        float s = 0.f;                                       // bandwagon of chunk `lane`, then prefix over chunks
                        if (lane < nch) for (int c = lane * L; c < (lane + 1) * L; ++c) s += bw[c];
#pragma unroll
                        for (int off = 1; off < 32; off <<= 1) {
                            const float o = __shfl_up_sync(CSIM_FULL, s, off);
                            if (lane >= off) s += o;
                        }
                        if (lane < nch) bwp[lane] = s;
                    }
                }
            }
            for (int c = lane; c < C; c += 32) tally[c] = 0;
            __syncwarp();
            const bool sticky_live = have_prev && has_sticky;

            int cnt0 = 0;                                        // run-off (fast2): this lane's votes for ef0
            if (!have_eff) {
                // =============== full-field round: NE electors of a lane at once =====================
                for (int qi = 0; qi < m_quads; ++qi) {
                    const int q = lane + 32 * qi;
                    uint32_t w[4];
                    philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)q, a.b.seed_lo, a.b.seed_hi, w);
#pragma unroll 1
                    for (int g = 0; g < 4 / NE; ++g) {
                        const int e0 = 4 * q + NE * g;
                        Elector32 el[NE];
                        float xe[NE], see[NE], sadd[NE];
                        int jp[NE];
                        unsigned long long r2[NE], acc[NE];
#pragma unroll
                        for (int i = 0; i < NE; ++i) {
                            el[i] = load_elector(e0 + i, sticky_live);
                            corrections32<MODEL, FACT>(kx, e0 + i, el[i], see[i], sadd[i], jp[i]);
                            xe[i] = el[i].x;
                            r2[i] = f2_pack(el[i].R, el[i].R);
                            acc[i] = 0ull;
                        }
                        const int je = e0 / L;                   // the electors of a group share their chunk (L % 4 == 0)
                        float* pc[NE];
                        float corr[NE];                           // running self-vote / stickiness correction of the prefix
#pragma unroll
                        for (int i = 0; i < NE; ++i) { pc[i] = pref + (size_t)i * nch * 32 + lane; corr[i] = 0.f; }
#pragma unroll 1
                        for (int j = 0; j < nch; ++j) {
                            chunk_acc<NE, LC, FACT>(kx, j, xe, r2, acc);
                            const bool at_self = j == je;
#pragma unroll
                            for (int i = 0; i < NE; ++i) {
                                if (at_self) corr[i] -= see[i];
                                if (MODEL && j == jp[i]) corr[i] += sadd[i];
                                *pc[i] = (f2_lo(acc[i]) + f2_hi(acc[i])) + corr[i];     // running prefix, m units
                                pc[i] += 32;
                            }
                        }
#pragma unroll 1
                        for (int i = 0; i < NE; ++i) {
                            const int wi = NE * g + i;
                            const uint32_t wa = wi == 0 ? w[0] : (wi == 1 ? w[1] : (wi == 2 ? w[2] : w[3]));
                            Elector32 ei = el[0];
#pragma unroll
                            for (int m = 1; m < NE; ++m) if (i == m) ei = el[m];
                            const int vote = finish_draw<LC, MODEL, FACT, PROBE>(kx, pref + (size_t)i * nch * 32 + lane, 32, e0 + i,
                                                                                 u24_from_word(wa), ei);
                            atomicAdd(&tally[vote], 1);
                            if (MODEL) votes[e0 + i] = vote;   // read only by this lane for its own electors: in-place is safe
                        }
                    }
                }
                // =============== leftover electors: P lanes per elector, every P-th chunk each ========
                for (int gb = 0; gb < n_left; gb += 32) {
                    const int n_it = min(32, n_left - gb);
                    int P = 1;
                    while (P * 2 * n_it <= 32) P *= 2;
                    const int eli = lane / P, sub = lane - eli * P;
                    const bool act = eli < n_it;
                    const int e = e_left0 + gb + (act ? eli : 0);
                    uint32_t w[4];
                    philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)(e >> 2), a.b.seed_lo, a.b.seed_hi, w);
                    const int sel = e & 3;
                    const uint32_t wa = sel == 0 ? w[0] : (sel == 1 ? w[1] : (sel == 2 ? w[2] : w[3]));
                    const Elector32 el = load_elector(e, sticky_live);
                    float see, sadd; int jp;
                    corrections32<MODEL, FACT>(kx, e, el, see, sadd, jp);
                    const float xe1[1] = {el.x};
                    const unsigned long long r2[1] = {f2_pack(el.R, el.R)};
                    const int je = e / L;
                    float* cl = pref + eli * nch;                // this elector's chunk sums, then prefixes (m units)
                    __syncwarp();
                    if (act) {
                        for (int j = sub; j < nch; j += P) {
                            unsigned long long acc[1] = {f2_pack((j == je) ? -see : 0.f, (MODEL && j == jp) ? sadd : 0.f)};
                            chunk_acc<1, LC, FACT>(kx, j, xe1, r2, acc);
                            cl[j] = f2_lo(acc[0]) + f2_hi(acc[0]);
                        }
                    }
                    __syncwarp();
                    if (act && sub == 0) {
                        float run = 0.f;
                        for (int j = 0; j < nch; ++j) { run += cl[j]; cl[j] = run; }
                        const int vote = finish_draw<LC, MODEL, FACT, PROBE>(kx, cl, 1, e, u24_from_word(wa), el);
                        atomicAdd(&tally[vote], 1);
                        if (MODEL) votes[e] = vote;
                    }
                    __syncwarp();
                }
            } else {
                // =============== run-off round: one elector per slot, first k columns ==================
                uint32_t w[4] = {0, 0, 0, 0};
                for (int n = 0; n < n_slots; ++n) {
                    int e, q; bool gen;
                    if (n < n_main) { q = lane + 32 * (n >> 2); e = 4 * q + (n & 3); gen = (n & 3) == 0; }
                    else { e = e_single0 + 32 * (n - n_main); q = e >> 2; gen = true; }
                    if (gen) philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)q, a.b.seed_lo, a.b.seed_hi, w);
                    const int sel = e & 3;
                    const uint32_t wa = sel == 0 ? w[0] : (sel == 1 ? w[1] : (sel == 2 ? w[2] : w[3]));
                    const float u = u24_from_word(wa);
                    const Elector32 el = load_elector(e, sticky_live);
                    int vote;
                    if (fast2) {
                        // two-candidate run-off: columns 0 and 1, renormalised (simulate.py:144-154)
                        const float v0 = full_score32<MODEL, FACT>(kx, e, 0, el);
                        const float v1 = full_score32<MODEL, FACT>(kx, e, 1, el);
                        const float sum = v0 + v1;
                        int j;
                        if (!(sum > 0.0f)) j = (u * 2.0f >= 1.0f) ? 1 : 0;       // simulate.py:151 uniform fallback
                        else j = (v0 > u * sum || !(v1 > 0.0f)) ? 0 : 1;
                        cnt0 += (j == 0);
                        vote = j == 0 ? ef0 : ef1;
                    } else {
                        float sum = 0.f;
                        for (int j = 0; j < k; ++j) sum += full_score32<MODEL, FACT>(kx, e, j, el);
                        int jsel;
                        if (!(sum > 0.0f)) {
                            jsel = (int)(u * (float)k);
                        } else {
                            const float tgt = u * sum;
                            float cum = 0.f; jsel = -1; int lastpos = 0;
                            for (int j = 0; j < k; ++j) {
                                const float v = full_score32<MODEL, FACT>(kx, e, j, el);
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
                    if (MODEL) votes[e] = vote;
                }
            }
            int n0 = 0, n1 = 0;
            if (have_eff && fast2) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) cnt0 += __shfl_xor_sync(CSIM_FULL, cnt0, off);
                n0 = cnt0; n1 = E - cnt0;
                if (lane == 0) { tally[ef0] = n0; tally[ef1] = n1; }
            }
            __syncwarp();
            if (PROBE) break;                                    // one round is all a probe plays
            // ---------------- close the round -----------------------------------------------------
            if (a.out.round_tallies)
                for (int c = lane; c < C; c += 32)
                    a.out.round_tallies[(t * (uint64_t)R + (uint64_t)(r - 1)) * (uint64_t)C + c] = tally[c];
            if (MODEL) for (int c = lane; c < C; c += 32) prev_tally[c] = tally[c];
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
                if (mx >= need) { winner = arg; reason = CSIM_REASON_WINNER; break; }        // simulate.py:174-180
                if (r >= rr) { ef0 = arg; ef1 = second; have_eff = true; }                   // simulate.py:183-186
            } else {
                int mx, arg;
                warp_argmax_first(tally, C, lane, mx, arg);
                if (mx >= need) { winner = arg; reason = CSIM_REASON_WINNER; break; }
                if (r >= rr && k > 0) { warp_top_k(tally, C, k, lane, eff, mark); have_eff = true; }
            }
            __syncwarp();
        }
        // ---------------- trial outputs -----------------------------------------------------------
        write_trial_outputs(a.out, t, set, C, R, lane, tally, winner, rounds, reason);
        acc_rounds += (unsigned long long)rounds; acc_trials += 1;
        __syncwarp();
    }
    if (cur_set >= 0 && lane == 0 && a.out.totals) {
        atomicAdd(&a.out.totals[2 * cur_set], acc_rounds); atomicAdd(&a.out.totals[2 * cur_set + 1], acc_trials);
    }
}
