// dense32.cuh — the fp32 dense engine: every trial re-evaluates all E x C scores each round.
//
// This is the kernel the roofline is quoted on (SURVEY §8d "canonical dense per-trial
// formulation").  One warp owns one trial and walks it from round 1 to its end, then takes the
// next trial of its CTA's work item (a CTA walks items of ONE parameter set at a time, so the
// per-set tables below are CTA-shared; results never depend on the schedule).  After the tables
// of a set are staged there is no block-level barrier; only __syncwarp.
//
// Full-field round, for every elector e (candidates are cut into chunks of L, L = 8 up to E = 256):
//
//   phase 1   chunk_sum[j] = sum over c in chunk j of  exp(-beta_r |x_e - x_c|) * pw_c
//             evaluated for NE (= 2) electors of a lane at once, so every candidate operand fetched
//             from shared memory (a warp-uniform 128-bit load) feeds NE x 4 pair evaluations; the running
//             prefix over chunks goes to a per-lane column of shared memory.  Nothing else happens in
//             this loop — 36 of its 44 instructions are the pair arithmetic and its operand loads.
//             The terms that need no exponential enter when a prefix is PROBED:
//               self vote       = - s_ee  from e's chunk on           (model.py:418-424)     one compare + add
//               stickiness      = + st * s_e,prev  from prev's chunk on (model.py:357-367)   one compare + add
//               regional bonus  = bonus * #{c <= chunk j : region_c == region_e}   (roster table RBp)
//               bandwagon       = sum of bw_c over chunks <= j                     (per trial-round BWp)
//   phase 2   t = u * S; a bound-free binary search over the prefix column (1 + log2 P probes) finds the
//             chunk that t falls into; ONLY that chunk's L candidates are re-evaluated with every term
//             inline, in roster order, to find the first c with cum > t
//             (== searchsorted(cumsum(p)/sum, u, 'right'), simulate.py:153).
//   The E mod 128 electors that do not fill a quad per lane are split P lanes per elector (P = 4 at
//   E = 135), each lane summing every P-th chunk, so the leftover costs ~1/4 of a full pass.
//
// Factorised evaluation (template flag FACT, the default whenever beta_r >= 0 and
// beta_r * half-range(x) <= 26): with x' = x - x_mid,
//     exp(-beta |x_e - x_c|) = min( e^{-beta x'_e} e^{+beta x'_c} ,  e^{+beta x'_e} e^{-beta x'_c} )
//                            = A_e * min( BP_c , R_e * BQ_c ) / pw_c     with A = e^{-beta x'}, R = e^{+2 beta x'},
//                                                                        BP = e^{+beta x'} pw, BQ = e^{-beta x'} pw
// so a pair costs ONE multiply, a min and an add instead of FADD + FMUL + MUFU.EX2 + FFMA; the multiply and
// the add run as sm_100 packed-fp32 instructions (FMUL2 / FADD2, two pairs per issue slot): 2 issue slots
// per pair.  The factor A_e multiplies the chunk prefix once, when it is probed.  The per-round vectors
// come from k_round_tables32 (fp64 exp, rounded once).  FACT = false keeps the MUFU.EX2 form (any roster,
// any beta).
//
// Candidate operands of one round ("round table" T, shared memory).  Two operands per candidate: P (BP, or the papabile
// weight in the MUFU form) and X (BQ, or the ideology).  L = 8: interleaved per chunk, [P0..3 | X0..3 | P4..7 | X4..7 | pad],
// so ONE pointer that advances 80 bytes per chunk serves the four 128-bit loads of a chunk at immediate offsets; generic
// L: two planes [CP] + [CP].  FACT: the tables of ALL R rounds of the current parameter set are staged once per CTA
// (R x 2 CP floats: 21.8 KB at E = 135, R = 20) when they fit, else each warp stages the round it plays.
//
// Run-off rounds (simulate.py:144-154,183-186) only need the first k columns of the row
// (P[e, :k] renormalised), so they cost k pair evaluations per elector; with k = 2 the tally is a
// warp-reduced count and the round closes in O(1).
//
// EXT (template flag; chunks of 8, model wiring, factorised form; launched by eng_full.cu as the FULL engine's fast path): candidate
// fatigue and stop-candidate in the same two-phase draw.
//   fatigue (model.py:372-379) scales whole columns by a per-trial factor phi_c: a round with fatigued candidates plays from the
//     warp's own copy of the round table with phi folded in, the bandwagon row is stored as bw * phi, and the regional prefix the
//     probes read is rebuilt per round from bit rows (popc of fatigue bits & region bits): no extra instruction per pair;
//   stop-candidate (model.py:384-415) scales, for a TRIGGERED elector, the whole score of the candidates acceptable to it: in a
//     round with threats the pair loop carries every term and a second, byte-masked accumulator (chunk_ext8); "triggered" =
//     (threat bits & ~acceptable bits) != 0; acceptable(e, c) is decided in fp64 once per parameter set (k_full_accmask bits,
//     k_ext_accbytes bytes);
//   the fatigue set is the reference's (simulate.py:84-118): tallies of the last F rounds in a ring of bytes, the immune top-n on
//     packed integer keys with an fp64 tie routine out of line.
// The quad-dealt and the leftover electors of a round go through ONE loop over "steps" so that the per-elector tail exists once
// (instruction footprint: DESIGN.md 5.4b), one elector per lane (CSIM_EXT_NE), 20 warps per SM at 96 registers.
#pragma once
#include "common.cuh"

struct Dense32Args {
    RosterDev ro;
    BatchDev b;
    TrialOut out;
    const float* nbeta;     // [sets][R]  -beta_r * log2(e)
    const float* tab;       // [sets][R][4][nch*L]  A, 1/A, BP, BQ (FACT only; k_round_tables32)
    int wiring;
    int L;                  // candidates per chunk (multiple of 4)
    int nch;                // number of chunks, nch * L >= C, nch <= 32
    int kmax;
    int warps_per_cta;
    WorkQueue wq;           // dynamic trial scheduling (common.cuh)
    int cta_tabs;           // 1: the round tables of the whole parameter set are CTA-shared (n_tab rounds); 0: per warp, per round
    int n_tab;              // rounds staged per set when cta_tabs (FACT: R; MUFU form: 1 — its operands do not depend on the round)
    int n_tab_full;         // of those, the first n_tab_full are whole tables (the rounds that can be full-field); the later ones are
                            // run-off rounds, which only ever read the first kmax candidates: nk_chunks chunks each (L = 8 only)
    int nk_chunks;
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
    // EXT instantiation only (candidate fatigue + stop-candidate live, launched by eng_full.cu as the FULL engine's fast path)
    const uint32_t* accmask;            // [sets][E][accw] bit c of row e: candidate c is ACCEPTABLE to elector e (k_full_accmask, fp64-exact)
    int accw;
    const uint2* accb;                  // [sets][nch][epitch]: the same bits of (elector slot, chunk) as 8 bytes, 0x40 = acceptable (k_ext_accbytes)
    int epitch;
    int Fmax;                           // longest fatigue window of the batch (rows of the per-warp tally history)
    size_t o_ext;                       // byte offset of the EXT part of the CTA-shared memory (accb | regb | regbits)
    const uint8_t* probe_fatigued;      // PROBE: the fatigue mask to apply (or null)
};

#ifndef CSIM_NE
#define CSIM_NE 2          // electors of a lane evaluated together in phase 1 (1, 2 or 4)
#endif
#ifndef CSIM_MINB
#define CSIM_MINB 3        // resident CTAs per SM the register allocation is bounded for
#endif
#ifndef CSIM_DENSE_PIPELINE
#define CSIM_DENSE_PIPELINE 1   // software-pipeline the operand loads of the pair loop (L = 8, factorised form)
#endif
#ifndef CSIM_EXT_THREADS
#define CSIM_EXT_THREADS 640    // EXT: threads per CTA the register allocation is bounded for (one CTA per SM): 20 warps at 96 registers.
#endif                          // Measured with one elector per lane (eng_full.cu: CSIM_EXT_NE = 1), ms per 250 k trials with fatigue + stop live:
                                // 512 / 576 / 608 / 640 / 768 threads = 14.04 / 14.40 / 13.76 / 12.94 / 14.33 (above 640 ptxas drops to 80 registers
                                // and spills 200 B; shared memory holds 20 warps); two electors per lane at 512 threads, 128 registers: 14.90
#ifndef CSIM_EXT_PLAY4
#define CSIM_EXT_PLAY4 0        // EXT, k = 2 run-off rounds: 1 = four unrolled copies of the per-elector body (as the plain kernel), 0 = one
#endif
#define DENSE32_PREF_COLS (CSIM_NE > 1 ? CSIM_NE : 1)
// Shared-memory layouts whose chunks are GATHERED per lane by the re-scan (each lane re-reads the chunk ITS target fell into, with
// 128-bit loads): with L = 8 a chunk is padded so that consecutive chunks start in different 16-byte bank groups — a round table
// chunk holds 16 operands + 4 floats of padding (stride 80 B: 5 j mod 8 walks all eight groups; at 64 B every chunk started in
// group 0 or 4 and the gathers serialised four-fold: 37 % of the kernel's shared wavefronts were conflicts), the per-candidate
// arrays (regional rows, bandwagon) 8 + 4 floats (3 j mod 8).  Generic L: plain arrays.
__host__ __device__ inline int dense32_tab_floats(int nch, int L) { return L == 8 ? nch * 20 : 2 * nch * L; }      // one round table
__host__ __device__ inline int dense32_cpp(int nch, int L) { return L == 8 ? nch * 12 : nch * L; }                  // padded per-candidate array
__host__ __device__ inline int dense32_rb_pitch(int nch, int L) { return dense32_cpp(nch, L) + 4; }
// byte offsets of the CTA-shared arrays: reg i32[CP] | RBp f32[G][nch] | RBf f32[G][pitch] | round tables f32[n_tab][2 CP]; all 16-byte aligned
__host__ __device__ inline size_t dense32_o_rbp(int nch, int L) { return (size_t)nch * L * 4; }
__host__ __device__ inline size_t dense32_o_rbf(int nch, int L, int G) { return dense32_o_rbp(nch, L) + (((size_t)nch * G * 4 + 15) & ~(size_t)15); }
__host__ __device__ inline size_t dense32_o_tabs(int nch, int L, int G) { return dense32_o_rbf(nch, L, G) + (size_t)G * dense32_rb_pitch(nch, L) * 4; }
__host__ __device__ inline size_t dense32_cta_smem(int nch, int L, int G, int n_tab_full, int n_tab_short = 0, int nk_chunks = 0) {
    return dense32_o_tabs(nch, L, G) + ((size_t)n_tab_full * dense32_tab_floats(nch, L) + (size_t)n_tab_short * dense32_tab_floats(nk_chunks, L)) * 4;
}
__host__ __device__ inline size_t dense32_warp_smem(int E, int nch, int L, int kmax, bool model, bool warp_tab) {
    const size_t CP = (size_t)nch * L, Ei = ((size_t)E + 3) & ~(size_t)3, n4 = ((size_t)nch + 3) & ~(size_t)3;
    size_t n = (warp_tab ? (size_t)dense32_tab_floats(nch, L) * 4 : 0) /*own round table*/ + (size_t)DENSE32_PREF_COLS * nch * 32 * 4 /*prefix columns*/ +
               Ei * 4 /*tally*/ + (((size_t)kmax + 3) & ~(size_t)3) * 4 /*eff*/ + Ei /*mark*/ + 16 + 2 * 32 * 16 /*wq, tailw: Philox words (16-byte aligned)*/;
    if (model) n += (size_t)dense32_cpp(nch, L) * 4 /*bw*/ + n4 * 4 /*bwp*/ + Ei * 4 * 2 /*prev_tally, votes*/;
    (void)CP;
    return (n + 15) & ~(size_t)15;
}

// EXT instantiation: what candidate fatigue + stop-candidate add to the shared-memory footprint (the warp part follows the arrays of
// dense32_warp_smem(..., model = true, warp_tab = true): a fatigued round plays from the warp's own, phi-folded copy of its round table)
__host__ __device__ inline size_t dense32_ext_warp_smem(int E, int nch, int G, int Fmax) {
    const size_t Ei = ((size_t)E + 3) & ~(size_t)3, n4 = ((size_t)nch + 3) & ~(size_t)3;
    const size_t n = (size_t)dense32_cpp(nch, 8) * 4 /*phi*/ + (size_t)G * n4 * 4 /*rbpw*/ + 2 * 8 * 4 /*fatbits, thrbits*/ + (size_t)(Fmax < 1 ? 1 : Fmax) * Ei /*hist*/;
    return (n + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t dense32_ext_cta_smem(int E, int nch, int G, int epitch) {
    const size_t n = (size_t)nch * epitch * 8 /*accb*/ + (size_t)G * nch * 8 /*regb*/ + (size_t)G * ((E + 31) / 32) * 4 /*regbits*/;
    return (n + 15) & ~(size_t)15;
}
// row of elector e in the acceptable-bytes table: the order in which the lanes of the trial kernel visit electors (quad-dealt
// electors 4 (lane + 32 qi) + i sit at 128 qi + 32 i + lane, so the 32 lanes of a pass read 32 consecutive 8-byte entries)
__host__ __device__ inline int dense32_slot(int e, int E) {
    return e < (E & ~127) ? ((e >> 7) << 7) + ((e & 3) << 5) + ((e >> 2) & 31) : e;
}

// sm_100 packed fp32 (two lanes of a 64-bit register pair per instruction)
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
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

// EXT: acceptable(e, c) = !(|x_e - x_c| > D) in fp64 exactly as model.py:390, for every set with stop-candidate enabled, as one byte per
// candidate (0x40 = acceptable: the top byte of the float 2.0) grouped by (chunk of 8 candidates, elector slot).
// grid = (ceil(E * nch / 128), sets), block = 128.
static __global__ void k_ext_accbytes(RosterDev ro, const DevSet* sets, int nch, int epitch, uint2* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, set = blockIdx.y;
    const int E = ro.E;
    if (i >= E * nch) return;
    const int e = i / nch, j = i - e * nch;
    const double xe = ro.x64[e], D = sets[set].stop_D;
    uint32_t lo = 0, hi = 0;
    for (int q = 0; q < 8; ++q) {
        const int c = 8 * j + q;
        if (c < E && !(fabs(xe - ro.x64[c]) > D)) {
            if (q < 4) lo |= 0x40u << (8 * q); else hi |= 0x40u << (8 * (q - 4));
        }
    }
    out[((size_t)set * nch + j) * epitch + dense32_slot(e, E)] = make_uint2(lo, hi);
}

// index of candidate c's P operand inside a round table; its X operand sits XOFF floats further
template <int LC> __device__ __forceinline__ int tab_off(int c) { return LC == 8 ? (c >> 3) * 20 + ((c & 4) << 1) + (c & 3) : c; }
// index of candidate c in the padded per-candidate arrays (regional rows, bandwagon)
template <int LC> __device__ __forceinline__ int pad_c(int c) { return LC == 8 ? c + ((c >> 3) << 2) : c; }

struct LaneCtx32 {              // per-warp / per-round constants of the draw
    const float* T;             // this round's table (see the header): P = BP | pw, X = BQ | x
    int xoff;                   // floats from a candidate's P operand to its X operand (4 when L = 8, else CP)
    const int32_t* reg; const float* rbp; const float* rbf; int rb_pitch;
    const float* bw; const float* bwp;   // bandwagon per candidate / its prefix over chunks (MODEL)
    float nb, bonus, st1p;
    int C, G, L, nch, step0;
    bool use_bw;
    float* dbg_scores; float* dbg_prefix; int dbg_cap;      // PROBE only
    // EXT only: phi = fatigue factor per candidate (padded like bw; 1 where not fatigued), bw then holds bandwagon * phi; rbp_cur = the
    // regional prefix the probes use (the roster's counts, or this round's phi-weighted ones); flat_in_m = this round's prefix columns
    // already hold every term (stop-candidate live: the masked pair loop), so a probe adds no table term
    const float* phi; const float* rbp_cur; float boost; bool flat_in_m;
};

// One elector's scalars.  FACT: score(e,c) = A * min(BP_c, R * BQ_c);  else A = 1 and x is the ideology.
struct Elector32 { float x, A, Ainv, R; int g, prev; bool trig; const uint32_t* arow; };     // trig / arow: EXT only (stop-candidate, model.py:384-415)

// The two point rules of a row as corrections of its chunk prefixes, in m units (score = A * m): the self score leaves from
// chunk je on (model.py:418-424), stickiness adds from chunk jp on (model.py:357-367).  Applied when a prefix is probed.
struct Corr32 { float see, sadd; int je, jp; };

// the exponential part of one score in "m" units (score = A * m)
template <int LC, bool FACT>
__device__ __forceinline__ float pair_m32(const LaneCtx32& k, int c, const Elector32& el) {
    const int o = tab_off<LC>(c);
    return FACT ? fminf(k.T[o], el.R * k.T[o + k.xoff]) : ex2_approx(fabsf(el.x - k.T[o + k.xoff]) * k.nb) * k.T[o];
}

// every term of one score, inline (run-off rounds and the generic re-scan)
// bit c of an elector's "acceptable" row (EXT)
__device__ __forceinline__ bool acc_bit(const uint32_t* arow, int c) { return (__ldg(arow + (c >> 5)) >> (c & 31)) & 1u; }

// EXT (run-off rounds: k.T is the round's plain table there): fatigue scales the score, bandwagon included (k.bw holds bandwagon * phi),
// and a triggered elector boosts the candidates acceptable to it (model.py:372-415)
template <int LC, bool MODEL, bool FACT, bool EXT = false>
__device__ __forceinline__ float full_score32(const LaneCtx32& k, int e, int c, const Elector32& el) {
    float v = el.A * pair_m32<LC, FACT>(k, c, el);
    if (k.reg[c] == el.g) v += k.bonus;
    if (EXT) v *= k.phi[pad_c<LC>(c)];
    if (MODEL) {
        if (k.use_bw) v += k.bw[pad_c<LC>(c)];
        if (c == el.prev) v *= k.st1p;
    }
    if (EXT) { if (el.trig && acc_bit(el.arow, c)) v *= k.boost; }
    return c == e ? 0.0f : v;
}

// Phase 1 of one chunk for NE electors of a lane: acc (packed pair of partial sums, m units) += chunk j.
// The candidate operands are loaded once (warp-uniform 128-bit loads) and reused by all NE electors.
// LC == 8: tc points at the chunk's 16 interleaved floats; generic L: the chunk's P plane, X plane xoff floats further.
template <int NE, int LC, bool FACT>
__device__ __forceinline__ void chunk_acc(const LaneCtx32& k, const float* tc, const float (&xe)[NE], const float (&Re)[NE],
                                          unsigned long long (&acc)[NE]) {
    if (LC == 8) {
        const ulonglong2* t4 = reinterpret_cast<const ulonglong2*>(tc);
        if (FACT) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const ulonglong2 b = t4[2 * h], c = t4[2 * h + 1];            // BP, BQ of four candidates
#pragma unroll
                for (int t = 0; t < NE; ++t) {
                    const unsigned long long r2 = f2_pack(Re[t], Re[t]);
                    unsigned long long t2 = f2_mul(r2, c.x);
                    acc[t] = f2_add(acc[t], f2_pack(fminf(f2_lo(b.x), f2_lo(t2)), fminf(f2_hi(b.x), f2_hi(t2))));
                    t2 = f2_mul(r2, c.y);
                    acc[t] = f2_add(acc[t], f2_pack(fminf(f2_lo(b.y), f2_lo(t2)), fminf(f2_hi(b.y), f2_hi(t2))));
                }
            }
        } else {
            const float4* f4 = reinterpret_cast<const float4*>(tc);
            const float nb = k.nb;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 pv = f4[2 * h], xv = f4[2 * h + 1];
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
    } else {
        const int L = k.L;
        const float4* p4 = reinterpret_cast<const float4*>(tc);
        const float4* x4 = reinterpret_cast<const float4*>(tc + k.xoff);
        const float nb = k.nb;
#pragma unroll 2
        for (int i = 0; i < (L >> 2); ++i) {
            const float4 pv = p4[i], xv = x4[i];
#pragma unroll
            for (int t = 0; t < NE; ++t) {
                float a0 = f2_lo(acc[t]), a1 = f2_hi(acc[t]);
                if (FACT) {
                    a0 += fminf(pv.x, Re[t] * xv.x); a1 += fminf(pv.y, Re[t] * xv.y);
                    a0 += fminf(pv.z, Re[t] * xv.z); a1 += fminf(pv.w, Re[t] * xv.w);
                } else {
                    a0 = fmaf(ex2_approx(fabsf(xe[t] - xv.x) * nb), pv.x, a0);
                    a1 = fmaf(ex2_approx(fabsf(xe[t] - xv.y) * nb), pv.y, a1);
                    a0 = fmaf(ex2_approx(fabsf(xe[t] - xv.z) * nb), pv.z, a0);
                    a1 = fmaf(ex2_approx(fabsf(xe[t] - xv.w) * nb), pv.w, a1);
                }
                acc[t] = f2_pack(a0, a1);
            }
        }
    }
}

// The factorised pair arithmetic of one chunk of 8 candidates on operands already in registers (b = BP, c = BQ of candidates
// 0-3 and 4-7): the software-pipelined form of the loop loads the NEXT chunk's operands while this one is evaluated.
template <int NE>
__device__ __forceinline__ void chunk_math8(const ulonglong2& b0, const ulonglong2& c0, const ulonglong2& b1, const ulonglong2& c1,
                                            const float (&Re)[NE], unsigned long long (&acc)[NE]) {
#pragma unroll
    for (int t = 0; t < NE; ++t) {
        const unsigned long long r2 = f2_pack(Re[t], Re[t]);
        unsigned long long t2 = f2_mul(r2, c0.x);
        acc[t] = f2_add(acc[t], f2_pack(fminf(f2_lo(b0.x), f2_lo(t2)), fminf(f2_hi(b0.x), f2_hi(t2))));
        t2 = f2_mul(r2, c0.y);
        acc[t] = f2_add(acc[t], f2_pack(fminf(f2_lo(b0.y), f2_lo(t2)), fminf(f2_hi(b0.y), f2_hi(t2))));
        t2 = f2_mul(r2, c1.x);
        acc[t] = f2_add(acc[t], f2_pack(fminf(f2_lo(b1.x), f2_lo(t2)), fminf(f2_hi(b1.x), f2_hi(t2))));
        t2 = f2_mul(r2, c1.y);
        acc[t] = f2_add(acc[t], f2_pack(fminf(f2_lo(b1.y), f2_lo(t2)), fminf(f2_hi(b1.y), f2_hi(t2))));
    }
}

// EXT, a round with stop-candidate threats: the pair loop carries EVERY term of the score, because a triggered elector multiplies
// the whole score of the candidates acceptable to it (model.py:400-415) and "acceptable" is a per-(elector, candidate) fact.
// In m units (score / A_e), for NE electors of a lane and one chunk of 8 candidates:
//     mm  = min(BP'_c, R_e BQ'_c) + Ainv_e * ( [region_c == region_e] * bonus * phi_c + bandwagon_c * phi_c )     (BP', BQ': phi folded)
//     acc += mm ;  accA += [acceptable(e, c)] * 2 mm        (the caller adds (boost - 1) / 2 * accA for a triggered elector)
// The two 0/1 facts come as bytes (0x40 = yes) that ONE byte permute turns into the floats 2.0 / 0.0 (the byte becomes the float's
// top byte), so a pair costs 1 FMNMX + 2 PRMT on the ALU pipe and 2.5 packed instructions on the FMA pipe; the candidate operands
// (table, phi, bandwagon) are warp-uniform 128-bit loads shared by the NE electors.
__device__ __forceinline__ float byte_w2(uint32_t w, int i) { return __uint_as_float(__byte_perm(w, 0u, 0x0444u | ((uint32_t)i << 12))); }

template <int NE>
__device__ __forceinline__ void chunk_ext8(const float* tc, const float* ph, const float* bwc, float hbonus, const uint2 (&ab)[NE], const uint2 (&rb)[NE],
                                           const float (&Re)[NE], const float (&Ai)[NE], unsigned long long (&acc)[NE], unsigned long long (&accA)[NE]) {
    const ulonglong2* t4 = reinterpret_cast<const ulonglong2*>(tc);
    const ulonglong2* p4 = reinterpret_cast<const ulonglong2*>(ph);
    const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(bwc);
    const unsigned long long hb2 = f2_pack(hbonus, hbonus);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const ulonglong2 b = t4[2 * h], c = t4[2 * h + 1];                    // BP', BQ' of four candidates
        const ulonglong2 f = p4[h], w = w4[h];                                // phi, bandwagon * phi
        const unsigned long long rbx = f2_mul(hb2, f.x), rby = f2_mul(hb2, f.y);   // bonus / 2 * phi
#pragma unroll
        for (int t = 0; t < NE; ++t) {
            const unsigned long long r2 = f2_pack(Re[t], Re[t]), a2 = f2_pack(Ai[t], Ai[t]);
            const uint32_t rw = h ? rb[t].y : rb[t].x, aw = h ? ab[t].y : ab[t].x;
            {
                const unsigned long long t2 = f2_mul(r2, c.x);
                const unsigned long long m2 = f2_pack(fminf(f2_lo(b.x), f2_lo(t2)), fminf(f2_hi(b.x), f2_hi(t2)));
                const unsigned long long fl = f2_fma(f2_pack(byte_w2(rw, 0), byte_w2(rw, 1)), rbx, w.x);
                const unsigned long long mm = f2_fma(a2, fl, m2);
                acc[t] = f2_add(acc[t], mm);
                accA[t] = f2_fma(f2_pack(byte_w2(aw, 0), byte_w2(aw, 1)), mm, accA[t]);
            }
            {
                const unsigned long long t2 = f2_mul(r2, c.y);
                const unsigned long long m2 = f2_pack(fminf(f2_lo(b.y), f2_lo(t2)), fminf(f2_hi(b.y), f2_hi(t2)));
                const unsigned long long fl = f2_fma(f2_pack(byte_w2(rw, 2), byte_w2(rw, 3)), rby, w.y);
                const unsigned long long mm = f2_fma(a2, fl, m2);
                acc[t] = f2_add(acc[t], mm);
                accA[t] = f2_fma(f2_pack(byte_w2(aw, 2), byte_w2(aw, 3)), mm, accA[t]);
            }
        }
    }
}

// EXT: is some threat of this round unacceptable to the elector (model.py:384-400)?  thr = the round's threat bits, arow = the elector's
// "acceptable" bits, W words each.
__device__ __forceinline__ bool ext_triggered(const uint32_t* thr, const uint32_t* arow, int W) {
    uint32_t any = 0;
#pragma unroll 1
    for (int w = 0; w < W; ++w) any |= thr[w] & ~__ldg(arow + w);
    return any != 0;
}

// Per-elector corrections to the exponential-part prefix, in m units: the self score (removed from e's
// chunk on, model.py:418-424) and stickiness (added from prev's chunk on, model.py:357-367).
template <int LC, bool MODEL, bool FACT>
__device__ __forceinline__ Corr32 corrections32(const LaneCtx32& k, int e, const Elector32& el) {
    Corr32 cr;
    float flat = k.bonus;                                  // the self pair is of e's own region
    if (MODEL && k.use_bw) flat += k.bw[pad_c<LC>(e)];
    cr.see = pair_m32<LC, FACT>(k, e, el) + flat * el.Ainv;   // same expression as phase 1, so it cancels
    cr.je = e / k.L;
    cr.jp = 0x7fffffff; cr.sadd = 0.f;
    if (MODEL && el.prev >= 0 && el.prev != e) {
        float fl = (k.reg[el.prev] == el.g) ? k.bonus : 0.f;
        if (k.use_bw) fl += k.bw[pad_c<LC>(el.prev)];
        cr.sadd = (pair_m32<LC, FACT>(k, el.prev, el) + fl * el.Ainv) * (k.st1p - 1.0f);
        cr.jp = el.prev / k.L;
    }
    return cr;
}

// EXT form of the same corrections (full-field rounds: k.T is the phi-folded table when a candidate is fatigued, k.bw = bandwagon * phi):
// the whole self score and the whole stickiness increment, in m units, boost included.
template <int LC, bool FACT>
__device__ __forceinline__ Corr32 corrections32_ext(const LaneCtx32& k, int e, const Elector32& el) {
    Corr32 cr;
    float fe = k.bonus * k.phi[pad_c<LC>(e)];              // the self pair is of e's own region
    if (k.use_bw) fe += k.bw[pad_c<LC>(e)];
    cr.see = fmaf(fe, el.Ainv, pair_m32<LC, FACT>(k, e, el));
    if (el.trig && acc_bit(el.arow, e)) cr.see *= k.boost;
    cr.je = e / k.L;
    cr.jp = 0x7fffffff; cr.sadd = 0.f;
    if (el.prev >= 0 && el.prev != e) {
        float fl = (k.reg[el.prev] == el.g) ? k.bonus * k.phi[pad_c<LC>(el.prev)] : 0.f;
        if (k.use_bw) fl += k.bw[pad_c<LC>(el.prev)];
        cr.sadd = fmaf(fl, el.Ainv, pair_m32<LC, FACT>(k, el.prev, el)) * (k.st1p - 1.0f);
        if (el.trig && acc_bit(el.arow, el.prev)) cr.sadd *= k.boost;
        cr.jp = el.prev / k.L;
    }
    return cr;
}

// prefix of the scores over chunks 0..j:  A * (exponential part + point rules) + regional bonus + bandwagon
template <bool MODEL, bool EXT = false>
__device__ __forceinline__ float probe32(const LaneCtx32& k, const float* pref, int pstride, const float* rbrow, float A, const Corr32& cr, int j) {
    float m = pref[j * pstride];
    m = j >= cr.je ? m - cr.see : m;
    if (MODEL) m = j >= cr.jp ? m + cr.sadd : m;
    if (EXT) { if (k.flat_in_m) return A * m; }             // warp-uniform
    float v = fmaf(k.bonus, rbrow[j], A * m);
    if (MODEL && k.use_bw) v += k.bwp[j];
    return v;
}

// The 8 scores of chunk jstar of elector e's row with every term inline (LC == 8): what the re-scan of finish_draw sums.
template <bool MODEL, bool FACT, bool EXT = false>
__device__ __forceinline__ void rescan8_scores(const LaneCtx32& k, int jstar, int e, const Elector32& el, float (&v)[8]) {
    const int c0 = jstar * 8;
    const float4* t4 = reinterpret_cast<const float4*>(k.T + 20 * jstar);
    const float4 pa = t4[0], xa = t4[1], pb = t4[2], xb = t4[3];           // P = BP | pw, X = BQ | x
    const float* rbrow = k.rbf + el.g * k.rb_pitch + 12 * jstar;           // bonus where candidate c0+i is of e's region, else 0 (padded chunks)
    const float4 ra = *reinterpret_cast<const float4*>(rbrow), rb = *reinterpret_cast<const float4*>(rbrow + 4);
    float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
    const int is = e - c0;                                                 // position of e itself / of prev inside this chunk (else outside 0..7)
    int ip = -1;
    if (MODEL) {
        ip = el.prev - c0;
        if (k.use_bw) { ba = *reinterpret_cast<const float4*>(k.bw + 12 * jstar); bb = *reinterpret_cast<const float4*>(k.bw + 12 * jstar + 4); }
    }
    const float nb = k.nb;
    // EXT: the table already carries phi (fatigue) on the exponential part and k.bw on the bandwagon; the regional bonus takes it here;
    // a triggered elector boosts the candidates of this chunk that are acceptable to it
    float4 fa = make_float4(1.f, 1.f, 1.f, 1.f), fb = fa;
    uint32_t abits = 0;
    if (EXT) {
        fa = *reinterpret_cast<const float4*>(k.phi + 12 * jstar); fb = *reinterpret_cast<const float4*>(k.phi + 12 * jstar + 4);
        if (el.trig) abits = (__ldg(el.arow + (c0 >> 5)) >> (c0 & 31)) & 0xFFu;
    }
#define CSIM_RESCAN(i, xv, pv, rv, bv, fv)                                      \
    {                                                                       \
        const float m = FACT ? fminf((pv), el.R * (xv)) : ex2_approx(fabsf(el.x - (xv)) * nb) * (pv); \
        float s = fmaf(el.A, m, EXT ? (rv) * (fv) : (rv));                  \
        if (MODEL) { s += (bv); if (ip == (i)) s *= k.st1p; }               \
        if (EXT) { if ((abits >> (i)) & 1u) s *= k.boost; }                 \
        if (is == (i)) s = 0.0f;                                            \
        v[i] = s;                                                           \
    }
    CSIM_RESCAN(0, xa.x, pa.x, ra.x, ba.x, fa.x) CSIM_RESCAN(1, xa.y, pa.y, ra.y, ba.y, fa.y) CSIM_RESCAN(2, xa.z, pa.z, ra.z, ba.z, fa.z) CSIM_RESCAN(3, xa.w, pa.w, ra.w, ba.w, fa.w)
    CSIM_RESCAN(4, xb.x, pb.x, rb.x, bb.x, fb.x) CSIM_RESCAN(5, xb.y, pb.y, rb.y, bb.y, fb.y) CSIM_RESCAN(6, xb.z, pb.z, rb.z, bb.z, fb.z) CSIM_RESCAN(7, xb.w, pb.w, rb.w, bb.w, fb.w)
#undef CSIM_RESCAN
}

// The draw of a row whose scores all vanished (model.py:447-449: the row becomes uniform INCLUDING self) — out of line, it is
// (almost) never taken and its instructions should not sit in every draw.
static __device__ __noinline__ int uniform_row_vote(float u, int C) {
    const int j = (int)(u * (float)C);
    return j < C ? j : C - 1;
}

// Phase 2: given the exponential-part prefix over chunks (pref[j * pstride], m units), draw the vote.
template <int LC, bool MODEL, bool FACT, bool PROBE = false, bool EXT = false>
__device__ __forceinline__ int finish_draw(const LaneCtx32& k, const float* pref, int pstride, int e, float u, const Elector32& el, const Corr32& cr) {
    const int C = k.C, L = LC > 0 ? LC : k.L, nch = k.nch;
    const float* rbrow = (EXT ? k.rbp_cur : k.rbp) + el.g * nch;
    if (PROBE) {            // csim_engine_probs: the numbers this draw is made from
        for (int j = 0; j < nch; ++j) k.dbg_prefix[(size_t)e * k.dbg_cap + j] = probe32<MODEL, EXT>(k, pref, pstride, rbrow, el.A, cr, j);
        for (int js = 0; js < nch; ++js) {
            if (LC == 8) {
                float v[8];
                rescan8_scores<MODEL, FACT, EXT>(k, js, e, el, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) if (js * 8 + i < C) k.dbg_scores[(size_t)e * C + js * 8 + i] = v[i];
            } else {
                for (int i = 0; i < L; ++i) if (js * L + i < C) k.dbg_scores[(size_t)e * C + js * L + i] = full_score32<LC, MODEL, FACT>(k, e, js * L + i, el);
            }
        }
    }
    const float S = probe32<MODEL, EXT>(k, pref, pstride, rbrow, el.A, cr, nch - 1);
    if (__builtin_expect(!(S > 0.0f), 0))          // all scores vanished: the reference makes the row uniform INCLUDING self (model.py:447-449)
        return uniform_row_vote(u, C);
    const float tgt = u * S;
    // Bound-free search for the number of chunks whose prefix is <= tgt, in [0, nch - 1] (see prefix.cuh / table.cuh): with
    // P = step0 = the largest power of two <= nch, one first probe picks between the windows [0, P) and [nch - P, nch); the
    // halving steps stay inside the window, and the last probe taken IS the prefix below the chosen chunk.  nch <= 32, so the
    // halving steps are 16, 8, 4, 2, 1 at most: unrolled with compile-time offsets, the ones >= step0 skipped (uniform branch).
    int pos = 0;
    float below = 0.0f;
    if (nch > k.step0) {
        const float f = probe32<MODEL, EXT>(k, pref, pstride, rbrow, el.A, cr, nch - k.step0 - 1);
        if (f <= tgt) { pos = nch - k.step0; below = f; }
    }
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        if (step < k.step0) {
            const float f = probe32<MODEL, EXT>(k, pref, pstride, rbrow, el.A, cr, pos + step - 1);
            if (f <= tgt) { pos += step; below = f; }
        }
    }
    const int jstar = pos;
    const int c0 = jstar * L;
    float cum = below;
    int cnt = 0;
    if (LC == 8) {
        float v[8];
        rescan8_scores<MODEL, FACT, EXT>(k, jstar, e, el, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) { cum += v[i]; cnt += (cum <= tgt) ? 1 : 0; }
    } else {
        for (int i = 0; i < L; ++i) {
            const int c = c0 + i;
            const float v = c < C ? full_score32<LC, MODEL, FACT>(k, e, c, el) : 0.0f;
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

// EXT, candidate fatigue: among the unmarked candidates whose vote sum over the window is S, the one with the largest MEAN SHARE as the
// reference computes it in fp64 (simulate.py:88-95: each round's votes / E, summed in round order, / rounds), ties -> highest index.
// Out of line: only reached when two candidates tie on the integer sum.
static __device__ __noinline__ int ext_top_mean_exact(const uint8_t* hist, const uint8_t* mark, const int* isum, int S, int n_hist, int nh, int Fmax,
                                                      int Ei, int C, int E, int lane) {
    double bv = -1.0; int ba = -1;
    for (int c = lane; c < C; c += 32) {
        if (mark[c] || isum[c] != S) continue;
        double acc = __ddiv_rn((double)hist[(size_t)((n_hist - nh) % Fmax) * Ei + c], (double)E);
        for (int q = 1; q < nh; ++q) acc = __dadd_rn(acc, __ddiv_rn((double)hist[(size_t)((n_hist - nh + q) % Fmax) * Ei + c], (double)E));
        const double av = __ddiv_rn(acc, (double)nh);
        if (ba < 0 || av >= bv) { bv = av; ba = c; }
    }
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(CSIM_FULL, bv, off);
        const int oa = __shfl_xor_sync(CSIM_FULL, ba, off);
        if (oa >= 0 && (ba < 0 || ov > bv || (ov == bv && oa > ba))) { bv = ov; ba = oa; }
    }
    return ba;
}

// One round's table from its source into shared memory, by `nthr` threads of which this is number `tid`:
// FACT: src = the round's [4][CP] block of Dense32Args::tab (planes 2, 3 = BP, BQ);  MUFU form: P = papabile weight, X = ideology.
template <int LC, bool FACT>
__device__ __forceinline__ void stage_round_table(float* dst, const Dense32Args& a, const float* src, float pwf, int CP, int C, int tid, int nthr,
                                                  int n_cand = 0x7fffffff) {
    const int xoff = LC == 8 ? 4 : CP;
    const int n = n_cand < CP ? n_cand : CP;                 // run-off rounds: only the first chunks are ever read
    for (int c = tid; c < n; c += nthr) {
        float P, X;
        if (FACT) { P = __ldg(src + 2 * CP + c); X = __ldg(src + 3 * CP + c); }
        else { P = c < C ? (a.ro.pap[c] ? pwf : 1.0f) : 0.0f; X = c < C ? a.ro.x32[c] : 0.0f; }
        const int o = tab_off<LC>(c);
        dst[o] = P; dst[o + xoff] = X;
    }
}

template <int LC, bool MODEL, bool FACT, bool PROBE = false, bool EXT = false>
__global__ void __launch_bounds__(EXT ? CSIM_EXT_THREADS : 256, EXT ? 1 : CSIM_MINB) k_trials_dense32(Dense32Args a) {
    static_assert(!EXT || (LC == 8 && MODEL && FACT), "the EXT instantiation exists for chunks of 8, the model wiring and the factorised exponential");
    constexpr int NE = CSIM_NE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int E = a.ro.E, C = E, R = a.b.R, G = a.ro.G;
    const int L = LC > 0 ? LC : a.L;
    const int nch = a.nch;
    const int CP = nch * L;
    const int Ei = (E + 3) & ~3;
    const int rbpitch = dense32_rb_pitch(nch, L);
    const int TF = dense32_tab_floats(nch, L), CPP = dense32_cpp(nch, L);
    const int TK = dense32_tab_floats(a.nk_chunks, L);       // a run-off round's short table (cta_tabs)
    // ---- CTA-shared roster tables ------------------------------------------------------------
    int32_t* reg = (int32_t*)smem_raw;
    float* rbp = (float*)(smem_raw + dense32_o_rbp(nch, L));           // [G][nch]: #{c in chunks 0..j : region_c == g}
    float* rbf = (float*)(smem_raw + dense32_o_rbf(nch, L, G));        // [G][rbpitch]: the current set's bonus where region_c == g, else 0
    float* tabs = (float*)(smem_raw + dense32_o_tabs(nch, L, G));      // [n_tab][2 CP] round tables of the current set (cta_tabs)
    // EXT: acceptable-bytes of the current set [nch][epitch], region-match bytes [G][nch] and region bit rows [G][Wb] of the roster
    uint2* accb_s = nullptr; uint2* regb = nullptr; uint32_t* regbits = nullptr;
    const int Wb = (C + 31) >> 5;
    if (EXT) {
        accb_s = (uint2*)(smem_raw + a.o_ext); regb = accb_s + (size_t)nch * a.epitch; regbits = (uint32_t*)(regb + G * nch);
        for (int i = threadIdx.x; i < G * nch; i += blockDim.x) {
            const int g = i / nch, j = i - g * nch;
            uint32_t lo = 0, hi = 0;
            for (int q = 0; q < 8; ++q) {
                const int c = 8 * j + q;
                if (c < C && a.ro.region[c] == g) { if (q < 4) lo |= 0x40u << (8 * q); else hi |= 0x40u << (8 * (q - 4)); }
            }
            regb[i] = make_uint2(lo, hi);
        }
        for (int i = threadIdx.x; i < G * Wb; i += blockDim.x) {
            const int g = i / Wb, w = i - g * Wb;
            uint32_t m = 0;
            for (int b = 0; b < 32; ++b) { const int c = 32 * w + b; if (c < C && a.ro.region[c] == g) m |= 1u << b; }
            regbits[i] = m;
        }
    }
    for (int c = threadIdx.x; c < CP; c += blockDim.x) reg[c] = c < C ? a.ro.region[c] : -1;
    for (int i = threadIdx.x; i < nch * G; i += blockDim.x) {
        const int j = i / G, g = i - j * G;
        int n = 0;
        for (int c = 0; c < (j + 1) * L && c < C; ++c) n += a.ro.region[c] == g;
        rbp[g * nch + j] = (float)n;
    }
    // ---- per-warp state ----------------------------------------------------------------------
    unsigned char* base = smem_raw + a.cta_smem + (size_t)wib * a.warp_smem;
    // 16-byte aligned float arrays first (they are read with 128-bit loads), then the int arrays
    float* wtab = (float*)base;                 if (!a.cta_tabs || EXT) base += sizeof(float) * TF;       // own round table (fallback; EXT: the phi-folded copy)
    float* pref = (float*)base;                 base += sizeof(float) * DENSE32_PREF_COLS * nch * 32;   // prefix columns [j][t][lane]
    float* bw = nullptr; float* bwp = nullptr;
    if (MODEL) {
        bw = (float*)base;                      base += sizeof(float) * CPP;
        bwp = (float*)base;                     base += sizeof(float) * ((nch + 3) & ~3);
    }
    int32_t* tally = (int32_t*)base;            base += sizeof(int32_t) * Ei;
    int32_t* prev_tally = nullptr; int32_t* votes = nullptr;
    if (MODEL) {
        prev_tally = (int32_t*)base;            base += sizeof(int32_t) * Ei;
        votes = (int32_t*)base;                 base += sizeof(int32_t) * Ei;
    }
    int32_t* eff = (int32_t*)base;              base += sizeof(int32_t) * ((a.kmax + 3) & ~3);
    uint8_t* mark = (uint8_t*)base;               base += (Ei + 15) & ~15;
    uint32_t* wq = (uint32_t*)base;               // [32][4]: the quad of Philox words of each lane's current four electors (16-byte slots)
    uint32_t* tailw = wq + 32 * 4;                // [32][4]: the quads of the electors beyond the last whole 128, for several rounds ahead
    // EXT per-warp state: fatigue factor per candidate (padded like bw), phi-weighted regional prefix, fatigue / threat bit rows, and
    // the tallies of the last Fmax rounds (bytes: E <= 255)
    float* phi = nullptr; float* rbpw = nullptr; uint32_t* fatbits = nullptr; uint32_t* thrbits = nullptr; uint8_t* hist = nullptr;
    if (EXT) {
        unsigned char* xb = (unsigned char*)(tailw + 32 * 4);
        phi = (float*)xb;          xb += sizeof(float) * CPP;
        rbpw = (float*)xb;         xb += sizeof(float) * G * ((nch + 3) & ~3);
        fatbits = (uint32_t*)xb;   xb += 32;
        thrbits = (uint32_t*)xb;   xb += 32;
        hist = xb;
    }

    const int m_quads = E >> 7;                              // whole Philox quads per lane
    const int n_main = 4 * m_quads;
    const int e_left0 = 128 * m_quads;                       // first elector that does not fill a quad per lane
    const int n_left = E - e_left0;
    const int e_single0 = e_left0 + lane;
    const int n_single = e_single0 < E ? (E - e_single0 + 31) >> 5 : 0;
    const int n_slots = n_main + n_single;
    const bool top2_ok = C >= 2 && C < 32768;

    LaneCtx32 kx;
    kx.T = tabs; kx.xoff = LC == 8 ? 4 : CP;
    kx.reg = reg; kx.rbp = rbp; kx.rbf = rbf; kx.rb_pitch = rbpitch; kx.bw = bw; kx.bwp = bwp;
    kx.C = C; kx.G = G; kx.L = L; kx.nch = nch; kx.use_bw = false; kx.nb = 0.f; kx.bonus = 0.f; kx.st1p = 1.f;
    kx.step0 = 1;
    while (kx.step0 * 2 <= nch) kx.step0 *= 2;
    kx.dbg_scores = a.probe_scores; kx.dbg_prefix = a.probe_prefix; kx.dbg_cap = a.probe_cap;
    kx.phi = phi; kx.rbp_cur = rbp; kx.boost = 1.f; kx.flat_in_m = false;
    // EXT: the current set's fatigue / stop-candidate constants, and this round's number of threats (0: no elector can be triggered)
    bool en_fat = false, en_stop = false, neg_T = false;
    int Fw = 0, fat_top_n = 0, fat_cnt = 0, stop_cnt = 0, n_thr = 0;
    float pen = 1.f, bm1h = 0.f;

    int cur_set = -1;
    float bw_strength = 0.f, pwf = 1.f;
    int need = 0, k = 0, rr = 0;
    bool has_sticky = false, has_bw = false;
    unsigned long long acc_rounds = 0, acc_trials = 0;
    uint32_t tabA = 0;                        // element offset of this round's A | 1/A | BP | BQ block inside a.tab (32-bit: the launcher checks the size)

    // one elector's scalars for this round
    auto load_elector = [&](int e, bool sticky_live, int trig_known = -1) {
        Elector32 el;
        if (FACT) { el.A = __ldg(a.tab + (tabA + (uint32_t)e)); el.Ainv = __ldg(a.tab + (tabA + (uint32_t)(CP + e))); el.R = el.Ainv * el.Ainv; el.x = 0.f; }
        else { el.A = 1.f; el.Ainv = 1.f; el.R = 1.f; el.x = __ldg(a.ro.x32 + e); }
        el.g = reg[e];
        el.prev = (MODEL && sticky_live) ? votes[e] : -1;
        el.trig = false; el.arow = nullptr;
        if (EXT) {
            if (n_thr > 0) {                                 // stop-candidate: triggered when some threat is unacceptable to e (model.py:384-400)
                el.arow = a.accmask + ((size_t)cur_set * E + e) * a.accw;
                el.trig = trig_known >= 0 ? trig_known != 0 : ext_triggered(thrbits, el.arow, Wb);
            }
        }
        return el;
    };

    // The electors beyond the last whole 128 (n_left of them, nlq quads of Philox words per round): ONE Philox call per lane stashes
    // their words for rpf = 32 / nlq rounds ahead (the counter is (sim, round, quad), so any lane can compute any round's quad).
    const int nlq = (n_left + 3) >> 2;
    const int rpf = nlq > 0 ? 32 / nlq : 0;          // rounds per fill (n_left <= 127 -> nlq <= 32 -> rpf >= 1)
    const int q_left0 = e_left0 >> 2;

    __shared__ int s_skip[2];
    for (int si = 0; si < a.b.n_sets; ++si) {
        const int set = cta_set_of_step(si, a.b.n_sets);         // uniform over the CTA
        if (threadIdx.x == 0) s_skip[si & 1] = set_exhausted(a.wq, set, a.b.n_sims) ? 1 : 0;
        __syncthreads();                                        // the answer is published; readers of the previous set's tables are done
        if (s_skip[si & 1]) continue;
        {
            if (cur_set >= 0 && lane == 0 && a.out.totals && acc_trials) {
                atomicAdd(&a.out.totals[2 * cur_set], acc_rounds); atomicAdd(&a.out.totals[2 * cur_set + 1], acc_trials);
            }
            acc_rounds = acc_trials = 0;
            cur_set = set;
            const DevSet* S = &a.b.sets[set];
            kx.bonus = S->f_bonus; kx.st1p = S->f_st1p; bw_strength = S->f_bandwagon; pwf = S->f_pwf;
            need = S->need; k = S->k; rr = S->rr;
            has_sticky = MODEL && S->has_sticky; has_bw = MODEL && S->has_bw;
            for (int i = threadIdx.x; i < G * CP; i += blockDim.x) {
                const int g = i / CP, c = i - g * CP;
                rbf[g * rbpitch + pad_c<LC>(c)] = (c < C && a.ro.region[c] == g) ? kx.bonus : 0.0f;
            }
            if (a.cta_tabs)
                for (int rt = 0; rt < a.n_tab; ++rt) {
                    const bool whole = rt < a.n_tab_full;
                    float* dst = whole ? tabs + (size_t)rt * TF : tabs + (size_t)a.n_tab_full * TF + (size_t)(rt - a.n_tab_full) * TK;
                    stage_round_table<LC, FACT>(dst, a, FACT ? a.tab + ((size_t)set * R + rt) * 4 * CP : nullptr, pwf, CP, C, threadIdx.x, blockDim.x,
                                                whole ? CP : a.nk_chunks * L);
                }
            if (MODEL) for (int c = lane; c < CPP; c += 32) bw[c] = 0.0f;
            if (EXT) {
                en_fat = S->en_fat != 0; en_stop = S->en_stop != 0; neg_T = 0.0 > S->stop_T;
                Fw = S->fat_rounds; fat_top_n = S->fat_top_n; fat_cnt = S->fat_cnt; stop_cnt = S->stop_cnt;
                pen = S->f_fat_pen; kx.boost = S->f_stop_boost; bm1h = 0.5f * (S->f_stop_boost - 1.0f);
                if (en_stop) {
                    const uint2* src = a.accb + (size_t)set * nch * a.epitch;
                    for (int i = threadIdx.x; i < nch * a.epitch; i += blockDim.x) accb_s[i] = __ldg(src + i);
                }
            }
            __syncthreads();
        }
        const float* nbeta = a.nbeta + (size_t)set * R;
        for (uint64_t i0 = warp_claim_trials(a.wq, set, lane); i0 < a.b.n_sims; i0 = warp_claim_trials(a.wq, set, lane))
        for (int ti = 0; ti < a.wq.chunk; ++ti) {
            const uint64_t i = i0 + (uint64_t)ti;
            if (i >= a.b.n_sims) break;
            const uint64_t t = (uint64_t)set * a.b.n_sims + i;
            const uint64_t sim = a.b.sim_id_begin + i;
            const uint32_t s_lo = (uint32_t)sim, s_hi = (uint32_t)(sim >> 32);
            const bool fast2 = (k == 2) && top2_ok;                  // two-candidate run-off: O(1) closing
            bool have_eff = false, have_prev = false;
            int winner = -1, rounds = 0, reason = CSIM_REASON_MAX_ROUNDS;
            int ef0 = 0, ef1 = 0;                                    // run-off pair when fast2
            int tail_r0 = -(1 << 30);                                // first round of the current tail stash (none yet)
            if (R == 0) { for (int c = lane; c < C; c += 32) tally[c] = 0; __syncwarp(); }
            int n_hist = 0;                                          // EXT: rounds in the tally history
            int prev_mx = 0;                                         // EXT: the largest tally of the previous round (the closing computes it)
            int hrow = 0;                                            // EXT: row of the history ring the next tally goes to (= n_hist % Fmax)
            if (EXT) {
#pragma unroll 1
                for (int c = lane; c < CPP; c += 32) { phi[c] = 1.0f; bw[c] = 0.0f; }
                n_thr = 0;
            }
            if (PROBE && MODEL && (a.probe_prev_vote || a.probe_prev_count)) {      // start from the given previous-round state
                for (int c = lane; c < C; c += 32) {
                    prev_tally[c] = a.probe_prev_count ? a.probe_prev_count[c] : 0;
                    votes[c] = a.probe_prev_vote ? a.probe_prev_vote[c] : -1;
                    if (EXT) hist[c] = (uint8_t)prev_tally[c];
                }
                have_prev = true; n_hist = 1; hrow = a.Fmax > 1 ? 1 : 0;
                __syncwarp();
                if (EXT) { int arg0; warp_argmax_first(prev_tally, C, lane, prev_mx, arg0); }
            }

            for (int r = PROBE ? a.probe_round : 1; r <= R; ++r) {
                rounds = r;
                kx.nb = __ldg(&nbeta[r - 1]);
                kx.use_bw = false;
                if (FACT) tabA = (uint32_t)(set * R + (r - 1)) * (uint32_t)(4 * CP);
                if (a.cta_tabs) {
                    const int rt = FACT ? r - 1 : 0;
                    kx.T = rt < a.n_tab_full ? tabs + (size_t)rt * TF : tabs + (size_t)a.n_tab_full * TF + (size_t)(rt - a.n_tab_full) * TK;
                } else {                                             // stage this round's table (trial-invariant vectors) for this warp
                    __syncwarp();
                    stage_round_table<LC, FACT>(wtab, a, FACT ? a.tab + tabA : nullptr, pwf, CP, C, lane, 32);
                    kx.T = wtab;
                }
                bool use_fat = false;
                if (EXT) {
                    n_thr = 0;
                    __syncwarp();
                    // ---- candidate fatigue: the set, simulate.py:84-118, decided in fp64 exactly as exact64.cuh decides it; the
                    //      factor per candidate, model.py:372-379
                    if (en_fat) {
                        bool window = false;
                        if (r > Fw) {
#pragma unroll 1
                            for (int c = lane; c < C; c += 32) mark[c] = 0;
                            __syncwarp();
                            const int nh = n_hist < Fw ? n_hist : Fw;
                            if (fat_top_n > 0 && n_hist > 0 && nh > 0) {
                                // top-n by mean share over the window, ties -> highest index.  The mean is the reference's fp64 expression
                                // (sum of votes / E, / rounds); it is ordered like the integer vote sum except between candidates whose
                                // sums are EQUAL (different compositions may round differently), so the ranking runs on packed integer
                                // keys (sum << 8 | index, E <= 255) and only a tie at the top goes through the fp64 routine.
                                int* isum = (int*)pref;                          // the prefix columns are not live between rounds
#pragma unroll 1
                                for (int c = lane; c < C; c += 32) isum[c] = 0;
#pragma unroll 1
                                for (int q = 0; q < nh; ++q) {
                                    int row = hrow - nh + q;                     // the ring row of the q-th round of the window
                                    row = row < 0 ? row + a.Fmax : row;
                                    const uint8_t* hr = hist + row * Ei;
#pragma unroll 1
                                    for (int c = lane; c < C; c += 32) isum[c] += (int)hr[c];
                                }
                                __syncwarp();
                                const int n_imm = fat_top_n < C ? fat_top_n : C;
#pragma unroll 1
                                for (int j = 0; j < n_imm; ++j) {
                                    int best = -1;
#pragma unroll 1
                                    for (int c = lane; c < C; c += 32) if (!mark[c]) best = max(best, (isum[c] << 8) | c);
#pragma unroll
                                    for (int off = 16; off > 0; off >>= 1) best = max(best, __shfl_xor_sync(CSIM_FULL, best, off));
                                    int pick = best & 255;
                                    if (nh > 1) {
                                        bool tie = false;
#pragma unroll 1
                                        for (int c = lane; c < C; c += 32) tie = tie || (!mark[c] && c != pick && isum[c] == (best >> 8));
                                        if (__any_sync(CSIM_FULL, tie)) pick = ext_top_mean_exact(hist, mark, isum, best >> 8, n_hist, nh, a.Fmax, Ei, C, E, lane);
                                    }
                                    if (lane == 0) mark[pick] = 1;
                                    __syncwarp();
                                }
                            }
                            window = n_hist >= Fw;
                        }
                        uint32_t anyf = 0;
#pragma unroll 1
                        for (int c0 = 0; c0 < 32 * Wb; c0 += 32) {
                            const int c = c0 + lane;
                            bool f = false;
                            if (window && c < C && !mark[c]) {
                                f = true;
#pragma unroll 1
                                for (int q = 0; q < Fw && f; ++q) {                // share < threshold, as an integer test (DevSet::fat_cnt)
                                    int row = hrow - Fw + q;
                                    row = row < 0 ? row + a.Fmax : row;
                                    f = (int)hist[row * Ei + c] < fat_cnt;
                                }
                            }
                            if (PROBE && a.probe_fatigued) f = c < C && a.probe_fatigued[c] != 0;
                            const uint32_t m = __ballot_sync(CSIM_FULL, f);
                            anyf |= m;
                            if (lane == 0) fatbits[c0 >> 5] = m;
                            if (c < CP) phi[pad_c<LC>(c)] = f ? pen : 1.0f;
                        }
                        use_fat = anyf != 0;
                    }
                    // ---- stop-candidate threats, model.py:384-400: candidates whose last share exceeds the threshold.  Round 1 has no
                    //      previous shares: the reference tests 0.0 > threshold against every candidate (simulate.py:73, model.py:397),
                    //      so with a NEGATIVE threat share every candidate is a threat already then
                    if (en_stop && (have_prev || neg_T)) {
                        const bool all_thr = !have_prev;
#pragma unroll 1
                        for (int c0 = 0; c0 < 32 * Wb; c0 += 32) {
                            const int c = c0 + lane;
                            const bool t = c < C && (all_thr || prev_tally[c] >= stop_cnt);   // share > threshold (DevSet::stop_cnt)
                            const uint32_t m = __ballot_sync(CSIM_FULL, t);
                            n_thr += __popc(m);
                            if (lane == 0) thrbits[c0 >> 5] = m;
                        }
                    }
                    __syncwarp();
                    kx.flat_in_m = n_thr > 0;
                    kx.rbp_cur = rbp;
                    if (!have_eff && use_fat) {                  // a full-field round with fatigued candidates plays from the phi-folded table
                        const float* Ts = kx.T;
#pragma unroll 1
                        for (int c = lane; c < CP; c += 32) {
                            const int o = tab_off<LC>(c);
                            const float f = phi[pad_c<LC>(c)];
                            wtab[o] = Ts[o] * f; wtab[o + 4] = Ts[o + 4] * f;
                        }
                        kx.T = wtab;
                        if (n_thr == 0) {                        // regional prefix with phi: (same-region candidates) - (1 - pen) (fatigued ones)
#pragma unroll 1
                            for (int g = lane; g < G; g += 32) {
                                int nf = 0;
#pragma unroll 1
                                for (int j = 0; j < nch; ++j) {
                                    const int sh = 8 * (j & 3);
                                    nf += __popc((fatbits[j >> 2] >> sh) & (regbits[g * Wb + (j >> 2)] >> sh) & 0xFFu);
                                    rbpw[g * nch + j] = (rbp[g * nch + j] - (float)nf) + pen * (float)nf;
                                }
                            }
                            kx.rbp_cur = rbpw;
                        }
                    }
                }
                if (MODEL) {
                    if (has_bw && have_prev) {                                   // model.py:326-354
                        int mx, arg;
                        if constexpr (EXT) { mx = prev_mx; arg = 0; }           // the previous round's closing already found it
                        else warp_argmax_first(prev_tally, C, lane, mx, arg);
                        (void)arg;
                        if (mx > 0) {
                            kx.use_bw = true;
                            const float inv = 1.0f / (float)mx;
                            // EXT: a run-off round only ever reads the first k candidates' terms (no full-field round follows it)
                            const int nbw = (EXT && have_eff) ? min(C, max(k, 2)) : C;
                            for (int c = lane; c < nbw; c += 32) {
                                float v = (float)prev_tally[c] * inv * bw_strength;
                                if (EXT) v *= phi[pad_c<LC>(c)];                  // fatigue scales the whole score (model.py:379)
                                bw[pad_c<LC>(c)] = v;
                            }
                            __syncwarp();
                            // (EXT: the prefix over chunks is probed only in a full-field round without stop-candidate threats)
                            if (!EXT || (!have_eff && n_thr == 0)) {
                                float s = 0.f;                                   // bandwagon of chunk `lane`, then prefix over chunks
                                if (lane < nch) for (int c = lane * L; c < (lane + 1) * L; ++c) s += bw[pad_c<LC>(c)];
#pragma unroll
                                for (int off = 1; off < 32; off <<= 1) {
                                    const float o = __shfl_up_sync(CSIM_FULL, s, off);
                                    if (lane >= off) s += o;
                                }
                                if (lane < nch) bwp[lane] = s;
                            }
                        }
                    }
                }
                for (int c4 = lane; c4 < (Ei >> 2); c4 += 32) reinterpret_cast<int4*>(tally)[c4] = make_int4(0, 0, 0, 0);
                if (nlq > 0 && (unsigned)(r - tail_r0) >= (unsigned)rpf) {         // (re)fill the tail stash: rounds r .. r + rpf - 1
                    tail_r0 = r;
                    const int rd = lane / nlq, tq = lane - rd * nlq;
                    if (rd < rpf) {
                        uint32_t w[4];
                        philox4x32_10(s_lo, s_hi, (uint32_t)(r + rd), (uint32_t)(q_left0 + tq), a.b.seed_lo, a.b.seed_hi, w);
                        *reinterpret_cast<uint4*>(tailw + 4 * lane) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                __syncwarp();
                const uint32_t* tail_r = tailw + 4 * (r - tail_r0) * nlq - e_left0;  // word of leftover elector e this round: tail_r[e]
                const bool sticky_live = have_prev && has_sticky;

                int cnt0 = 0;                                        // run-off (fast2): this lane's votes for ef0
                if (!have_eff) {
                  if constexpr (EXT) {
                    // =============== full-field round, EXT: ONE loop over the steps of the round ==========================
                    // (a step = NE quad-dealt electors of every lane, or up to 32 leftover electors P lanes each), so that the
                    // ~500-instruction per-elector tail (corrections + draw) exists ONCE, inline: with one copy per kind of step and
                    // five of the run-off body the hot code was ~80 KB and the kernel instruction-fetch bound (no_instruction 5.0 per
                    // issue, profiles/r02b_ncu_ext_v1_summary.txt); an out-of-line routine instead put the round's constants
                    // (LaneCtx32) into local memory (long_scoreboard 0.1 -> 2.7 per issue, v2).
                    constexpr int SPQ = 4 / NE;                                  // steps per quad iteration
                    const int n_main_steps = m_quads * SPQ;
                    const int n_steps = n_main_steps + ((n_left + 31) >> 5);
                    const float hbonus = 0.5f * kx.bonus;
                    const bool masked = n_thr > 0;                               // stop-candidate live this round: the masked pair loop
#pragma unroll 1
                    for (int step = 0; step < n_steps; ++step) {
                        int nd, e_first, pstride, wsel = 0;
                        const float* pbase;
                        bool act_draw = true;
                        uint32_t trigm = 0;                                      // bit i = draw i's elector is triggered
                        if (step < n_main_steps) {
                            const int qi = step / SPQ, g = step - qi * SPQ;
                            const int q = lane + 32 * qi;
                            if (g == 0) {
                                uint32_t w[4];
                                philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)q, a.b.seed_lo, a.b.seed_hi, w);
                                *reinterpret_cast<uint4*>(wq + 4 * lane) = make_uint4(w[0], w[1], w[2], w[3]);
                            }
                            const int e0 = 4 * q + NE * g;
                            float Re[NE], Ai[NE];
                            unsigned long long acc[NE];
#pragma unroll
                            for (int i = 0; i < NE; ++i) {
                                const float Ainv = __ldg(a.tab + (tabA + (uint32_t)(CP + e0 + i)));
                                Re[i] = Ainv * Ainv; Ai[i] = Ainv; acc[i] = 0ull;
                            }
                            const float* tc = kx.T;
                            float* pc = pref + lane;                             // [j][t][lane]
                            if (masked) {
                                float bmh[NE];
                                const uint2* abp[NE]; const uint2* rbq[NE];
                                unsigned long long accA[NE];
#pragma unroll
                                for (int i = 0; i < NE; ++i) {
                                    const int e = e0 + i;
                                    const bool tg = ext_triggered(thrbits, a.accmask + ((size_t)cur_set * E + e) * a.accw, Wb);
                                    trigm |= (tg ? 1u : 0u) << i;
                                    bmh[i] = tg ? bm1h : 0.0f;
                                    abp[i] = accb_s + ((qi << 7) + ((NE * g + i) << 5) + lane);     // dense32_slot(e, E)
                                    rbq[i] = regb + reg[e] * nch;
                                    accA[i] = 0ull;
                                }
                                const float* ph = phi; const float* bwc = bw;
                                const int epitch = a.epitch;
#pragma unroll 1
                                for (int j = 0; j < nch; ++j) {
                                    uint2 ab[NE], rb[NE];
#pragma unroll
                                    for (int i = 0; i < NE; ++i) { ab[i] = abp[i][j * epitch]; rb[i] = rbq[i][j]; }
                                    chunk_ext8<NE>(tc, ph, bwc, hbonus, ab, rb, Re, Ai, acc, accA);
#pragma unroll
                                    for (int i = 0; i < NE; ++i)
                                        pc[i * 32] = fmaf(bmh[i], f2_lo(accA[i]) + f2_hi(accA[i]), f2_lo(acc[i]) + f2_hi(acc[i]));
                                    tc += 20; ph += 12; bwc += 12; pc += NE * 32;
                                }
                            } else {
                                const ulonglong2* t4 = reinterpret_cast<const ulonglong2*>(tc);
                                ulonglong2 nb0 = t4[0], nc0 = t4[1], nb1 = t4[2], nc1 = t4[3];
#pragma unroll 2
                                for (int j = 0; j < nch; ++j) {
                                    const ulonglong2 b0 = nb0, c0 = nc0, b1 = nb1, c1 = nc1;
                                    t4 += 5;                                                     // 80 bytes per chunk
                                    nb0 = t4[0]; nc0 = t4[1]; nb1 = t4[2]; nc1 = t4[3];
                                    chunk_math8<NE>(b0, c0, b1, c1, Re, acc);
#pragma unroll
                                    for (int i = 0; i < NE; ++i) pc[i * 32] = f2_lo(acc[i]) + f2_hi(acc[i]);   // running prefix, m units
                                    pc += NE * 32;
                                }
                            }
                            nd = NE; e_first = e0; pbase = pref + lane; pstride = NE * 32; wsel = NE * g;
                        } else {
                            // leftover electors: P lanes per elector, a contiguous range of chunks each (see the non-EXT form below)
                            const int gb = (step - n_main_steps) << 5;
                            const int n_it = min(32, n_left - gb);
                            int P = 1;
                            while (P * 2 * n_it <= 32) P *= 2;
                            const int eli = lane / P, sub = lane - eli * P;
                            const bool act = eli < n_it;
                            const int e = e_left0 + gb + (act ? eli : 0);
                            const float Ainv = __ldg(a.tab + (tabA + (uint32_t)(CP + e)));
                            const float Re1[1] = {Ainv * Ainv}, Ai1[1] = {Ainv};
                            float* cl = pref + eli * nch;                        // this elector's chunk prefixes (m units)
                            const int per = (nch + P - 1) / P;
                            const int j_lo = min(nch, sub * per), j_hi = min(nch, j_lo + per);
                            float run = 0.f;
                            __syncwarp();
                            if (masked) {
                                const bool tg = ext_triggered(thrbits, a.accmask + ((size_t)cur_set * E + e) * a.accw, Wb);
                                trigm = tg ? 1u : 0u;
                                if (act) {
                                    const float bmh1 = tg ? bm1h : 0.0f;
                                    const uint2* abp = accb_s + e;               // dense32_slot(e, E) == e for the leftover electors
                                    const uint2* rbq = regb + reg[e] * nch;
#pragma unroll 1
                                    for (int j = j_lo; j < j_hi; ++j) {
                                        unsigned long long acc[1] = {0ull}, accA[1] = {0ull};
                                        const uint2 ab[1] = {abp[j * a.epitch]}, rb[1] = {rbq[j]};
                                        chunk_ext8<1>(kx.T + j * 20, phi + j * 12, bw + j * 12, hbonus, ab, rb, Re1, Ai1, acc, accA);
                                        run += fmaf(bmh1, f2_lo(accA[0]) + f2_hi(accA[0]), f2_lo(acc[0]) + f2_hi(acc[0]));
                                        cl[j] = run;
                                    }
                                }
                            } else if (act) {
                                const float xe1[1] = {0.f};
#pragma unroll 1
                                for (int j = j_lo; j < j_hi; ++j) {
                                    unsigned long long acc[1] = {0ull};
                                    chunk_acc<1, LC, FACT>(kx, kx.T + j * 20, xe1, Re1, acc);
                                    run += f2_lo(acc[0]) + f2_hi(acc[0]);
                                    cl[j] = run;
                                }
                            }
                            float off = run;                                     // inclusive scan over the P lanes of an elector, then exclusive
                            for (int d = 1; d < P; d <<= 1) {
                                const float o = __shfl_up_sync(CSIM_FULL, off, d);
                                if (sub >= d) off += o;
                            }
                            off -= run;
                            if (act && sub > 0) for (int j = j_lo; j < j_hi; ++j) cl[j] += off;
                            __syncwarp();
                            nd = 1; e_first = e; pbase = cl; pstride = 1; act_draw = act && sub == 0;
                        }
                        // ---- the per-elector tail: scalars, point-rule corrections, search + re-scan, tally
#pragma unroll 1
                        for (int i = 0; i < nd; ++i) {
                            if (act_draw) {
                                const int e = e_first + i;
                                const Elector32 el = load_elector(e, sticky_live, (int)((trigm >> i) & 1u));
                                const uint32_t wa = step < n_main_steps ? wq[4 * lane + wsel + i] : tail_r[e];
                                const Corr32 cr = corrections32_ext<LC, FACT>(kx, e, el);
                                const int vote = finish_draw<LC, MODEL, FACT, PROBE, true>(kx, pbase + i * 32, pstride, e, u24_from_word(wa), el, cr);
                                atomicAdd(&tally[vote], 1);
                                votes[e] = vote;                                 // read only by this lane for its own electors: in-place is safe
                            }
                        }
                        if (step >= n_main_steps) __syncwarp();
                    }
                  } else {
                    // =============== full-field round: NE electors of a lane at once =====================
                    for (int qi = 0; qi < m_quads; ++qi) {
                        const int q = lane + 32 * qi;
                        {   // this lane's four words go to its own slot of shared memory: nothing of them stays live across the pair loop
                            uint32_t w[4];
                            philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)q, a.b.seed_lo, a.b.seed_hi, w);
                            *reinterpret_cast<uint4*>(wq + 4 * lane) = make_uint4(w[0], w[1], w[2], w[3]);
                        }
#pragma unroll 1
                        for (int g = 0; g < 4 / NE; ++g) {
                            const int e0 = 4 * q + NE * g;
                            float xe[NE], Re[NE];
                            unsigned long long acc[NE];
#pragma unroll
                            for (int i = 0; i < NE; ++i) {          // all the pair loop needs of an elector: R_e (or its ideology)
                                if (FACT) { const float Ainv = __ldg(a.tab + (tabA + (uint32_t)(CP + e0 + i))); Re[i] = Ainv * Ainv; xe[i] = 0.f; }
                                else { Re[i] = 1.f; xe[i] = __ldg(a.ro.x32 + e0 + i); }
                                acc[i] = 0ull;
                            }
                            // phase 1: nothing but the pair arithmetic, its operand loads and one prefix store per elector
                            const float* tc = kx.T;
                            const int tstep = LC == 8 ? 20 : L;
                            float* pc = pref + lane;                                    // [j][t][lane]
                            if (LC == 8 && FACT && CSIM_DENSE_PIPELINE) {
                                // software-pipelined: chunk j + 1's operands are in flight while chunk j is evaluated (the load past the
                                // last chunk reads the next table / the warp's own arrays: inside the allocation, never used)
                                const ulonglong2* t4 = reinterpret_cast<const ulonglong2*>(tc);
                                ulonglong2 nb0 = t4[0], nc0 = t4[1], nb1 = t4[2], nc1 = t4[3];
#pragma unroll 2
                                for (int j = 0; j < nch; ++j) {
                                    const ulonglong2 b0 = nb0, c0 = nc0, b1 = nb1, c1 = nc1;
                                    t4 += 5;                                                     // 80 bytes per chunk
                                    nb0 = t4[0]; nc0 = t4[1]; nb1 = t4[2]; nc1 = t4[3];
                                    chunk_math8<NE>(b0, c0, b1, c1, Re, acc);
#pragma unroll
                                    for (int i = 0; i < NE; ++i) pc[i * 32] = f2_lo(acc[i]) + f2_hi(acc[i]);   // running prefix, m units
                                    pc += NE * 32;
                                }
                            } else {
#pragma unroll 2
                                for (int j = 0; j < nch; ++j) {
                                    chunk_acc<NE, LC, FACT>(kx, tc, xe, Re, acc);
#pragma unroll
                                    for (int i = 0; i < NE; ++i) pc[i * 32] = f2_lo(acc[i]) + f2_hi(acc[i]);   // running prefix, m units
                                    tc += tstep; pc += NE * 32;
                                }
                            }
                            // phase 2: the elector's scalars and point-rule corrections are only needed here
#pragma unroll 1
                            for (int i = 0; i < NE; ++i) {
                                const int e = e0 + i;
                                const Elector32 el = load_elector(e, sticky_live);
                                const Corr32 cr = corrections32<LC, MODEL, FACT>(kx, e, el);
                                const uint32_t wa = wq[4 * lane + NE * g + i];
                                const int vote = finish_draw<LC, MODEL, FACT, PROBE>(kx, pref + i * 32 + lane, NE * 32, e, u24_from_word(wa), el, cr);
                                atomicAdd(&tally[vote], 1);
                                if (MODEL) votes[e] = vote;        // read only by this lane for its own electors: in-place is safe
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
                        const Elector32 el = load_elector(e, sticky_live);
                        const float xe1[1] = {el.x}, Re1[1] = {el.R};
                        const int tstep = LC == 8 ? 20 : L;
                        float* cl = pref + eli * nch;                // this elector's chunk prefixes (m units)
                        // lane `sub` of an elector sums a contiguous range of its chunks (running prefix inside the range), the P range
                        // totals are scanned with shuffles, and every lane adds the total of the ranges before its own
                        const int per = (nch + P - 1) / P;
                        const int j_lo = min(nch, sub * per), j_hi = min(nch, j_lo + per);
                        float run = 0.f;
                        __syncwarp();
                        if (act) {
                            for (int j = j_lo; j < j_hi; ++j) {
                                unsigned long long acc[1] = {0ull};
                                chunk_acc<1, LC, FACT>(kx, kx.T + j * tstep, xe1, Re1, acc);
                                run += f2_lo(acc[0]) + f2_hi(acc[0]);
                                cl[j] = run;
                            }
                        }
                        float off = run;                             // inclusive scan over the P lanes of an elector, then exclusive
                        for (int d = 1; d < P; d <<= 1) {
                            const float o = __shfl_up_sync(CSIM_FULL, off, d);
                            if (sub >= d) off += o;
                        }
                        off -= run;
                        if (act && sub > 0) for (int j = j_lo; j < j_hi; ++j) cl[j] += off;
                        __syncwarp();
                        if (act && sub == 0) {
                            const Corr32 cr = corrections32<LC, MODEL, FACT>(kx, e, el);
                            const int vote = finish_draw<LC, MODEL, FACT, PROBE>(kx, cl, 1, e, u24_from_word(tail_r[e]), el, cr);
                            atomicAdd(&tally[vote], 1);
                            if (MODEL) votes[e] = vote;
                        }
                        __syncwarp();
                    }
                  }
                } else {
                    // =============== run-off round: first k columns (simulate.py:144-154) ==================
                    if (fast2) {
                        // k = 2: columns 0 and 1 are the same two candidates for every elector; their operands of this round, their
                        // regions and bandwagon terms live in registers, and a vote is decided in m units (score / A_e: the
                        // comparison v0 > u (v0 + v1) does not depend on the positive factor A_e), so an elector costs one load.
                        const int o0 = tab_off<LC>(0), o1 = tab_off<LC>(1);
                        float P0 = kx.T[o0], X0 = kx.T[o0 + kx.xoff], P1 = kx.T[o1], X1 = kx.T[o1 + kx.xoff];
                        const int g0 = reg[0], g1 = reg[1];
                        float f0 = 0.f, f1 = 0.f;
                        if (MODEL && kx.use_bw) { f0 = bw[pad_c<LC>(0)]; f1 = bw[pad_c<LC>(1)]; }
                        const float bonus = kx.bonus, nb = kx.nb, st1p = kx.st1p;
                        float bon0 = bonus, bon1 = bonus;
                        if (EXT) {                                   // fatigue scales the two candidates' whole scores (f0, f1 already carry it)
                            const float ph0 = phi[pad_c<LC>(0)], ph1 = phi[pad_c<LC>(1)];
                            P0 *= ph0; X0 *= ph0; P1 *= ph1; X1 *= ph1; bon0 *= ph0; bon1 *= ph1;
                        }
                        auto play = [&](int e, uint32_t wa) {
                            const float u = u24_from_word(wa);
                            const int g = reg[e];
                            const float b0 = (g == g0 ? bon0 : 0.0f) + f0, b1 = (g == g1 ? bon1 : 0.0f) + f1;
                            float w0, w1;
                            if (FACT) {
                                const float Ainv = __ldg(a.tab + (tabA + (uint32_t)(CP + e))), Rr = Ainv * Ainv;
                                w0 = fmaf(b0, Ainv, fminf(P0, Rr * X0));
                                w1 = fmaf(b1, Ainv, fminf(P1, Rr * X1));
                            } else {
                                const float xe = __ldg(a.ro.x32 + e);
                                w0 = fmaf(ex2_approx(fabsf(xe - X0) * nb), P0, b0);
                                w1 = fmaf(ex2_approx(fabsf(xe - X1) * nb), P1, b1);
                            }
                            if (MODEL && sticky_live) {
                                const int prev = votes[e];
                                if (prev == 0) w0 *= st1p;
                                if (prev == 1) w1 *= st1p;
                            }
                            if (EXT) {
                                if (n_thr > 0) {                     // a triggered elector boosts the candidates acceptable to it (model.py:400-415)
                                    const uint32_t* arow = a.accmask + ((size_t)cur_set * E + e) * a.accw;
                                    if (ext_triggered(thrbits, arow, Wb)) {
                                        const uint32_t aw = __ldg(arow);
                                        if (aw & 1u) w0 *= kx.boost;
                                        if (aw & 2u) w1 *= kx.boost;
                                    }
                                }
                            }
                            if (e == 0) w0 = 0.0f;
                            if (e == 1) w1 = 0.0f;
                            const float sum = w0 + w1;
                            int j;
                            if (!(sum > 0.0f)) j = (u * 2.0f >= 1.0f) ? 1 : 0;       // simulate.py:151 uniform fallback
                            else j = (w0 > u * sum || !(w1 > 0.0f)) ? 0 : 1;
                            cnt0 += (j == 0);
                            if (MODEL) votes[e] = j == 0 ? ef0 : ef1;
                        };
                        if constexpr (EXT && !CSIM_EXT_PLAY4) {
                            // ONE copy of the per-elector body (it is ~3x the plain one: the kernel's hot code has to stay inside the
                            // instruction cache, see the header of the EXT section in DESIGN.md)
                            uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll 1
                            for (int n = 0; n < n_slots; ++n) {
                                int e;
                                uint32_t wa;
                                if (n < n_main) {
                                    const int q = lane + 32 * (n >> 2);
                                    e = 4 * q + (n & 3);
                                    if ((n & 3) == 0) philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)q, a.b.seed_lo, a.b.seed_hi, w);
                                    const int sel = n & 3;
                                    wa = sel == 0 ? w[0] : (sel == 1 ? w[1] : (sel == 2 ? w[2] : w[3]));
                                } else {
                                    e = e_single0 + 32 * (n - n_main);
                                    wa = tail_r[e];
                                }
                                play(e, wa);
                            }
                        } else {
                            for (int qi = 0; qi < m_quads; ++qi) {
                                const int q = lane + 32 * qi;
                                uint32_t w[4];
                                philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)q, a.b.seed_lo, a.b.seed_hi, w);
#pragma unroll
                                for (int i = 0; i < 4; ++i) play(4 * q + i, w[i]);
                            }
                            for (int e = e_single0; e < E; e += 32) play(e, tail_r[e]);
                        }
                    } else {
                        uint32_t w[4] = {0, 0, 0, 0};
                        for (int n = 0; n < n_slots; ++n) {
                            int e, q; bool gen;
                            if (n < n_main) { q = lane + 32 * (n >> 2); e = 4 * q + (n & 3); gen = (n & 3) == 0; }
                            else { e = e_single0 + 32 * (n - n_main); q = e >> 2; gen = false; }
                            if (gen) philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)q, a.b.seed_lo, a.b.seed_hi, w);
                            const int sel = e & 3;
                            const uint32_t wa = n >= n_main ? tail_r[e] : (sel == 0 ? w[0] : (sel == 1 ? w[1] : (sel == 2 ? w[2] : w[3])));
                            const float u = u24_from_word(wa);
                            const Elector32 el = load_elector(e, sticky_live);
                            float sum = 0.f;
                            for (int j = 0; j < k; ++j) sum += full_score32<LC, MODEL, FACT, EXT>(kx, e, j, el);
                            int jsel;
                            if (!(sum > 0.0f)) {
                                jsel = (int)(u * (float)k);
                            } else {
                                const float tgt = u * sum;
                                float cum = 0.f; jsel = -1; int lastpos = 0;
                                for (int j = 0; j < k; ++j) {
                                    const float v = full_score32<LC, MODEL, FACT, EXT>(kx, e, j, el);
                                    cum += v;
                                    if (v > 0.0f) lastpos = j;
                                    if (jsel < 0 && v > 0.0f && cum > tgt) jsel = j;
                                }
                                if (jsel < 0) jsel = lastpos;
                            }
                            if (jsel >= k) jsel = k - 1;
                            const int vote = eff[jsel];
                            atomicAdd(&tally[vote], 1);
                            if (MODEL) votes[e] = vote;
                        }
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
                if (EXT) {
                    uint8_t* hr = hist + hrow * Ei;
#pragma unroll 1
                    for (int c = lane; c < C; c += 32) hr[c] = (uint8_t)tally[c];
                    ++n_hist; hrow = hrow + 1 == a.Fmax ? 0 : hrow + 1;
                }
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
                    if (EXT) prev_mx = mx;
                    if (mx >= need) { winner = arg; reason = CSIM_REASON_WINNER; break; }        // simulate.py:174-180
                    if (r >= rr) { ef0 = arg; ef1 = second; have_eff = true; }                   // simulate.py:183-186
                } else {
                    int mx, arg;
                    warp_argmax_first(tally, C, lane, mx, arg);
                    if (EXT) prev_mx = mx;
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
    }
    if (cur_set >= 0 && lane == 0 && a.out.totals && acc_trials) {
        atomicAdd(&a.out.totals[2 * cur_set], acc_rounds); atomicAdd(&a.out.totals[2 * cur_set + 1], acc_trials);
    }
}
