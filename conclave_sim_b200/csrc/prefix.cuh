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
//   1. S = F_e(last), target = u S, a binary search over the chunk prefixes (17 at E = 134: 5 probes; at most 128: 7 probes), and
//   2. a re-scan of the ONE chunk the target falls into, every term inline in roster order, to find the
//      first c with cum > target  (== searchsorted(cumsum(p)/sum, u, 'right'), simulate.py:153).
// That is the dense fp32 engine (dense32.cuh) with its phase 1 — the E x C exponentials every trial
// re-evaluates — replaced by a table lookup: O(log C + L) work per elector-round instead of O(C).  Under
// the ref_cli wiring the bandwagon / stickiness terms vanish and F_e(j) is one load.
//
// The tables of ALL rounds of one parameter set are small (E x 17 floats per round, 182 KB for 20 rounds
// at E = 134) and live in shared memory (1 CTA per SM, staged once per set by TMA bulk copies on an mbarrier), so — unlike the
// full-resolution CDF tables of table.cuh, which only fit one round at a time — trials stay independent:
// one warp owns one trial from round 1 to its end, then takes the next.  When they do not fit (cfg-5:
// 1024 electors x 50 rounds = 6.7 MB) the probes read them through L2 instead (STAGE = false).
//
// Table rows are stored in the order the lanes visit electors (prefix_slot) with an odd row pitch, so the
// 32 lanes of a warp probing the same chunk column hit 32 different banks.
#pragma once
#include "common.cuh"

#define CSIM_PREFIX_MAX_CHUNKS 128     // chunk prefixes per row (the bandwagon scan walks them 32 at a time)
#ifndef CSIM_PREFIX_THREADS
#define CSIM_PREFIX_THREADS 1024       // up to 32 warps per CTA, 1 CTA per SM (64 registers): measured 768 / 896 / 1024 threads =
                                       // 16.35 / 15.95 / 15.42 ms (ref_cli) and 20.4 / 20.06 / 20.03 ms (model) per 1.25 M trials
#endif

__host__ __device__ inline int prefix_cpp(int nch, int Ls);

struct PrefixArgs {
    RosterDev ro;
    BatchDev b;
    TrialOut out;
    const float* nbeta;     // [sets][R]  -beta_r * log2(e)
    const float* T;         // chunk-prefix tables of the full-field rounds: [sets][Rt][rows][stride] when staged in shared
                            // memory (odd row pitch: conflict-free), [sets][Rt][stride][rows] (column-major) when probed through
                            // L2: the 32 lanes of a pass own 32 consecutive rows, so a probe of one chunk column is one 128-byte line
    int col_major;          // 1: the second layout
    int Rt;                 // rounds that can be full-field: min(R, max(runoff_threshold_rounds, 1)) when a run-off exists, else R
    int Ls;                 // log2(candidates per chunk); L = 1 << Ls >= 8
    int nch;                // chunks, nch * L >= C, nch <= CSIM_PREFIX_MAX_CHUNKS
    int stride;             // row pitch in floats (odd, >= nch)
    int rows;               // rows per round
    int e_quad_end;         // electors below this are dealt as whole Philox quads per lane, the rest one per lane
    int kmax;
    int warps_per_cta;
    WorkQueue wq;           // dynamic trial scheduling (common.cuh)
    size_t set_pitch;       // floats between the tables of consecutive sets (16-byte multiple)
    size_t tab_smem;        // bytes of the staged tables (0 when STAGE = false)
    size_t cta_smem;        // staged tables + roster tables
    size_t warp_smem;       // per-warp state
    // shared-memory byte offsets, computed on the host (prefix_layout) so that the kernel reads them straight from the
    // constant bank instead of keeping a dozen addresses in registers:
    uint32_t o_xs, o_pw, o_rb, o_xg, o_nbs;            // from the start of dynamic shared memory
    uint32_t w_bwp, w_tally, w_eff, w_votes, w_mark;   // from the warp's block (bandwagon f32[CPP] sits at 0)
    uint32_t cp4;                                      // 4 * CPP: byte pitch of the per-candidate arrays
    int nch4, step0_4;                                 // 4 nch; 4 x (largest power of two <= nch)
    // PROBE instantiation only (csim_engine_probs): see Dense32Args
    int probe_round;
    const int32_t* probe_prev_vote;
    const int32_t* probe_prev_count;
    float* probe_scores;
    float* probe_prefix;
    int probe_cap;
};

// fills the offsets above from (E, G, R, Ls, nch, rows, kmax, tab_smem); returns nothing the caller has to size by hand
__host__ inline void prefix_layout(PrefixArgs& a, bool model) {
    const uint32_t CPP = (uint32_t)prefix_cpp(a.nch, a.Ls), G = (uint32_t)a.ro.G;
    const uint32_t Ei = ((uint32_t)a.ro.E + 7u) & ~7u, rows8 = ((uint32_t)a.rows + 7u) & ~7u;
    a.cp4 = 4u * CPP;
    uint32_t o = (uint32_t)a.tab_smem;
    a.o_xs = o; o += a.cp4;
    a.o_pw = o; o += a.cp4;
    a.o_rb = o; o += a.cp4 * G;
    a.o_xg = o; o += 8u * rows8;
    a.o_nbs = o;
    uint32_t w = 0;
    if (model) { w += a.cp4; a.w_bwp = w; w += 4u * CSIM_PREFIX_MAX_CHUNKS; } else a.w_bwp = 0;
    a.w_tally = w; w += 4u * Ei;
    a.w_eff = w; w += 4u * (((uint32_t)a.kmax + 3u) & ~3u);
    a.w_votes = w; if (model) w += 2u * rows8;
    a.w_mark = w;
    a.nch4 = 4 * a.nch;
    a.step0_4 = 4;
    while (a.step0_4 * 2 <= a.nch4) a.step0_4 *= 2;
}

// Row of elector e inside one round's table: quad-dealt electors (e = 4 (lane + 32 qi) + i) sit at
// 128 qi + 32 i + lane, so one pass of the warp (fixed i) touches 32 consecutive rows.
__host__ __device__ inline int prefix_slot(int e, int e_quad_end) {
    return e < e_quad_end ? ((e >> 7) << 7) + ((e & 3) << 5) + ((e >> 2) & 31) : e;
}
// Per-candidate arrays are stored with 4 floats of padding after every chunk of L: a lane reads its chunk with
// 128-bit loads, and with a pitch of L + 4 floats the chunks start in 8 different 4-bank groups instead of 4 (L = 8).
__host__ __device__ inline int prefix_cpp(int nch, int Ls) { return nch * ((1 << Ls) + 4); }
__host__ __device__ inline size_t prefix_roster_smem(int CPP, int G, int R, int rows) {
    const size_t rows8 = ((size_t)rows + 7) & ~(size_t)7;
    return ((size_t)CPP * 8 /*xs, pw*/ + (size_t)CPP * G * 4 /*rb*/ + rows8 * 8 /*xg*/ + (size_t)R * 4 /*nbeta*/ + 31) & ~(size_t)15;
}
__host__ __device__ inline size_t prefix_warp_smem(int E, int CPP, int kmax, int rows, bool model) {
    const size_t Ei = ((size_t)E + 7) & ~(size_t)7, rows8 = ((size_t)rows + 7) & ~(size_t)7;
    size_t n = Ei * 4 /*tally*/ + (((size_t)kmax + 3) & ~(size_t)3) * 4 /*eff*/ + Ei /*mark*/;
    if (model) n += (size_t)CPP * 4 /*bw*/ + CSIM_PREFIX_MAX_CHUNKS * 4 /*bwp*/ + rows8 * 2 /*votes*/;
    return (n + 15) & ~(size_t)15;
}

// Host-side plan of one launch: how electors are dealt to lanes, the chunk size, whether the tables are staged in
// shared memory, warps per CTA and every shared-memory offset.  Pure arithmetic (no CUDA calls): csim_prefix_plan_query
// exposes it so that the CPU test-suite can check the budgeting for any roster shape.
//   budget       = opt-in shared memory per block of the device (232448 B on sm_100)
//   allow_stage  = false forces the global-memory (L2) path
// Returns false when even one warp does not fit.
//   force_Ls     > 0: use chunks of 1 << force_Ls candidates (tuning / tests), else the rule below
__host__ inline bool prefix_plan(PrefixArgs& a, int E, int G, int R, int Rt, int kmax, bool model, size_t budget, bool allow_stage,
                                 int force_Ls = 0) {
    a.ro.E = E; a.ro.G = G; a.kmax = kmax;
    // electors are dealt to lanes as whole Philox quads (128 per group of 4 passes); a tail of <= 64 electors goes one per lane
    const int n_left = E & 127;
    a.e_quad_end = (E & ~127) + (n_left > 64 ? 128 : 0);
    a.rows = E > a.e_quad_end ? E : a.e_quad_end;
    a.Rt = Rt < 1 ? 1 : Rt;
    const int max_wpc = CSIM_PREFIX_THREADS / 32;
    int Ls_min = 3;
    while (((E + (1 << Ls_min) - 1) >> Ls_min) > CSIM_PREFIX_MAX_CHUNKS) ++Ls_min;
    // tables probed through L1 / L2 (large rosters): re-scanning a chunk costs ~14 instructions per candidate, a probe one
    // cache round trip.  Measured (100 k trials, elector-rounds/s, chunks of 8 / 16 / 32, best layout below):
    //   E = 512   ref_cli 1.46e11 / 1.34e11 / 0.93e11     model 9.3e10 / 8.3e10 / 6.9e10
    //   E = 1024  ref_cli 8.5e10 / 7.9e10 / 6.7e10        model 3.5e10 / 3.8e10 / 3.3e10
    const int Ls_want = (model && E > 512) ? 4 : 3;
    int Ls_pref = Ls_min > Ls_want ? Ls_min : Ls_want;
    if (force_Ls >= 3) { Ls_pref = force_Ls < Ls_min ? Ls_min : force_Ls; Ls_min = Ls_pref; }
    int Ls = -1, wpc = 0;
    bool stage = false;
    // smallest chunk (fewest re-scanned candidates) whose tables fit in shared memory next to >= 8 warps; a chunk of more
    // than 16 candidates costs more to re-scan than probing the tables through L2 does
    for (int l = Ls_min; l <= (force_Ls >= 3 ? Ls_min : 4) && allow_stage; ++l) {
        const int nch = (E + (1 << l) - 1) >> l, stride = nch | 1, CPP = prefix_cpp(nch, l);
        const size_t tab = (((size_t)a.Rt * a.rows * stride * 4) + 15) & ~(size_t)15;
        const size_t fixed = tab + prefix_roster_smem(CPP, G, R, a.rows), per_warp = prefix_warp_smem(E, CPP, kmax, a.rows, model);
        if (fixed + 8 * per_warp > budget) continue;
        Ls = l; stage = true;
        const size_t fit = (budget - fixed) / per_warp;
        wpc = (int)(fit < (size_t)max_wpc ? fit : (size_t)max_wpc);
        break;
    }
    if (!stage) {
        Ls = Ls_pref;
        const int nch = (E + (1 << Ls) - 1) >> Ls, CPP = prefix_cpp(nch, Ls);
        const size_t fixed = prefix_roster_smem(CPP, G, R, a.rows), per_warp = prefix_warp_smem(E, CPP, kmax, a.rows, model);
        if (fixed + per_warp > budget) return false;
        const size_t fit = (budget - fixed) / per_warp;
        wpc = (int)(fit < (size_t)max_wpc ? fit : (size_t)max_wpc);
    }
    a.Ls = Ls; a.nch = (E + (1 << Ls) - 1) >> Ls; a.stride = a.nch | 1;
    const int CPP = prefix_cpp(a.nch, Ls);
    a.set_pitch = (((size_t)a.Rt * a.rows * a.stride) + 3) & ~(size_t)3;
    a.tab_smem = stage ? a.set_pitch * 4 : 0;
    a.cta_smem = a.tab_smem + prefix_roster_smem(CPP, G, R, a.rows);
    a.warp_smem = prefix_warp_smem(E, CPP, kmax, a.rows, model);
    a.warps_per_cta = wpc;
    // layout of a table that stays in global memory: row-major keeps the 6-7 probes of one draw inside the 2-3 lines of its
    // row — best while a round's table fits in what the shared-memory carve-out leaves of L1 (E = 512 ref_cli: 1.46e11 vs
    // 1.32e11); beyond that column-major wins: a pass of 32 lanes owns 32 consecutive rows, so a probe of one chunk column
    // is one 128-byte line (E = 1024 ref_cli: 8.5e10 vs 6.0e10; E = 512 model, where shared memory leaves little L1: 9.3e10 vs 5.1e10)
    const size_t l1_left = (size_t)228 * 1024 > a.cta_smem + a.warp_smem * (size_t)wpc ? (size_t)228 * 1024 - (a.cta_smem + a.warp_smem * (size_t)wpc) : 0;
    a.col_major = (!stage && (size_t)a.rows * a.stride * 4 > l1_left) ? 1 : 0;
    prefix_layout(a, model);
    return true;
}

// K0: chunk-prefix tables.  One warp per (set, round, elector), lane j sums chunk j in fp64 (the order of the
// terms is model.py:311-324: exp, papabile factor, regional bonus), an inclusive warp scan makes the prefixes.
// grid = (ceil(E / 8), sets * Rt), block = 256.
static __global__ void k_prefix_tables32(RosterDev ro, const DevSet* sets, const double* betas, int R, int Rt, int Ls, int nch, int stride,
                                  int rows, int e_quad_end, size_t set_pitch, int col_major, float* T, int sr0) {
    const int lane = threadIdx.x & 31, e = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int sr = sr0 + blockIdx.y;
    const int set = sr / Rt, r0 = sr - set * Rt;                        // r0 = round - 1
    const int C = ro.E;
    if (e >= C) return;
    const DevSet* S = &sets[set];
    const double beta = betas[(size_t)set * R + r0], xe = ro.x64[e], pwf = S->pwf, bonus = S->bonus;
    const int ge = ro.region[e];
    const int slot = prefix_slot(e, e_quad_end);
    float* tab = T + (size_t)set * set_pitch + (size_t)r0 * rows * stride;        // this round's table
    float* row = tab + (size_t)slot * stride;                                      // row-major
    double carry = 0.0;                                              // prefix up to the previous group of 32 chunks
    for (int j0 = 0; j0 < stride; j0 += 32) {
        const int j = j0 + lane;
        double s = 0.0;
        if (j < nch) {
            const int c1 = min(C, (j + 1) << Ls);
            for (int c = j << Ls; c < c1; ++c) {
                if (c == e) continue;                                        // model.py:418-424
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
        s += carry;
        if (j < stride) {                                            // entries beyond nch repeat the total (row pitch padding)
            if (col_major) tab[(size_t)j * rows + slot] = (float)s; else row[j] = (float)s;
        }
        carry = __shfl_sync(CSIM_FULL, s, 31);
    }
}

// ---- shared memory through 32-bit window addresses ---------------------------------------------------
// The draw touches ~10 shared arrays; as generic pointers they cost 20 registers and the compiler
// re-derives them inside the hot loop.  One 32-bit address per array and explicit ld.shared keeps the
// draw in registers.  (`asm volatile`: never merged or hoisted across a round; ptxas still schedules them.)
__device__ __forceinline__ float lds_f(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_i(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_u16(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return (int)v; }
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void lds_xg(uint32_t a, float& x, int& g) { asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=f"(x), "=r"(g) : "r"(a)); }
__device__ __forceinline__ void sts_u16(uint32_t a, int v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ void red_inc(uint32_t a) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); }

// Shared arrays of a draw, as window addresses = one of two bases + an offset from the constant bank:
//   CTA  (sb): xs f32[CPP] ideology, pw f32[CPP] papabile weight (0 = padding), rb f32[G][CPP] regional bonus of region g
//              (all in padded candidate order, prefix_cpp), xg {ideology, region} of the electors in slot order (8 B each)
//   warp (wb): bw f32[CPP] bandwagon, bwp f32[32] its prefix over chunks, tally i32[E], votes u16[rows] (slot order)
struct PrefixCtx {
    uint32_t sb, wb;
    float nb, st1p;
};
#define PFX_XS(k, a)    ((k).sb + (a).o_xs)
#define PFX_PW(k, a)    ((k).sb + (a).o_pw)
#define PFX_RB(k, a)    ((k).sb + (a).o_rb)
#define PFX_XG(k, a)    ((k).sb + (a).o_xg)
#define PFX_BW(k, a)    ((k).wb)
#define PFX_BWP(k, a)   ((k).wb + (a).w_bwp)
#define PFX_TALLY(k, a) ((k).wb + (a).w_tally)
#define PFX_VOTES(k, a) ((k).wb + (a).w_votes)

// every term of one score, inline (run-off rounds; c is warp-uniform)
template <bool MODEL>
__device__ __forceinline__ float prefix_full_score(const PrefixCtx& k, const PrefixArgs& a, int e, int c, float xe, int ge, int prev) {
    const uint32_t o = 4u * (uint32_t)(c + ((c >> a.Ls) << 2));
    float v = fmaf(ex2_approx(fabsf(xe - lds_f(PFX_XS(k, a) + o)) * k.nb), lds_f(PFX_PW(k, a) + o), lds_f(PFX_RB(k, a) + ge * a.cp4 + o));
    if (MODEL) {
        v += lds_f(PFX_BW(k, a) + o);
        if (c == prev) v *= k.st1p;
    }
    return c == e ? 0.0f : v;
}

// The 8 re-scan scores of the block of 8 candidates starting at cb (padded offset o), every term inline in roster order.
template <bool MODEL>
__device__ __forceinline__ void prefix_scores8(const PrefixCtx& k, const PrefixArgs& a, uint32_t rbrow, float xe, float nb, int e, int prev, int cb,
                                               uint32_t o, float (&v)[8]) {
    const float4 xa = lds_f4(PFX_XS(k, a) + o), xb = lds_f4(PFX_XS(k, a) + o + 16);
    const float4 pa = lds_f4(PFX_PW(k, a) + o), pb = lds_f4(PFX_PW(k, a) + o + 16);     // 0 beyond C: padding scores nothing
    const float4 ra = lds_f4(rbrow + o), rb = lds_f4(rbrow + o + 16);        // regional bonus of e's region
    float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
    if (MODEL) { ba = lds_f4(PFX_BW(k, a) + o); bb = lds_f4(PFX_BW(k, a) + o + 16); }
    const int is = e - cb;                               // position of e itself / of prev inside this block (else outside 0..7)
    const int ip = MODEL ? prev - cb : -1;
#define CSIM_PRESCAN(i, xv, pv, rv, bv)                                                  \
    {                                                                                \
        float s = fmaf(ex2_approx(fabsf(xe - (xv)) * nb), (pv), (rv));               \
        if (MODEL) { s += (bv); if (ip == (i)) s *= k.st1p; }                        \
        if (is == (i)) s = 0.0f;                                                     \
        v[i] = s;                                                                    \
    }
    CSIM_PRESCAN(0, xa.x, pa.x, ra.x, ba.x) CSIM_PRESCAN(1, xa.y, pa.y, ra.y, ba.y)
    CSIM_PRESCAN(2, xa.z, pa.z, ra.z, ba.z) CSIM_PRESCAN(3, xa.w, pa.w, ra.w, ba.w)
    CSIM_PRESCAN(4, xb.x, pb.x, rb.x, bb.x) CSIM_PRESCAN(5, xb.y, pb.y, rb.y, bb.y)
    CSIM_PRESCAN(6, xb.z, pb.z, rb.z, bb.z) CSIM_PRESCAN(7, xb.w, pb.w, rb.w, bb.w)
#undef CSIM_PRESCAN
}

// One vote of a full-field round.  STAGE: trow = shared address of this elector's row of the round's chunk-prefix
// table; else Tg = its first entry in the global table (row- or column-major, PrefixArgs::col_major).
template <bool LC8, bool MODEL, bool STAGE, bool PROBE = false>
__device__ __forceinline__ int prefix_draw(const PrefixCtx& k, const PrefixArgs& a, uint32_t trow, const float* __restrict__ Tg, int e, int slot,
                                           float u, int prev) {
    const int Ls = LC8 ? 3 : a.Ls, nch4 = a.nch4, C = a.ro.E;
    const uint32_t gstep = a.col_major ? (uint32_t)a.rows : 1u;       // column-major: consecutive chunks are `rows` floats apart
    float xe; int ge;
    lds_xg(PFX_XG(k, a) + 8 * slot, xe, ge);
    const uint32_t rbrow = PFX_RB(k, a) + ge * a.cp4;
    const float nb = k.nb;
    float selfbw = 0.f, sadd = 0.f;
    const int je4 = (e >> Ls) << 2;
    int jp4 = 0x7fffffff;
    if (MODEL) {
        selfbw = lds_f(PFX_BW(k, a) + 4 * (e + je4));        // bwp counts e's own bandwagon term; its score is 0 (model.py:418-424)
        if (prev >= 0 && prev != e) {                       // stickiness, model.py:357-367
            jp4 = (prev >> Ls) << 2;
            const uint32_t o = 4u * (uint32_t)(prev + jp4);
            float b = fmaf(ex2_approx(fabsf(xe - lds_f(PFX_XS(k, a) + o)) * nb), lds_f(PFX_PW(k, a) + o), lds_f(rbrow + o));
            b += lds_f(PFX_BW(k, a) + o);
            sadd = b * (k.st1p - 1.0f);
        }
    }
    auto F = [&](int j4) -> float {                         // prefix of the row over chunks 0..j  (j4 = 4 j)
        // staged: this elector's row in shared memory; else Tg = its first entry in the global table, gstep floats per chunk
        float v = STAGE ? lds_f(trow + j4) : __ldg(reinterpret_cast<const float*>(reinterpret_cast<const char*>(Tg) + (size_t)j4 * gstep));
        if (MODEL) {
            v += lds_f(PFX_BWP(k, a) + j4);
            v = (j4 >= je4) ? v - selfbw : v;
            v = (j4 >= jp4) ? v + sadd : v;
        }
        return v;
    };
    if (PROBE) {                  // csim_engine_probs: the numbers this draw is made from
        for (int j4 = 0; j4 < nch4; j4 += 4) a.probe_prefix[(size_t)e * a.probe_cap + (j4 >> 2)] = F(j4);
        for (int cb = 0; cb < C; cb += 8) {
            float v[8];
            prefix_scores8<MODEL>(k, a, rbrow, xe, nb, e, prev, cb, 4u * (uint32_t)cb + 4u * (uint32_t)((cb >> Ls) << 2), v);
#pragma unroll
            for (int i = 0; i < 8; ++i) if (cb + i < C) a.probe_scores[(size_t)e * C + cb + i] = v[i];
        }
    }
    const float S = F(nch4 - 4);
    if (!(S > 0.0f)) {            // all scores vanished: the reference makes the row uniform INCLUDING self (model.py:447-449)
        const int j = (int)(u * (float)C);
        return j < C ? j : C - 1;
    }
    const float tgt = u * S;
    // Bound-free search for the number of chunks whose prefix is <= tgt, in [0, nch - 1] (the last prefix is S itself; should
    // u S round up to S the vote belongs to the last chunk).  With P = the largest power of two <= nch, one first probe picks
    // between the overlapping windows [0, P) and [nch - P, nch); the halving steps then never leave the window: no clamps.
    int pos4 = 0;                 // 4 x number of chunks whose prefix is <= tgt
    float cum = 0.0f;             // prefix below the chunk the target falls into
    const int P4 = a.step0_4;
    if (nch4 > P4) {
        const float f = F(nch4 - P4 - 4);
        const bool le = f <= tgt;
        pos4 = le ? nch4 - P4 : 0;
        cum = le ? f : 0.0f;
    }
    for (int st = P4 >> 1; st >= 4; st >>= 1) {
        const float f = F(pos4 + st - 4);
        const bool le = f <= tgt;
        pos4 = le ? pos4 + st : pos4;
        cum = le ? f : cum;
    }
    const int c0 = (pos4 >> 2) << Ls;
    int cnt = 0;
    const int nb8 = LC8 ? 1 : (1 << (Ls - 3));
    for (int bi = 0; bi < nb8; ++bi) {                      // blocks of 8 candidates of the chunk, in roster order
        const int cb = c0 + 8 * bi;
        const uint32_t o = 4u * (uint32_t)cb + 4u * (uint32_t)pos4;     // padded order: 4 extra floats per chunk before this one
        float v[8];
        prefix_scores8<MODEL>(k, a, rbrow, xe, nb, e, prev, cb, o, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) { cum += v[i]; cnt += (cum <= tgt) ? 1 : 0; }
        if (cnt < 8 * (bi + 1)) break;                       // the crossing is inside this block
    }
    int vote = c0 + cnt;                                     // == first c with cum > tgt
    const int maxc = min(c0 + (1 << Ls) - 1, C - 1);
    vote = vote > maxc ? maxc : vote;                        // rounding at the chunk's upper edge
    if (vote == e) vote = e > 0 ? e - 1 : (C > 1 ? 1 : 0);
    return vote;
}

// One thread: issue the set's tables as bulk copies (TMA 1-D, UBLKCP) of <= 32 KB, all completing on the CTA's mbarrier (armed
// once at kernel start; every staging is one phase of it and the waiters track the parity).  Kept out of line so that its
// registers do not weigh on the trial loop.
static __device__ __noinline__ void prefix_stage_tables(uint32_t dst, const float* src, uint32_t total, uint32_t bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // earlier generic-proxy accesses of dst are ordered before the async writes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
    const char* gbase = reinterpret_cast<const char*>(src);
    for (uint32_t off = 0; off < total; off += 32768u) {
        const uint32_t nbytes = min(32768u, total - off);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + off), "l"(gbase + off), "r"(nbytes), "r"(bar) : "memory");
    }
}

template <bool LC8, bool MODEL, bool STAGE, bool PROBE = false>
__global__ void __launch_bounds__(CSIM_PREFIX_THREADS, 1) k_trials_prefix(const __grid_constant__ PrefixArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    asm volatile("" : "+r"(lane));       // opaque: kept in a register instead of being re-derived from %tid in every pass
    const int E = a.ro.E, C = E, R = a.b.R, G = a.ro.G;
    const int Ls = LC8 ? 3 : a.Ls, nch = a.nch, Lp = (1 << Ls) + 4, CPP = nch * Lp;
    auto pc = [&](int c) { return c + ((c >> Ls) << 2); };   // index of candidate c in the padded arrays
    // ---- CTA-shared: staged tables of the current set, then roster tables (offsets: prefix_layout) ---------
    float* xs = (float*)(smem_raw + a.o_xs);
    float* pwv = (float*)(smem_raw + a.o_pw);                                       // papabile weight of the current set (0 = padding)
    float* rbv = (float*)(smem_raw + a.o_rb);                                       // [g][c]: bonus if region_c == g else 0
    int2* xg = (int2*)(smem_raw + a.o_xg);                                          // {ideology bits, region} by elector slot
    float* nbs = (float*)(smem_raw + a.o_nbs);                                      // [R] -beta_r log2 e of the current set
    for (int i = threadIdx.x; i < CPP; i += blockDim.x) { xs[i] = 0.0f; pwv[i] = 0.0f; }
    for (int i = threadIdx.x; i < CPP * G; i += blockDim.x) rbv[i] = 0.0f;
    __shared__ __align__(8) unsigned long long stage_bar;                   // mbarrier of the staged tables: one phase per staged set
    uint32_t stage_phase = 0;
    if (STAGE && threadIdx.x == 0) mbar_init((uint32_t)__cvta_generic_to_shared(&stage_bar), 1);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        xs[pc(c)] = a.ro.x32[c];
        xg[prefix_slot(c, a.e_quad_end)] = make_int2(__float_as_int(a.ro.x32[c]), a.ro.region[c]);
    }
    // ---- per-warp state ----------------------------------------------------------------------------
    unsigned char* base = smem_raw + a.cta_smem + (size_t)wib * a.warp_smem;
    float* bw = (float*)base;                                                       // MODEL only
    float* bwp = (float*)(base + a.w_bwp);                                          // MODEL only
    int32_t* tally = (int32_t*)(base + a.w_tally);
    int32_t* eff = (int32_t*)(base + a.w_eff);
    uint8_t* mark = base + a.w_mark;

    const bool top2_ok = C >= 2 && C < 32768;
    const size_t round_pitch = (size_t)a.rows * a.stride;
    const uint32_t stride4 = 4u * (uint32_t)a.stride;

    PrefixCtx kx;
    kx.sb = (uint32_t)__cvta_generic_to_shared(smem_raw);
    kx.wb = kx.sb + (uint32_t)a.cta_smem + (uint32_t)wib * (uint32_t)a.warp_smem;
    asm volatile("" : "+r"(kx.sb), "+r"(kx.wb));   // opaque for the same reason (the shared-window base needs an S2R)
    kx.nb = 0.f; kx.st1p = 1.f;
    const uint32_t Ts_a = kx.sb;

    // the electors of one round, each lane with its uniform (common.cuh: deal_electors)
    auto for_each_elector = [&](int r, uint32_t s_lo, uint32_t s_hi, auto&& fn) {
        deal_electors(lane, E, a.e_quad_end, r, s_lo, s_hi, a.b.seed_lo, a.b.seed_hi, fn);
    };

    int cur_set = -1;
    float bw_strength = 0.f, bonus_s = 0.f;
    int need = 0, k = 0, rr = 0;
    bool has_sticky = false, has_bw = false;
    unsigned long long acc_rounds = 0, acc_trials = 0;
    const float* Tset = a.T;

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
            const DevSet* S = &a.b.sets[set];
            kx.st1p = S->f_st1p; bw_strength = S->f_bandwagon; bonus_s = S->f_bonus;
            need = S->need; k = S->k; rr = S->rr;
            has_sticky = MODEL && S->has_sticky; has_bw = MODEL && S->has_bw;
            const float pwf = S->f_pwf, bonus = S->f_bonus;
            Tset = a.T + (size_t)set * a.set_pitch;
            __syncthreads();                                    // readers of the previous set's tables are done
            if (STAGE && threadIdx.x == 0)
                prefix_stage_tables(Ts_a, Tset, (uint32_t)a.tab_smem, (uint32_t)__cvta_generic_to_shared(&stage_bar));
            for (int r = threadIdx.x; r < R; r += blockDim.x) nbs[r] = a.nbeta[(size_t)set * R + r];
            for (int c = threadIdx.x; c < C; c += blockDim.x) {
                pwv[pc(c)] = a.ro.pap[c] ? pwf : 1.0f;
                rbv[a.ro.region[c] * CPP + pc(c)] = bonus;
            }
            if (STAGE) { mbar_wait((uint32_t)__cvta_generic_to_shared(&stage_bar), stage_phase); stage_phase ^= 1u; }
            __syncthreads();                                    // tables landed, roster arrays of the set written
        }
        for (uint64_t i0 = warp_claim_trials(a.wq, set, lane); i0 < a.b.n_sims; i0 = warp_claim_trials(a.wq, set, lane))
        for (int ti = 0; ti < a.wq.chunk; ++ti) {
            const uint64_t i = i0 + (uint64_t)ti;
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
                for (int c = lane; c < CPP; c += 32) bw[c] = 0.0f;
                for (int j = lane; j < nch; j += 32) bwp[j] = 0.0f;
            }
            __syncwarp();
            if (PROBE && MODEL && (a.probe_prev_vote || a.probe_prev_count)) {   // start from the given previous-round state
                int m = 0;
                for (int c = lane; c < C; c += 32) {
                    const int v = a.probe_prev_count ? a.probe_prev_count[c] : 0;
                    tally[c] = v; m = max(m, v);
                    sts_u16(PFX_VOTES(kx, a) + 2 * prefix_slot(c, a.e_quad_end), a.probe_prev_vote ? a.probe_prev_vote[c] : 0xFFFF);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(CSIM_FULL, m, off));
                prev_mx = m; have_prev = true;
                __syncwarp();
            }

            for (int r = PROBE ? a.probe_round : 1; r <= R; ++r) {
                rounds = r;
                kx.nb = nbs[r - 1];
                const int rt = min(r, a.Rt) - 1;                     // rounds beyond Rt are run-off rounds: no table
                const uint32_t Tr_a = Ts_a + (uint32_t)((size_t)rt * round_pitch * 4);
                const float* Tr_g = Tset + (size_t)rt * round_pitch;
                float rb0 = 0.0f, rb1 = 0.0f;                        // bandwagon of candidates 0 / 1 (a k = 2 run-off round needs no more)
                if (MODEL && has_bw && have_prev && prev_mx > 0 && have_eff && fast2) {
                    const float inv = 1.0f / (float)prev_mx;
                    rb0 = (float)tally[0] * inv * bw_strength; rb1 = (float)tally[1] * inv * bw_strength;
                } else if (MODEL && has_bw && have_prev && prev_mx > 0) {        // bandwagon, model.py:326-354
                    const float inv = 1.0f / (float)prev_mx;
                    for (int c = lane; c < C; c += 32) bw[pc(c)] = (float)tally[c] * inv * bw_strength;   // tally = previous round's
                    __syncwarp();
                    float carry = 0.f;                                           // bandwagon per chunk, then prefix over chunks (32 at a time)
#pragma unroll 1
                    for (int j0 = 0; j0 < nch; j0 += 32) {
                        const int j = j0 + lane;
                        float s = 0.f;
                        if (j < nch) {
                            const float4* b4 = reinterpret_cast<const float4*>(bw + j * Lp);
                            for (int i = 0; i < (1 << (Ls - 2)); ++i) { const float4 v = b4[i]; s += (v.x + v.y) + (v.z + v.w); }
                        }
#pragma unroll
                        for (int off = 1; off < 32; off <<= 1) {
                            const float o = __shfl_up_sync(CSIM_FULL, s, off);
                            if (lane >= off) s += o;
                        }
                        s += carry;
                        if (j < nch) bwp[j] = s;
                        carry = __shfl_sync(CSIM_FULL, s, 31);
                    }
                }
                __syncwarp();
                for (int c = lane; c < C; c += 32) tally[c] = 0;
                __syncwarp();
                const bool sticky_live = MODEL && have_prev && has_sticky;

                int cnt0 = 0;                                        // run-off (fast2): this lane's votes for ef0
                if (!have_eff) {
                    // =============== full-field round ==============================================
                    for_each_elector(r, s_lo, s_hi, [&](int e, float u, int slot) {
                        int prev = sticky_live ? lds_u16(PFX_VOTES(kx, a) + 2 * slot) : -1;
                        if (PROBE && prev == 0xFFFF) prev = -1;              // probe input: this elector had no previous vote
                        const int vote = prefix_draw<LC8, MODEL, STAGE, PROBE>(kx, a, Tr_a + (uint32_t)slot * stride4, Tr_g + (a.col_major ? (size_t)slot : (size_t)slot * a.stride), e, slot, u, prev);
                        red_inc(PFX_TALLY(kx, a) + 4 * vote);
                        if (MODEL) sts_u16(PFX_VOTES(kx, a) + 2 * slot, vote);   // read back only by this lane: in place is safe
                    });
                } else {
                    // =============== run-off round: first k columns (simulate.py:144-154) ===========
                    if (fast2) {
                        // k = 2: columns 0 and 1 are the same two candidates for every elector: their constants live in registers
                        const uint32_t o1 = 4u * (uint32_t)pc(1);
                        const float x0 = lds_f(PFX_XS(kx, a)), x1 = lds_f(PFX_XS(kx, a) + o1);
                        const float p0 = lds_f(PFX_PW(kx, a)), p1 = lds_f(PFX_PW(kx, a) + o1);
                        const float b0 = rb0, b1 = rb1;
                        float xt; int g0, g1;
                        lds_xg(PFX_XG(kx, a) + 8 * prefix_slot(0, a.e_quad_end), xt, g0);
                        lds_xg(PFX_XG(kx, a) + 8 * prefix_slot(1, a.e_quad_end), xt, g1);
                        const float nb = kx.nb;
                        for_each_elector(r, s_lo, s_hi, [&](int e, float u, int slot) {
                            float xe; int ge;
                            lds_xg(PFX_XG(kx, a) + 8 * slot, xe, ge);
                            const int prev = sticky_live ? lds_u16(PFX_VOTES(kx, a) + 2 * slot) : -1;
                            float v0 = fmaf(ex2_approx(fabsf(xe - x0) * nb), p0, ge == g0 ? bonus_s : 0.0f) + b0;
                            float v1 = fmaf(ex2_approx(fabsf(xe - x1) * nb), p1, ge == g1 ? bonus_s : 0.0f) + b1;
                            if (MODEL) { if (prev == 0) v0 *= kx.st1p; if (prev == 1) v1 *= kx.st1p; }
                            if (e == 0) v0 = 0.0f;
                            if (e == 1) v1 = 0.0f;
                            const float sum = v0 + v1;
                            int j;
                            if (!(sum > 0.0f)) j = (u * 2.0f >= 1.0f) ? 1 : 0;       // simulate.py:151 uniform fallback
                            else j = (v0 > u * sum || !(v1 > 0.0f)) ? 0 : 1;
                            cnt0 += (j == 0);
                            if (MODEL) sts_u16(PFX_VOTES(kx, a) + 2 * slot, j == 0 ? ef0 : ef1);
                        });
                    } else {
                        for_each_elector(r, s_lo, s_hi, [&](int e, float u, int slot) {
                            float xe; int ge;
                            lds_xg(PFX_XG(kx, a) + 8 * slot, xe, ge);
                            const int prev = sticky_live ? lds_u16(PFX_VOTES(kx, a) + 2 * slot) : -1;
                            float sum = 0.f;
                            for (int j = 0; j < k; ++j) sum += prefix_full_score<MODEL>(kx, a, e, j, xe, ge, prev);
                            int jsel;
                            if (!(sum > 0.0f)) {
                                jsel = (int)(u * (float)k);
                            } else {
                                const float tgt = u * sum;
                                float cum = 0.f; jsel = -1; int lastpos = 0;
                                for (int j = 0; j < k; ++j) {
                                    const float v = prefix_full_score<MODEL>(kx, a, e, j, xe, ge, prev);
                                    cum += v;
                                    if (v > 0.0f) lastpos = j;
                                    if (jsel < 0 && v > 0.0f && cum > tgt) jsel = j;
                                }
                                if (jsel < 0) jsel = lastpos;
                            }
                            if (jsel >= k) jsel = k - 1;
                            const int vote = eff[jsel];
                            red_inc(PFX_TALLY(kx, a) + 4 * vote);
                            if (MODEL) sts_u16(PFX_VOTES(kx, a) + 2 * slot, vote);
                        });
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
            write_trial_outputs(a.out, t, set, C, R, lane, tally, winner, rounds, reason);
            acc_rounds += (unsigned long long)rounds; acc_trials += 1;
            __syncwarp();
        }
    }
    if (cur_set >= 0 && lane == 0 && a.out.totals && acc_trials) {
        atomicAdd(&a.out.totals[2 * cur_set], acc_rounds); atomicAdd(&a.out.totals[2 * cur_set + 1], acc_trials);
    }
}
