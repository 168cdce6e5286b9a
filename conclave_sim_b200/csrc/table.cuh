// table.cuh — the TABLE engine (CSIM_ENGINE_TABLE; what CSIM_ENGINE_AUTO picks for the ref_cli wiring).
//
// Under the ref_cli wiring the probability matrix depends on (roster, parameters, round) only — it is
// the same for every trial (the previous round's votes never reach the model).  So the
// per-round CDF rows are built ONCE per parameter set, with exactly the fp64 arithmetic of the exact
// engine (model.py:303-486, simulate.py:144-153), and a trial only has to draw:
//
//   full-field round r:   vote_e = #{c : cdf_r[e][c] <= u}      binary search over a row of C entries
//   run-off round r:      vote_e = eff[ #{j < k : cdfk_r[e][j] <= u} ]   (k = 2: one compare with cdfk_r[e][0])
//
// Work is organised round-major per CTA: a CTA owns a super-batch of (warps x 32) trials of ONE
// parameter set; for r = 1..(last full-field round) it stages cdf_r (E x C, 72 KB in fp32 at E = 134)
// into shared memory once and every warp then plays round r of its 32 trials, one trial at a time with
// the 32 lanes over electors.  Lane i keeps trial i's state (alive, winner, rounds, run-off pair) in
// registers.  Rounds of a trial do not depend on each other in this wiring except through "still
// alive", which is what makes the round-major order legal.  The run-off rounds are then played per
// trial from a small staged table.  When E x C does not fit in shared memory (cfg-5: 1024 x 1024) the
// rows are searched in global memory (L2-resident) instead — same code, different base pointer.
//
// Real = double: tables are the exact fp64 CDFs and uniforms have 53 bits (tape or Philox): results are
// bit-identical to the exact engine and to the oracle.  Real = float: the same CDFs rounded once to
// fp32, 24-bit uniforms: the fast path.
#pragma once
#include "common.cuh"
#include "exact64.cuh"

// ---------------------------------------------------------------------------------------------
// K0: CDF tables.  One thread per (set, round, elector) row.
//   cdf  [sets][R][ECp]    full-field rows [E][C] of one round, ECp = E*C rounded up to 4 elements so that every
//                          round starts 16-byte aligned  (RandomState.choice: cumsum(p) / cumsum(p)[-1])
//   cdfk [sets][R][E][kmax] run-off rows over the first k columns, renormalised as simulate.py:148-151
// ---------------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void cdf_row_exact64(F& prob, int ncol, double* out) {
    const double sm = np_pairwise(prob, ncol);                        // simulate.py:148 np.sum(prob_row)
    const bool close = fabs(sm - 1.0) <= (1e-8 + 1e-5 * 1.0);         // np.isclose(sm, 1.0)
    const int mode = close ? 0 : (sm > 0.0 ? 1 : 2);                  // simulate.py:151
    const double unif = 1.0 / (double)ncol;
    double acc = 0.0;
    for (int i = 0; i < ncol; ++i) {                                  // cdf = p.cumsum()
        const double v = mode == 0 ? prob(i) : (mode == 1 ? __ddiv_rn(prob(i), sm) : unif);
        acc = (i == 0) ? v : __dadd_rn(acc, v);
        out[i] = acc;
    }
    const double last = acc;
    for (int i = 0; i < ncol; ++i) out[i] = __ddiv_rn(out[i], last);  // cdf /= cdf[-1]
}

__host__ __device__ inline size_t cdf_round_stride(int E);

// How the trial kernel deals electors to lanes, and hence the order of the table rows.  Electors below e_quad_end go to lanes as
// whole Philox quads (lane l owns electors 4 (l + 32 qi) .. + 3: one Philox call yields its four uniforms, no shuffles); the tail
// of at most 64 goes one elector per lane.  So that the 32 lanes of a pass (fixed word index i) probe 32 CONSECUTIVE rows — same
// column, odd pitch: conflict-free — the row of elector e is stored at table_slot(e): 128 qi + 32 i + lane.
__host__ __device__ inline int table_e_quad_end(int E) { return (E & ~127) + ((E & 127) > 64 ? 128 : 0); }
__host__ __device__ inline int table_slot(int e, int e_quad_end) {
    return e < e_quad_end ? ((e >> 7) << 7) + ((e & 3) << 5) + ((e >> 2) & 31) : e;
}
__host__ __device__ inline int table_rows(int E) { const int q = table_e_quad_end(E); return E > q ? E : q; }
__host__ __device__ inline size_t cdf_round_stride(int E) { return ((size_t)table_rows(E) * E + 3) & ~(size_t)3; }

static __global__ void k_cdf_tables64(RosterDev ro, const DevSet* sets, const double* W, int R, int kmax,
                               double* cdf, double* cdfk, int sr0) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int sr = sr0 + blockIdx.y;             // set * R + (round - 1)
    const int E = ro.E, C = E;
    if (e >= E) return;
    const DevSet S = sets[sr / R];
    Row64 row;
    row.Wcol = W + (size_t)sr * C * ro.Epad + e; row.Epad = ro.Epad; row.e = e;
    row.prev = -1; row.sticky = false; row.st1p = 1.0; row.bw = nullptr; row.fat = nullptr; row.fat_pen = 1.0;
    row.x = ro.x64; row.xe = ro.x64[e]; row.stop_D = 0.0; row.stop_boost = 1.0; row.trig = false; row.degenerate = false;
    double rs = np_pairwise(row, C);                                  // model.py:451
    if (rs == 0.0) { row.degenerate = true; rs = np_pairwise(row, C); }
    Prob64 p{&row, rs};
    const int eqe = table_e_quad_end(E);
    cdf_row_exact64(p, C, cdf + (size_t)sr * cdf_round_stride(E) + (size_t)table_slot(e, eqe) * C);     // rows in lane-visiting order
    if (e == 0) {                                                    // rows no elector owns (a partly filled quad group) and the padding
        for (int sl = 0; eqe > E && sl < table_rows(E); ++sl) {      // (eqe <= E: the slots are a permutation of the electors)
            bool used = false;
            for (int q = 0; q < E && !used; ++q) used = table_slot(q, eqe) == sl;
            if (!used) for (int i = 0; i < C; ++i) cdf[(size_t)sr * cdf_round_stride(E) + (size_t)sl * C + i] = 1.0;
        }
        for (size_t i = (size_t)table_rows(E) * C; i < cdf_round_stride(E); ++i) cdf[(size_t)sr * cdf_round_stride(E) + i] = 1.0;
    }
    if (S.k > 0) cdf_row_exact64(p, S.k, cdfk + ((size_t)sr * E + e) * kmax);
}

static __global__ void k_f64_to_f32(const double* in, float* out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (float)in[i];
}

// ---------------------------------------------------------------------------------------------
// K1: trials.
// ---------------------------------------------------------------------------------------------
template <class Real>
struct TableArgs {
    RosterDev ro;
    BatchDev b;
    TrialOut out;
    const Real* cdf;        // [sets][R][ECp]  (ECp = cdf_round_stride(E))
    const Real* cdfk;       // [sets][R][E][kmax]
    int kmax;
    int warps_per_cta;
    int tpw;                // trials per warp per super-batch (<= 32: lane i holds trial i)
    int tpw_tail;           // ... of the small super-batches that the END of the job is cut into, so that the CTAs — which claim
                            // super-batches dynamically (wq.next[0] counts them) — run out of work within one small batch of each other
    uint64_t n_big_last;    // big super-batches of the LAST parameter set; the rest of its trials go in small ones
    WorkQueue wq;
    int stage;              // 1: E x C rows fit in shared memory and are staged per round
    int stage_q;            // 1: the [R][E] run-off first-column table is staged too
    int tail_words;         // fp32 only, 0 = off: Philox words kept per trial for the tail electors (4 per tail quad, <= 16).  The
                            // round-major order lets lane i compute the tail quads of trial i for the round ONCE for the whole
                            // warp (tail_words / 4 Philox calls per round instead of one per trial): see tailbuf in the kernel
    size_t cta_smem;        // staged table + run-off table
    size_t warp_smem;
};

// #{i : row[i] <= u}  == searchsorted(row, u, 'right'), for a row whose last entry is > u (a CDF row ends in 1.0 and u < 1).
// Bound-free form: with P = step0 = the largest power of two <= n, ONE first probe decides between the two overlapping windows
// [0, P) and [n - P, n) — the answer is >= n - P exactly when row[n - P - 1] <= u — and log2(P) halving steps then stay inside
// the chosen window by construction, so no probe needs a bound check or an index clamp.  Per probe: one load at a
// compile-time offset from a running pointer, one compare, one predicated add (the clamped form cost 8 instructions, 6 of them
// on the half-rate ALU pipe).  Same probe count as before (1 + log2 P).
template <class Real>
__device__ __forceinline__ int upper_bound_row(const Real* row, int n, int step0, Real u) {
    if (step0 == 128) {                                // 128 <= n <= 255: eight probes, fully unrolled
        const Real* p = row;
        if (n > 128) p += (row[n - 129] <= u) ? (n - 128) : 0;
#pragma unroll
        for (int step = 64; step > 0; step >>= 1) p += (p[step - 1] <= u) ? step : 0;
        return (int)(p - row);
    }
    int pos = (n > step0 && row[n - step0 - 1] <= u) ? n - step0 : 0;
    for (int step = step0 >> 1; step > 0; step >>= 1)
        if (row[pos + step - 1] <= u) pos += step;
    return pos;
}

// The same search on a row staged in shared memory, on a 32-bit shared-window byte address: every probe is ONE ld.shared at an
// immediate offset from the running address, one compare and one select-add (the generic-pointer form above makes the
// compiler re-derive an index and a LEA per probe).  128 <= n <= 255 only (the caller checks step0 == 128).
template <int OFF> __device__ __forceinline__ float lds_at(uint32_t a, float) {
    float v; asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF)); return v;
}
template <int OFF> __device__ __forceinline__ double lds_at(uint32_t a, double) {
    double v; asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF)); return v;
}
template <class Real>
__device__ __forceinline__ int upper_bound_row_smem128(uint32_t row_a, int n, Real u) {
    constexpr int W = (int)sizeof(Real);
    uint32_t a = row_a;
    if (n > 128) {                                     // window [n - 128, n) iff the (n - 128)-th entry is <= u
        const uint32_t d = (uint32_t)(n - 128) * W;
        const Real v = lds_at<-W>(a + d, Real(0));
        a += (v <= u) ? d : 0u;
    }
#define CSIM_TSTEP(S) { const Real v = lds_at<((S) - 1) * W>(a, Real(0)); a += (v <= u) ? (uint32_t)((S) * W) : 0u; }
    CSIM_TSTEP(64) CSIM_TSTEP(32) CSIM_TSTEP(16) CSIM_TSTEP(8) CSIM_TSTEP(4) CSIM_TSTEP(2) CSIM_TSTEP(1)
#undef CSIM_TSTEP
    return (int)((a - row_a) / W);
}

#ifndef CSIM_TABLE_WPC
#define CSIM_TABLE_WPC 16
#endif
template <class Real, bool STAGE>
__global__ void __launch_bounds__(CSIM_TABLE_WPC * 32, 1024 / (CSIM_TABLE_WPC * 32)) k_trials_table(TableArgs<Real> a) {
    constexpr bool F64 = sizeof(Real) == 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int E = a.ro.E, C = E, R = a.b.R, kmax = a.kmax;
    const int Ei = (E + 3) & ~3;
    Real* tab_s = (Real*)smem_raw;                               // staged cdf_r  [E][C]   (stage == 1)
    Real* q_s = tab_s + (STAGE ? cdf_round_stride(E) : 0);       // run-off first-column table [R][table_rows(E)], slot order
    unsigned char* base = smem_raw + a.cta_smem + (size_t)wib * a.warp_smem;
    int32_t* tally = (int32_t*)base;            base += sizeof(int32_t) * Ei;
    int32_t* eff = (int32_t*)base;              base += sizeof(int32_t) * ((kmax + 3) & ~3);
    uint8_t* mark = (uint8_t*)base;             base += (Ei + 15) & ~15;
    uint32_t* tailbuf = (uint32_t*)base;                                    // [32 trials][tail_words] (tail_words > 0 only)

    __shared__ __align__(8) unsigned long long stage_bar;                   // mbarrier of the staged table
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&stage_bar);
    const uint32_t tab_a = (uint32_t)__cvta_generic_to_shared(tab_s);
    uint32_t stage_phase = 0;
    if (STAGE) {
        if (threadIdx.x == 0) mbar_init(bar_a, 1);
        __syncthreads();
    }
    const int wpc = a.warps_per_cta;
    // super-batches: every set but the last in big ones (tpw trials per warp); the last set in n_big_last big ones and then small
    // ones (tpw_tail), so the job ends on short work items.  Claimed in order from one global counter.
    const uint64_t big_trials = (uint64_t)wpc * a.tpw, small_trials = (uint64_t)wpc * a.tpw_tail;
    const uint64_t sb_per_set = (a.b.n_sims + big_trials - 1) / big_trials;
    const uint64_t sb_before_last = sb_per_set * (uint64_t)(a.b.n_sets - 1);
    const uint64_t last_big_sims = a.n_big_last * big_trials;              // <= n_sims
    const uint64_t n_sb = sb_before_last + a.n_big_last + (a.b.n_sims - last_big_sims + small_trials - 1) / small_trials;
    __shared__ unsigned long long s_sb[2];
    int step0 = 1;
    while (step0 * 2 <= C) step0 *= 2;
    const bool top2_ok = C >= 2 && C < 32768;

    // Electors are dealt to lanes as whole Philox quads (table_slot above): lane l draws for electors 4 l .. 4 l + 3 with the
    // four words of ONE Philox call — no shuffles, no word selection — and the rows of a pass are 32 consecutive table rows.
    // fn(e, u, slot): slot = the elector's table row.  The tail (at most 64 electors) goes one elector per lane.
    const int eqe = table_e_quad_end(E);
    const int n_qit = eqe >> 7;
    const int q_rows = table_rows(E);                            // pitch of the staged run-off table (slot order)
    const int tw = F64 ? 0 : a.tail_words;
    // the tail quads of THIS lane's trial for round r -> tailbuf[lane][...]: one call per tail quad for the whole warp
    auto stage_tail_words = [&](int r, uint64_t my_sim) {
        for (int j = 0; j < (tw >> 2); ++j) {
            uint32_t w[4];
            philox4x32_10((uint32_t)my_sim, (uint32_t)(my_sim >> 32), (uint32_t)r, (uint32_t)((eqe >> 2) + j), a.b.seed_lo, a.b.seed_hi, w);
            *reinterpret_cast<uint4*>(tailbuf + lane * tw + 4 * j) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    };
    auto for_each_elector = [&](int r, uint64_t t, uint32_t s_lo, uint32_t s_hi, const uint32_t* tail, auto&& fn) {
        const bool use_tape = F64 && a.b.tape != nullptr;
        const double* tp = use_tape ? a.b.tape + (t * (uint64_t)R + (uint64_t)(r - 1)) * (uint64_t)E : nullptr;
        for (int qi = 0; qi < n_qit; ++qi) {
            const int q = lane + 32 * qi;
            if (4 * q >= E) continue;
            uint32_t w[4] = {0, 0, 0, 0}, wb[4] = {0, 0, 0, 0};
            if (!use_tape) {
                philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)q, a.b.seed_lo, a.b.seed_hi, w);
                if (F64) philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)q | 0x80000000u, a.b.seed_lo, a.b.seed_hi, wb);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = 4 * q + i;
                if (e < E) {
                    Real u;
                    if (use_tape) u = (Real)tp[e];
                    else u = F64 ? (Real)u53_from_words(w[i], wb[i]) : (Real)u24_from_word(w[i]);
                    fn(e, u, (qi << 7) + (i << 5) + lane);
                }
            }
        }
        for (int e = eqe + lane; e < E; e += 32) {
            Real u;
            if (use_tape) u = (Real)tp[e];
            else if (!F64 && tail) u = (Real)u24_from_word(tail[e - eqe]);       // staged for the whole warp by stage_tail_words
            else {
                uint32_t w[4], wb[4] = {0, 0, 0, 0};
                philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)(e >> 2), a.b.seed_lo, a.b.seed_hi, w);
                if (F64) philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)(e >> 2) | 0x80000000u, a.b.seed_lo, a.b.seed_hi, wb);
                const int sel = e & 3;
                const uint32_t wa = sel == 0 ? w[0] : (sel == 1 ? w[1] : (sel == 2 ? w[2] : w[3]));
                const uint32_t wbb = sel == 0 ? wb[0] : (sel == 1 ? wb[1] : (sel == 2 ? wb[2] : wb[3]));
                u = F64 ? (Real)u53_from_words(wa, wbb) : (Real)u24_from_word(wa);
            }
            fn(e, u, e);
        }
    };

    for (int it = 0;; ++it) {
        if (threadIdx.x == 0) s_sb[it & 1] = atomicAdd(a.wq.next, 1ULL);
        __syncthreads();                                         // (slot it & 1 is rewritten two barriers from now at the earliest)
        const uint64_t sb = s_sb[it & 1];
        if (sb >= n_sb) break;
        int set, tpw;
        uint64_t i0;                                             // first sim index of this warp's batch
        if (sb < sb_before_last) {
            set = (int)(sb / sb_per_set); tpw = a.tpw;
            i0 = (sb % sb_per_set) * big_trials + (uint64_t)wib * tpw;
        } else {
            const uint64_t j = sb - sb_before_last;
            set = a.b.n_sets - 1;
            if (j < a.n_big_last) { tpw = a.tpw; i0 = j * big_trials + (uint64_t)wib * tpw; }
            else { tpw = a.tpw_tail; i0 = last_big_sims + (j - a.n_big_last) * small_trials + (uint64_t)wib * tpw; }
        }
        const DevSet S = a.b.sets[set];
        const int k = S.k, need = S.need, rr = S.rr;
        const int r_full = k > 0 ? min(R, max(rr, 1)) : R;       // rounds 1..r_full are full-field (round 1 always is)
        const bool fast2 = (k == 2) && top2_ok;
        // lane i holds the state of trial i0 + i
        const uint64_t my_i = i0 + lane;
        bool alive = lane < tpw && my_i < a.b.n_sims;
        const bool exists = alive;
        int winner = -1, rounds = 0, ef0 = 0, ef1 = 0;
        const uint64_t my_t = (uint64_t)set * a.b.n_sims + my_i;

        // stage the run-off first-column table (k == 2 fast path) once per super-batch
        __syncthreads();
        if (fast2 && a.stage_q)                                  // [R][E] in slot order: a pass reads 32 consecutive entries
            for (int i = threadIdx.x; i < R * E; i += blockDim.x) {
                const int rq = i / E, eq = i - rq * E;
                q_s[rq * q_rows + table_slot(eq, eqe)] = a.cdfk[(((size_t)set * R) * E + i) * kmax];
            }

        // =============== full-field rounds, round-major ===========================================
        for (int r = 1; r <= r_full; ++r) {
            const Real* tab_g = a.cdf + ((size_t)set * R + (r - 1)) * cdf_round_stride(E);
            if (STAGE) {
                __syncthreads();                                  // previous round's readers are done
                // one bulk copy (TMA 1-D, UBLKCP) of the round's whole E x C table, completing on the CTA's mbarrier
                if (threadIdx.x == 0)
                    bulk_load_g2s((uint32_t)__cvta_generic_to_shared(tab_s), tab_g, (uint32_t)(cdf_round_stride(E) * sizeof(Real)), bar_a);
                mbar_wait(bar_a, stage_phase);
                stage_phase ^= 1u;
            }
            __syncthreads();
            if (!__any_sync(CSIM_FULL, alive)) continue;
            if (tw) { __syncwarp(); if (alive) stage_tail_words(r, a.b.sim_id_begin + my_i); __syncwarp(); }
            for (int ti = 0; ti < tpw; ++ti) {
                if (!__shfl_sync(CSIM_FULL, alive ? 1 : 0, ti)) continue;
                const uint64_t sim = a.b.sim_id_begin + i0 + ti;
                const uint64_t t = (uint64_t)set * a.b.n_sims + i0 + ti;
                const uint32_t s_lo = (uint32_t)sim, s_hi = (uint32_t)(sim >> 32);
                for (int c = lane; c < C; c += 32) tally[c] = 0;
                __syncwarp();
                for_each_elector(r, t, s_lo, s_hi, tw ? tailbuf + ti * tw : nullptr, [&](int e, Real u, int slot) {
                    // STAGE: the row lives in shared memory (a typed pointer keeps these LDS, not generic loads)
                    int vote;
                    if (STAGE && step0 == 128) vote = upper_bound_row_smem128<Real>(tab_a + (uint32_t)(slot * C) * (uint32_t)sizeof(Real), C, u);
                    else vote = STAGE ? upper_bound_row<Real>(tab_s + (size_t)slot * C, C, step0, u)
                                      : upper_bound_row<Real>(tab_g + (size_t)slot * C, C, step0, u);
                    vote = vote < C ? vote : C - 1;
                    atomicAdd(&tally[vote], 1);
                });
                __syncwarp();
                if (a.out.round_tallies)
                    for (int c = lane; c < C; c += 32)
                        a.out.round_tallies[(t * (uint64_t)R + (uint64_t)(r - 1)) * (uint64_t)C + c] = tally[c];
                int mx, arg, second = 0;
                if (top2_ok) {
                    uint32_t k1, k2;
                    warp_top2(tally, C, lane, k1, k2);
                    mx = (int)(k1 >> 16); arg = 0xFFFF - (int)(k1 & 0xFFFFu); second = 0xFFFF - (int)(k2 & 0xFFFFu);
                } else {
                    warp_argmax_first(tally, C, lane, mx, arg);
                }
                const bool won = mx >= need;                                        // simulate.py:174-180
                const bool last = won || r == R;
                if (last && a.out.final_votes) for (int c = lane; c < C; c += 32) a.out.final_votes[t * (uint64_t)C + c] = tally[c];
                if (lane == ti) {
                    rounds = r;
                    if (won) { winner = arg; alive = false; }
                    else if (r == R) { alive = false; }
                    else if (r >= rr && k > 0) { ef0 = arg; ef1 = second; }          // simulate.py:183-186 (k == 2)
                }
                __syncwarp();
            }
        }
        // =============== run-off rounds, k == 2: round-major like the full-field rounds ================
        // (lane i keeps the run-off pair of trial i; the tail quads of a round are computed once for the whole warp)
        if (fast2 && r_full < R) {
            for (int r = r_full + 1; r <= R; ++r) {
                if (!__any_sync(CSIM_FULL, alive)) break;
                if (tw) { __syncwarp(); if (alive) stage_tail_words(r, a.b.sim_id_begin + my_i); __syncwarp(); }
                for (int ti = 0; ti < tpw; ++ti) {
                    if (!__shfl_sync(CSIM_FULL, alive ? 1 : 0, ti)) continue;
                    const uint64_t sim = a.b.sim_id_begin + i0 + ti;
                    const uint64_t t = (uint64_t)set * a.b.n_sims + i0 + ti;
                    const int e0 = __shfl_sync(CSIM_FULL, ef0, ti), e1 = __shfl_sync(CSIM_FULL, ef1, ti);
                    int cnt0 = 0;
                    for_each_elector(r, t, (uint32_t)sim, (uint32_t)(sim >> 32), tw ? tailbuf + ti * tw : nullptr, [&](int e, Real u, int slot) {
                        const Real q0 = a.stage_q ? q_s[(r - 1) * q_rows + slot] : a.cdfk[(((size_t)set * R + (r - 1)) * E + e) * kmax];
                        cnt0 += (q0 <= u) ? 0 : 1;                                   // j = #{cdf_j <= u}; vote eff[0] iff cdf_0 > u
                    });
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) cnt0 += __shfl_xor_sync(CSIM_FULL, cnt0, off);
                    const int n0 = cnt0, n1 = E - cnt0;
                    int mx, arg, second;
                    const bool need_tally = a.out.round_tallies || a.out.final_votes || !(n0 > 0 && n1 > 0);
                    if (need_tally) {
                        __syncwarp();
                        for (int c = lane; c < C; c += 32) tally[c] = 0;
                        __syncwarp();
                        if (lane == 0) { tally[e0] = n0; tally[e1] = n1; }
                        __syncwarp();
                    }
                    if (n0 > 0 && n1 > 0) {
                        const bool first0 = n0 > n1 || (n0 == n1 && e0 < e1);
                        mx = first0 ? n0 : n1; arg = first0 ? e0 : e1; second = first0 ? e1 : e0;
                    } else {
                        // one of the pair got nothing: the second place is the lowest-index zero-tally candidate
                        uint32_t k1, k2;
                        warp_top2(tally, C, lane, k1, k2);
                        mx = (int)(k1 >> 16); arg = 0xFFFF - (int)(k1 & 0xFFFFu); second = 0xFFFF - (int)(k2 & 0xFFFFu);
                    }
                    if (a.out.round_tallies)
                        for (int c = lane; c < C; c += 32)
                            a.out.round_tallies[(t * (uint64_t)R + (uint64_t)(r - 1)) * (uint64_t)C + c] = tally[c];
                    const bool won = mx >= need;
                    if ((won || r == R) && a.out.final_votes)
                        for (int c = lane; c < C; c += 32) a.out.final_votes[t * (uint64_t)C + c] = tally[c];
                    if (lane == ti) {
                        rounds = r;
                        if (won) { winner = arg; alive = false; }
                        else if (r == R) { alive = false; }
                        else { ef0 = arg; ef1 = second; }
                    }
                    __syncwarp();
                }
            }
        }
        // =============== run-off rounds, generic k: per trial =======================================
        if (k > 0 && !fast2 && r_full < R) {
            for (int ti = 0; ti < tpw; ++ti) {
                if (!__shfl_sync(CSIM_FULL, alive ? 1 : 0, ti)) continue;
                const uint64_t sim = a.b.sim_id_begin + i0 + ti;
                const uint64_t t = (uint64_t)set * a.b.n_sims + i0 + ti;
                const uint32_t s_lo = (uint32_t)sim, s_hi = (uint32_t)(sim >> 32);
                int e0 = __shfl_sync(CSIM_FULL, ef0, ti), e1 = __shfl_sync(CSIM_FULL, ef1, ti);
                int win = -1, rnd = r_full;
                if (!fast2) {
                    // generic k: rebuild round r_full's stable top-k from its tally: replay that round's draws
                    // (rare path; k != 2).  The tally is recomputed exactly as in the loop above.
                    const Real* tab_g = a.cdf + ((size_t)set * R + (r_full - 1)) * cdf_round_stride(E);
                    for (int c = lane; c < C; c += 32) tally[c] = 0;
                    __syncwarp();
                    for_each_elector(r_full, t, s_lo, s_hi, nullptr, [&](int e, Real u, int slot) {
                        const int vote = upper_bound_row<Real>(tab_g + (size_t)slot * C, C, step0, u);
                        atomicAdd(&tally[vote < C ? vote : C - 1], 1);
                    });
                    __syncwarp();
                    warp_top_k(tally, C, k, lane, eff, mark);
                    __syncwarp();
                }
                for (int r = r_full + 1; r <= R; ++r) {
                    rnd = r;
                    int cnt0 = 0;
                    if (!fast2) { for (int c = lane; c < C; c += 32) tally[c] = 0; __syncwarp(); }
                    for_each_elector(r, t, s_lo, s_hi, nullptr, [&](int e, Real u, int slot) {
                        if (fast2) {
                            const Real q0 = a.stage_q ? q_s[(r - 1) * q_rows + slot] : a.cdfk[(((size_t)set * R + (r - 1)) * E + e) * kmax];
                            cnt0 += (q0 <= u) ? 0 : 1;                               // j = #{cdf_j <= u}; vote eff[0] iff cdf_0 > u
                        } else {
                            const Real* rowk = a.cdfk + (((size_t)set * R + (r - 1)) * E + e) * kmax;
                            int j = 0;
                            for (int i = 0; i < k; ++i) j += (rowk[i] <= u) ? 1 : 0;
                            atomicAdd(&tally[eff[j < k ? j : k - 1]], 1);
                        }
                    });
                    int mx, arg, second;
                    if (fast2) {
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) cnt0 += __shfl_xor_sync(CSIM_FULL, cnt0, off);
                        const int n0 = cnt0, n1 = E - cnt0;
                        if (a.out.round_tallies || a.out.final_votes) {
                            for (int c = lane; c < C; c += 32) tally[c] = 0;
                            __syncwarp();
                            if (lane == 0) { tally[e0] = n0; tally[e1] = n1; }
                            __syncwarp();
                        }
                        if (n0 > 0 && n1 > 0) {
                            const bool first0 = n0 > n1 || (n0 == n1 && e0 < e1);
                            mx = first0 ? n0 : n1; arg = first0 ? e0 : e1; second = first0 ? e1 : e0;
                        } else {
                            // one of the pair got nothing: the second place is the lowest-index zero-tally candidate
                            for (int c = lane; c < C; c += 32) tally[c] = 0;
                            __syncwarp();
                            if (lane == 0) { tally[e0] = n0; tally[e1] = n1; }
                            __syncwarp();
                            uint32_t k1, k2;
                            warp_top2(tally, C, lane, k1, k2);
                            mx = (int)(k1 >> 16); arg = 0xFFFF - (int)(k1 & 0xFFFFu); second = 0xFFFF - (int)(k2 & 0xFFFFu);
                        }
                    } else {
                        __syncwarp();
                        warp_argmax_first(tally, C, lane, mx, arg);
                        second = 0;
                    }
                    if (a.out.round_tallies)
                        for (int c = lane; c < C; c += 32)
                            a.out.round_tallies[(t * (uint64_t)R + (uint64_t)(r - 1)) * (uint64_t)C + c] = tally[c];
                    const bool won = mx >= need;
                    if ((won || r == R) && a.out.final_votes)
                        for (int c = lane; c < C; c += 32) a.out.final_votes[t * (uint64_t)C + c] = tally[c];
                    if (won) { win = arg; break; }
                    if (fast2) { e0 = arg; e1 = second; }
                    else { warp_top_k(tally, C, k, lane, eff, mark); }
                    __syncwarp();
                }
                if (lane == ti) { winner = win; rounds = rnd; alive = false; }
                __syncwarp();
            }
        }
        // =============== outputs ===================================================================
        if (exists && R == 0 && a.out.final_votes) for (int c = 0; c < C; ++c) a.out.final_votes[my_t * (uint64_t)C + c] = 0;
        if (exists) {
            a.out.winner[my_t] = winner; a.out.rounds[my_t] = rounds;
            if (a.out.reason) a.out.reason[my_t] = (uint8_t)(winner >= 0 ? CSIM_REASON_WINNER : CSIM_REASON_MAX_ROUNDS);
            if (a.out.winner_hist) atomicAdd(&a.out.winner_hist[(size_t)set * (C + 1) + (winner >= 0 ? winner : C)], 1ULL);
            if (a.out.rounds_hist && winner >= 0) atomicAdd(&a.out.rounds_hist[(size_t)set * (R + 1) + rounds], 1ULL);
            if (a.out.round_tallies)
                for (size_t i = (size_t)rounds * C; i < (size_t)R * C; ++i) a.out.round_tallies[my_t * (uint64_t)R * C + i] = -1;
        }
        if (a.out.totals) {
            unsigned long long rs = exists ? (unsigned long long)rounds : 0ULL, ts = exists ? 1ULL : 0ULL;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) { rs += __shfl_xor_sync(CSIM_FULL, rs, off); ts += __shfl_xor_sync(CSIM_FULL, ts, off); }
            if (lane == 0 && ts) { atomicAdd(&a.out.totals[2 * set], rs); atomicAdd(&a.out.totals[2 * set + 1], ts); }
        }
    }
}
