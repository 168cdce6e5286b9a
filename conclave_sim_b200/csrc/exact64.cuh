// exact64.cuh — the fp64 "exact" engine: bit-for-bit the arithmetic of the reference.
//
// One warp owns one trial.  Lanes own electors e = lane, lane+32, ...  Every fp64 value
// is produced by the same sequence of IEEE operations the reference performs through
// NumPy: scores (model.py:312-427), NumPy's pairwise row sum (model.py:451),
// division (model.py:481), the worker's np.sum / np.isclose re-normalisation of the row
// prefix (simulate.py:144-151), and RandomState.choice's cumsum / divide / searchsorted
// (simulate.py:153).  Products that feed additions use __dmul_rn/__dadd_rn so nvcc cannot
// contract them into FMAs (NumPy evaluates them as separate ufunc passes).
//
// The trial-invariant part of the score, W_r[e][c] = exp(-beta_r |x_e-x_c|) * pap_c + bonus*[same region],
// is built once per (parameter set, round) by k_base_scores64 and laid out [c][Epad] so that
// lanes (consecutive electors) read consecutive addresses.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------
// K0: base scores.  grid = (ceil(Epad/128), C, <= 65535 of the sets*R tables from sr0), block = 128 (threads over e).
// ---------------------------------------------------------------------------
static __global__ void k_base_scores64(RosterDev ro, const DevSet* sets, const double* betas /*[sets][R]*/, int R, double* W, int sr0) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    const int sr = sr0 + blockIdx.z;           // set * R + (round - 1); sr0: launches are tiled over the 65535 grid limit
    if (e >= ro.Epad) return;
    const int set = sr / R;
    double v = 0.0;
    if (e < ro.E) {
        const double beta = betas[sr];
        const double d = fabs(ro.x64[e] - ro.x64[c]);                 // model.py:312
        v = exp(__dmul_rn(-beta, d));                                 // model.py:313
        if (ro.pap[c]) v = __dmul_rn(v, sets[set].pwf);               // model.py:317-319
        if (ro.region[e] == ro.region[c]) v = __dadd_rn(v, sets[set].bonus);   // model.py:322-323
    }
    W[((size_t)sr * ro.E + c) * ro.Epad + e] = v;
}

// ---------------------------------------------------------------------------
// NumPy's pairwise summation (DOUBLE_pairwise_sum), restated over a functor f(i), i in [0, n).
// ---------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ double np_pairwise_leaf(F& f, int lo, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, f(lo + i));
        return res;
    }
    double r0 = f(lo), r1 = f(lo + 1), r2 = f(lo + 2), r3 = f(lo + 3), r4 = f(lo + 4), r5 = f(lo + 5), r6 = f(lo + 6), r7 = f(lo + 7);
    int i = 8;
    const int lim = n - (n & 7);
    for (; i < lim; i += 8) {
        r0 = __dadd_rn(r0, f(lo + i));     r1 = __dadd_rn(r1, f(lo + i + 1));
        r2 = __dadd_rn(r2, f(lo + i + 2)); r3 = __dadd_rn(r3, f(lo + i + 3));
        r4 = __dadd_rn(r4, f(lo + i + 4)); r5 = __dadd_rn(r5, f(lo + i + 5));
        r6 = __dadd_rn(r6, f(lo + i + 6)); r7 = __dadd_rn(r7, f(lo + i + 7));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (; i < n; ++i) res = __dadd_rn(res, f(lo + i));
    return res;
}

template <class F>
__device__ double np_pairwise(F& f, int n) {
    if (n <= 128) return np_pairwise_leaf(f, 0, n);
    // explicit post-order walk of: sum(lo,n) = sum(lo,n2) + sum(lo+n2,n-n2), n2 = n/2 - (n/2)%8
    int lo_s[16], n_s[16], st_s[16];
    double left_s[16];
    int sp = 0;
    lo_s[0] = 0; n_s[0] = n; st_s[0] = 0; sp = 1;
    double val = 0.0;
    bool have_val = false;
    while (sp > 0) {
        const int top = sp - 1;
        if (!have_val) {
            if (n_s[top] <= 128) { val = np_pairwise_leaf(f, lo_s[top], n_s[top]); have_val = true; --sp; }
            else {
                int n2 = n_s[top] / 2; n2 -= n2 & 7;
                st_s[top] = 1;
                lo_s[sp] = lo_s[top]; n_s[sp] = n2; st_s[sp] = 0; ++sp;
            }
        } else {
            if (st_s[top] == 1) {
                int n2 = n_s[top] / 2; n2 -= n2 & 7;
                left_s[top] = val; st_s[top] = 2; have_val = false;
                lo_s[sp] = lo_s[top] + n2; n_s[sp] = n_s[top] - n2; st_s[sp] = 0; ++sp;
            } else {
                val = __dadd_rn(left_s[top], val); --sp;
            }
        }
    }
    return val;
}

// ---------------------------------------------------------------------------
// One elector's row (model.py:326-449 on top of W_r).
// ---------------------------------------------------------------------------
struct Row64 {
    const double* Wcol;       // &W_r[0][e], stride Epad between candidates
    int Epad, e, prev;        // prev = elector's previous vote, -1 = none
    const double* bw;         // [C] bandwagon bonuses or null
    const uint8_t* fat;       // [C] fatigue mask or null
    const double* x;          // [C] ideology (stop-candidate distances)
    double xe, st1p, fat_pen, stop_D, stop_boost;
    bool sticky, trig, degenerate;
    const int32_t* cand = nullptr;   // optional column subset (model.py:280-296): column a is candidate cand[a]

    __device__ __forceinline__ double operator()(int a) const {
        if (degenerate) return 1e-9;                                  // model.py:447-449
        const int c = cand ? cand[a] : a;
        double s = Wcol[(size_t)c * Epad];
        if (bw) s = __dadd_rn(s, bw[c]);                              // model.py:354
        if (sticky && c == prev) s = __dmul_rn(s, st1p);              // model.py:363
        s = (s > 0.0 || s != s) ? s : 0.0;                            // model.py:369
        if (fat && fat[c]) s = __dmul_rn(s, fat_pen);                 // model.py:379
        if (trig && !(fabs(xe - x[c]) > stop_D)) s = __dmul_rn(s, stop_boost);   // model.py:414
        if (c == e) s = 0.0;                                          // model.py:422
        s = (s > 0.0 || s != s) ? s : 0.0;                            // model.py:427
        return s;
    }
};

// P[e, c] = s / rowsum, and the worker's prefix re-normalisation (simulate.py:144-151).
struct Prob64 {
    const Row64* row; double rs;
    __device__ __forceinline__ double operator()(int c) const { return __ddiv_rn((*row)(c), rs); }
};

// Returns #{i < ncol : cdf_i <= u}, cdf as RandomState.choice builds it.
__device__ __forceinline__ int draw_exact64(Row64& row, int C, int ncol, double u) {
    double rs = np_pairwise(row, C);                                  // model.py:451
    if (rs == 0.0) { row.degenerate = true; rs = np_pairwise(row, C); }
    Prob64 p{&row, rs};
    const double sm = np_pairwise(p, ncol);                           // simulate.py:148 np.sum(prob_row)
    const bool close = fabs(sm - 1.0) <= (1e-8 + 1e-5 * 1.0);         // np.isclose(sm, 1.0)
    const int mode = close ? 0 : (sm > 0.0 ? 1 : 2);                  // simulate.py:151
    const double unif = 1.0 / (double)ncol;
    double last = 0.0;
    for (int i = 0; i < ncol; ++i) {                                  // cdf = p.cumsum()
        const double v = mode == 0 ? p(i) : (mode == 1 ? __ddiv_rn(p(i), sm) : unif);
        last = (i == 0) ? v : __dadd_rn(last, v);
    }
    int cnt = 0;
    double acc = 0.0;
    for (int i = 0; i < ncol; ++i) {                                  // cdf /= cdf[-1]; searchsorted(u, 'right')
        const double v = mode == 0 ? p(i) : (mode == 1 ? __ddiv_rn(p(i), sm) : unif);
        acc = (i == 0) ? v : __dadd_rn(acc, v);
        cnt += (__ddiv_rn(acc, last) <= u) ? 1 : 0;
    }
    return cnt < ncol ? cnt : ncol - 1;
}

struct Exact64Args {
    RosterDev ro;
    BatchDev b;
    TrialOut out;
    const double* W;        // [sets][R][C][Epad]
    int wiring;
    int Fmax;               // ring size of the tally history (>= 1)
    int kmax;               // max k over sets (>= 1)
    int warps_per_cta;
    WorkQueue wq;           // dynamic trial scheduling (common.cuh), one counter over all (set, sim) pairs
    size_t warp_smem;       // bytes of shared memory per warp
};

__host__ __device__ inline size_t exact64_warp_smem(int E, int Fmax, int kmax) {
    size_t n = 0;
    n += sizeof(double) * 2 * (size_t)E;                    // bw, avg
    n += sizeof(int32_t) * (size_t)E * (4 + (size_t)Fmax);  // tally, prev_tally, votes, newvotes, hist[F]
    n += sizeof(int32_t) * (size_t)kmax;                    // eff
    n += 2 * (size_t)E;                                     // fat, mark
    return (n + 15) & ~(size_t)15;
}

#ifndef CSIM_EXACT_MINB
#define CSIM_EXACT_MINB 3      // 80 registers: the fp64 engine is latency-bound, three CTAs per SM nearly halve its time
#endif
static __global__ void __launch_bounds__(256, CSIM_EXACT_MINB) k_trials_exact64(Exact64Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int E = a.ro.E, C = E, R = a.b.R;
    unsigned char* base = smem_raw + (size_t)wib * a.warp_smem;
    double* bw = (double*)base;                 base += sizeof(double) * E;
    double* avg = (double*)base;                base += sizeof(double) * E;
    int32_t* tally = (int32_t*)base;            base += sizeof(int32_t) * E;
    int32_t* prev_tally = (int32_t*)base;       base += sizeof(int32_t) * E;
    int32_t* votes = (int32_t*)base;            base += sizeof(int32_t) * E;
    int32_t* newvotes = (int32_t*)base;         base += sizeof(int32_t) * E;
    int32_t* hist = (int32_t*)base;             base += sizeof(int32_t) * (size_t)E * a.Fmax;
    int32_t* eff = (int32_t*)base;              base += sizeof(int32_t) * a.kmax;
    uint8_t* fat = (uint8_t*)base;              base += E;
    uint8_t* mark = (uint8_t*)base;

    const uint64_t n_trials = (uint64_t)a.b.n_sets * a.b.n_sims;
    const bool model = a.wiring == CSIM_WIRING_MODEL;

    // trials are claimed dynamically from ONE counter over the flat (set, sim) index (common.cuh: WorkQueue; no per-set state is CTA-shared here)
    for (uint64_t t0 = warp_claim_trials(a.wq, 0, lane); t0 < n_trials; t0 = warp_claim_trials(a.wq, 0, lane))
    for (int ti = 0; ti < a.wq.chunk; ++ti) {
        const uint64_t t = t0 + (uint64_t)ti;
        if (t >= n_trials) break;
        const int set = (int)(t / a.b.n_sims);
        const uint64_t sim = a.b.sim_id_begin + (t % a.b.n_sims);
        const DevSet S = a.b.sets[set];
        const int k = S.k;
        bool have_eff = false, have_prev = false;
        int winner = -1, rounds = 0, n_hist = 0, reason = CSIM_REASON_MAX_ROUNDS;
        for (int c = lane; c < C; c += 32) { tally[c] = 0; prev_tally[c] = 0; }
        __syncwarp();

        for (int r = 1; r <= R; ++r) {
            rounds = r;
            const double* Wr = a.W + ((size_t)set * R + (r - 1)) * (size_t)C * a.ro.Epad;
            // ---- per-round model-wiring state from round r-1 -------------------------------
            bool use_bw = false, use_fat = false;
            if (model) {
                // fatigue set, simulate.py:84-118 (ties across the immunity boundary -> highest index)
                const int F = S.fat_rounds;
                for (int c = lane; c < C; c += 32) { fat[c] = 0; mark[c] = 0; }
                __syncwarp();
                if (S.en_fat && r > F) {
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
                        for (int j = 0; j < n_imm; ++j) {
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
                            for (int q = 0; q < F && low; ++q)      // share < threshold as an exact integer test (DevSet::fat_cnt)
                                low = hist[(size_t)((n_hist - F + q) % a.Fmax) * C + c] < S.fat_cnt;
                            fat[c] = low ? 1 : 0;
                        }
                        use_fat = true;
                    }
                    __syncwarp();
                }
                // bandwagon bonuses, model.py:326-354
                if (S.has_bw && have_prev) {
                    int mx, arg;
                    warp_argmax_first(prev_tally, C, lane, mx, arg);
                    if (mx > 0) {
                        use_bw = true;
                        for (int c = lane; c < C; c += 32)
                            bw[c] = prev_tally[c] > 0 ? __dmul_rn(__ddiv_rn((double)prev_tally[c], (double)mx), S.bandwagon) : 0.0;
                    }
                }
                __syncwarp();
            }
            // ---- every elector draws ---------------------------------------------------------
            for (int c = lane; c < C; c += 32) tally[c] = 0;
            __syncwarp();
            for (int e = lane; e < E; e += 32) {
                double u;
                if (a.b.tape) {
                    u = a.b.tape[(t * (uint64_t)R + (uint64_t)(r - 1)) * (uint64_t)E + e];
                } else {
                    uint32_t wa[4], wb[4];
                    philox4x32_10((uint32_t)sim, (uint32_t)(sim >> 32), (uint32_t)r, (uint32_t)(e >> 2), a.b.seed_lo, a.b.seed_hi, wa);
                    philox4x32_10((uint32_t)sim, (uint32_t)(sim >> 32), (uint32_t)r, (uint32_t)(e >> 2) | 0x80000000u, a.b.seed_lo, a.b.seed_hi, wb);
                    u = u53_from_words(wa[e & 3], wb[e & 3]);
                }
                Row64 row;
                row.Wcol = Wr + e; row.Epad = a.ro.Epad; row.e = e;
                row.prev = (model && have_prev) ? votes[e] : -1;
                row.sticky = model && have_prev && S.has_sticky;
                row.st1p = S.st1p;
                row.bw = use_bw ? bw : nullptr;
                row.fat = (model && S.en_fat && use_fat) ? fat : nullptr;
                row.fat_pen = S.fat_pen;
                row.x = a.ro.x64; row.xe = a.ro.x64[e];
                row.stop_D = S.stop_D; row.stop_boost = S.stop_boost;
                row.trig = false; row.degenerate = false;
                if (model && S.en_stop) {                                     // model.py:384-400
                    for (int q = 0; q < C; ++q) {                 // share > threshold as an exact integer test (DevSet::stop_cnt)
                        const bool threat = have_prev ? prev_tally[q] >= S.stop_cnt : 0.0 > S.stop_T;
                        if (threat && fabs(row.xe - a.ro.x64[q]) > S.stop_D) { row.trig = true; break; }
                    }
                }
                const int ncol = have_eff ? k : C;
                const int j = draw_exact64(row, C, ncol, u);
                const int v = have_eff ? eff[j] : j;
                newvotes[e] = v;
                atomicAdd(&tally[v], 1);
            }
            __syncwarp();
            // ---- close the round ---------------------------------------------------------------
            if (a.out.round_tallies)
                for (int c = lane; c < C; c += 32)
                    a.out.round_tallies[(t * (uint64_t)R + (uint64_t)(r - 1)) * (uint64_t)C + c] = tally[c];
            for (int c = lane; c < C; c += 32) {
                prev_tally[c] = tally[c];
                if (model) hist[(size_t)(n_hist % a.Fmax) * C + c] = tally[c];
            }
            for (int e = lane; e < E; e += 32) votes[e] = newvotes[e];
            have_prev = true; ++n_hist;
            __syncwarp();
            int mx, arg;
            warp_argmax_first(tally, C, lane, mx, arg);
            if (mx >= S.need) { winner = arg; reason = CSIM_REASON_WINNER; break; }     // simulate.py:174-180
            if (r >= S.rr && k > 0) {                                                    // simulate.py:183-186
                warp_top_k(tally, C, k, lane, eff, mark);
                have_eff = true;
            }
            __syncwarp();
        }
        // ---- trial outputs ---------------------------------------------------------------------
        if (a.out.final_votes) for (int c = lane; c < C; c += 32) a.out.final_votes[t * (uint64_t)C + c] = tally[c];
        if (a.out.round_tallies)
            for (size_t i = (size_t)rounds * C + lane; i < (size_t)R * C; i += 32) a.out.round_tallies[t * (uint64_t)R * C + i] = -1;
        if (lane == 0) {
            a.out.winner[t] = winner; a.out.rounds[t] = rounds;
            if (a.out.reason) a.out.reason[t] = (uint8_t)reason;
            if (a.out.winner_hist) atomicAdd(&a.out.winner_hist[(size_t)set * (C + 1) + (winner >= 0 ? winner : C)], 1ULL);
            if (a.out.rounds_hist && winner >= 0) atomicAdd(&a.out.rounds_hist[(size_t)set * (R + 1) + rounds], 1ULL);
            if (a.out.totals) { atomicAdd(&a.out.totals[2 * set], (unsigned long long)rounds); atomicAdd(&a.out.totals[2 * set + 1], 1ULL); }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// K2: the standalone probability matrix (TransitionModel.calculate_transition_probabilities).
// One thread per elector row; W is the [c][Epad] base-score table of this call's beta.
// ---------------------------------------------------------------------------
struct ProbArgs {
    RosterDev ro;
    DevSet S;
    const double* W;          // [C][Epad]
    const int32_t* prev_vote; // [E] or null
    const int32_t* prev_count;// [E] or null
    const uint8_t* fatigued;  // [E] or null
    const double* shares;     // [E] or null
    const int32_t* cand;      // [ncand] or null (all candidates)
    int ncand;
    double* bw;               // [E] scratch (device)
    double* P;                // [E][ncand] out
};

static __global__ void k_prob_bw64(ProbArgs a) {       // single block: bandwagon bonuses (model.py:326-354)
    __shared__ int smx;
    const int C = a.ro.E;
    if (threadIdx.x == 0) smx = 0;
    __syncthreads();
    int m = 0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) m = max(m, a.prev_count[c]);
    atomicMax(&smx, m);
    __syncthreads();
    const int mx = smx;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        a.bw[c] = (mx > 0 && a.prev_count[c] > 0) ? __dmul_rn(__ddiv_rn((double)a.prev_count[c], (double)mx), a.S.bandwagon) : 0.0;
}

static __global__ void k_prob_matrix64(ProbArgs a, int use_bw) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int E = a.ro.E, C = E;
    if (e >= E) return;
    const int NC = a.cand ? a.ncand : C;
    Row64 row;
    row.cand = a.cand;
    row.Wcol = a.W + e; row.Epad = a.ro.Epad; row.e = e;
    row.prev = a.prev_vote ? a.prev_vote[e] : -1;
    row.sticky = a.prev_vote != nullptr && a.S.has_sticky;
    row.st1p = a.S.st1p;
    row.bw = use_bw ? a.bw : nullptr;
    row.fat = (a.S.en_fat && a.fatigued) ? a.fatigued : nullptr;
    row.fat_pen = a.S.fat_pen;
    row.x = a.ro.x64; row.xe = a.ro.x64[e];
    row.stop_D = a.S.stop_D; row.stop_boost = a.S.stop_boost;
    row.trig = false; row.degenerate = false;
    if (a.S.en_stop && a.shares)
        for (int q = 0; q < C; ++q)
            if (fabs(row.xe - a.ro.x64[q]) > a.S.stop_D && a.shares[q] > a.S.stop_T) { row.trig = true; break; }
    double rs = np_pairwise(row, NC);
    if (rs == 0.0) { row.degenerate = true; rs = np_pairwise(row, NC); }
    for (int c = 0; c < NC; ++c) a.P[(size_t)e * NC + c] = __ddiv_rn(row(c), rs);
}
