/*
 * conclave_oracle.c — plain-C CPU restatement of the conclave voting engine.
 *
 * TEST INFRASTRUCTURE ONLY: the checker for tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Nothing under conclave_sim_b200/
 * links or loads it.  Parity status: PINNED — tests/test_oracle_c.py checks it against
 * the golden vectors the unmodified reference produced (oracle/gen_golden.py) and against
 * the NumPy restatement (oracle/conclave_oracle.py).
 *
 * Follows (file:line into /root/reference):
 *   co_beta_schedule   src/model.py:562-580
 *   co_prob_matrix     src/model.py:303-486
 *   co_pairwise_sum    numpy's pairwise summation (third-party, numpy>=1.23 per
 *                      requirements.txt:2), reached through np.sum at model.py:448,451
 *                      and simulate.py:148
 *   draw (upper bound of cdf)  numpy RandomState.choice as called at src/simulate.py:153
 *   co_run             src/simulate.py:66-206 (+ :84-118 fatigue set, :174-186 threshold/run-off)
 *                      and the histograms of src/simulate.py:356-387
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC -I../include conclave_oracle.c -o _build/libconclave_oracle.so -lm
 * (-O2 without -ffast-math: fp64 evaluation order is part of the contract.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "conclave_b200.h"

int co_version(void) { return 2; }

/* co_set_faithful(1): co_run re-evaluates the whole E x C matrix (exp included) for every trial and every
 * round, exactly as the reference does (simulate.py:121 calls the model each round; model.py:311-486),
 * instead of using trial-invariant per-round tables.  Results are identical; only the cost differs.
 * This is the mode the cpu_baseline / --impl reference legs of bench.py time. */
static int g_faithful = 0;
void co_set_faithful(int on) { g_faithful = on; }

/* ---------------- beta schedule: model.py:562-580 ---------------- */
static double beta_step(const csim_params* p, double beta, int r) {
    if (!p->enable_dynamic_beta || r <= 0) return beta;
    if (p->beta_mode == CSIM_BETA_STEPPED) {
        int iv = p->beta_increment_interval_rounds;
        if (iv > 0 && r % iv == 0) beta = beta + p->beta_increment_amount;
    } else if (p->beta_growth_rate > 1.0) {
        beta = beta * p->beta_growth_rate;
    }
    return beta;
}

void co_beta_schedule(const csim_params* p, int first_round, int n, double* out) {
    double b = p->initial_beta_weight;
    for (int r = 1; r < first_round + n; ++r) {
        b = beta_step(p, b, r);
        if (r >= first_round) out[r - first_round] = b;
    }
}

int co_need_votes(double thr, int E) {
    double t = thr * (double)E;
    int v = (int)floor(t);
    if (v < 0) v = 0;
    while ((double)v < t) ++v;
    while (v > 0 && (double)(v - 1) >= t) --v;
    return v;
}

/* ---------------- numpy pairwise sum ---------------- */
double co_pairwise_sum(const double* a, int n, int stride) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[(size_t)i * stride];
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[(size_t)j * stride];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[(size_t)(i + j) * stride];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[(size_t)i * stride];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return co_pairwise_sum(a, n2, stride) + co_pairwise_sum(a + (size_t)n2 * stride, n - n2, stride);
}

/* ---------------- base scores W[e][c] = exp(-beta d) * pap + regional: model.py:312-323 ---------------- */
static void base_scores(int E, const double* x, const int32_t* region, const uint8_t* pap,
                        const csim_params* p, double beta, double* W) {
    for (int e = 0; e < E; ++e)
        for (int c = 0; c < E; ++c) {
            double s = exp(-beta * fabs(x[e] - x[c]));
            if (pap[c]) s *= p->papabile_weight_factor;
            if (region[e] == region[c]) s += p->regional_affinity_bonus;
            W[(size_t)e * E + c] = s;
        }
}

/* model.py:326-486 on top of W.  Any of prev_vote / prev_count / fatigued / shares may be NULL. */
static void finish_matrix(int E, const double* x, const csim_params* p, const double* W,
                          const int32_t* prev_vote, const int32_t* prev_count,
                          const uint8_t* fatigued, const double* shares, double* P, double* bw /*[E] scratch*/) {
    int have_bw = 0;
    if (p->bandwagon_strength > 0 && prev_count) {
        int mx = 0;
        for (int c = 0; c < E; ++c) if (prev_count[c] > mx) mx = prev_count[c];
        if (mx > 0) {
            have_bw = 1;
            for (int c = 0; c < E; ++c)
                bw[c] = prev_count[c] > 0 ? ((double)prev_count[c] / (double)mx) * p->bandwagon_strength : 0.0;
        }
    }
    int do_fat = p->enable_candidate_fatigue && fatigued;
    int do_stop = p->enable_stop_candidate && shares;
    double D = p->stop_candidate_threshold_unacceptable_distance, T = p->stop_candidate_threat_min_vote_share;
    for (int e = 0; e < E; ++e) {
        double* row = P + (size_t)e * E;
        int trig = 0;
        if (do_stop)
            for (int k = 0; k < E && !trig; ++k)
                if (fabs(x[e] - x[k]) > D && shares[k] > T) trig = 1;
        for (int c = 0; c < E; ++c) {
            double s = W[(size_t)e * E + c];
            if (have_bw) s = s + bw[c];
            if (prev_vote && p->stickiness_factor > 0 && prev_vote[e] == c) s *= (1 + p->stickiness_factor);
            if (!(s > 0)) s = (s != s) ? s : 0.0;                 /* np.maximum(s, 0) keeps NaN */
            if (do_fat && fatigued[c]) s *= p->fatigue_penalty_factor;
            if (trig && !(fabs(x[e] - x[c]) > D)) s *= p->stop_candidate_boost_factor;
            if (c == e) s = 0.0;
            if (!(s > 0)) s = (s != s) ? s : 0.0;
            row[c] = s;
        }
        double rs = co_pairwise_sum(row, E, 1);
        if (rs == 0) {
            for (int c = 0; c < E; ++c) row[c] = 1e-9;
            rs = co_pairwise_sum(row, E, 1);
        }
        for (int c = 0; c < E; ++c) row[c] = row[c] / rs;
    }
}

void co_prob_matrix(int E, const double* x, const int32_t* region, const uint8_t* pap, const csim_params* p,
                    double beta, const int32_t* prev_vote, const int32_t* prev_count,
                    const uint8_t* fatigued, const double* shares, double* P) {
    if (E <= 0) return;
    double* W = (double*)malloc(sizeof(double) * (size_t)E * E);
    double* bw = (double*)malloc(sizeof(double) * E);
    base_scores(E, x, region, pap, p, beta, W);
    finish_matrix(E, x, p, W, prev_vote, prev_count, fatigued, shares, P, bw);
    free(W); free(bw);
}

/* ---------------- Philox4x32-10 (Salmon et al. SC'11) ---------------- */
void co_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

/* stream layout: see include/conclave_b200.h (csim_run_args.seed) and csrc/common.cuh */
void co_philox_uniforms(uint64_t seed, uint64_t sim, int round_num, int E, double* u) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int q = 0; q < (E + 3) / 4; ++q) {
        uint32_t ca[4] = {(uint32_t)sim, (uint32_t)(sim >> 32), (uint32_t)round_num, (uint32_t)q}, wa[4];
        uint32_t cb[4] = {(uint32_t)sim, (uint32_t)(sim >> 32), (uint32_t)round_num, (uint32_t)q | 0x80000000u}, wb[4];
        co_philox4x32_10(ca, key, wa);
        co_philox4x32_10(cb, key, wb);
        for (int i = 0; i < 4 && 4 * q + i < E; ++i) u[4 * q + i] = u53(wa[i], wb[i]);
    }
}

/* cdf of one probability row prefix, exactly as simulate.py:144-153 + numpy choice do it */
static void row_cdf(const double* prow, int k, int full, double* cdf) {
    double* row = cdf; /* build in place */
    if (full) {
        for (int i = 0; i < k; ++i) row[i] = prow[i];
        double sm = co_pairwise_sum(row, k, 1);
        if (!(fabs(sm - 1.0) <= 1e-8 + 1e-5)) {   /* np.isclose(sum, 1.0) false (never for a full row) */
            if (sm > 0) for (int i = 0; i < k; ++i) row[i] = row[i] / sm;
            else for (int i = 0; i < k; ++i) row[i] = 1.0 / k;
        }
    } else {
        double sm = co_pairwise_sum(prow, k, 1);
        if (!(fabs(sm - 1.0) <= 1e-8 + 1e-5)) {
            if (sm > 0) for (int i = 0; i < k; ++i) row[i] = prow[i] / sm;
            else for (int i = 0; i < k; ++i) row[i] = 1.0 / k;
        } else {
            for (int i = 0; i < k; ++i) row[i] = prow[i];
        }
    }
    double acc = 0.0;
    for (int i = 0; i < k; ++i) { acc = (i == 0) ? row[0] : acc + row[i]; cdf[i] = acc; }
    double last = cdf[k - 1];
    for (int i = 0; i < k; ++i) cdf[i] = cdf[i] / last;
}

static inline int upper_bound(const double* cdf, int n, double u) {
    int lo = 0, hi = n;                 /* #{i : cdf[i] <= u} == searchsorted(side='right') */
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
    return lo;
}

/* first k of a stable descending sort of tally (ties -> lowest index) == Series.nlargest(k) */
static void top_k(const int32_t* tally, int C, int k, int32_t* eff, uint8_t* used) {
    memset(used, 0, C);
    for (int j = 0; j < k && j < C; ++j) {
        int best = -1;
        for (int c = 0; c < C; ++c)
            if (!used[c] && (best < 0 || tally[c] > tally[best])) best = c;
        eff[j] = best; used[best] = 1;
    }
}

/* simulate.py:84-118; shares history = tallies history / E.  Boundary ties -> highest index immune. */
static void fatigue_mask(const csim_params* p, int r, int n_hist, const int32_t* hist /*[F][C] ring, oldest..newest handled by caller*/,
                         const int32_t* const* last /*pointers to last min(n_hist,F) tallies, oldest first*/, int nh,
                         int C, int E, uint8_t* mask, uint8_t* immune, double* avg) {
    (void)hist;
    memset(mask, 0, C);
    int F = p->fatigue_threshold_rounds;
    if (!p->enable_candidate_fatigue || r <= F) return;
    memset(immune, 0, C);
    int n = p->fatigue_top_n_immune;
    if (n > 0 && n_hist > 0 && nh > 0) {
        for (int c = 0; c < C; ++c) {
            double acc = (double)last[0][c] / (double)E;
            for (int k = 1; k < nh; ++k) acc += (double)last[k][c] / (double)E;
            avg[c] = acc / (double)nh;
        }
        for (int j = 0; j < n && j < C; ++j) {
            int best = -1;
            for (int c = 0; c < C; ++c)
                if (!immune[c] && (best < 0 || avg[c] >= avg[best])) best = c;   /* >= : highest index among equals */
            immune[best] = 1;
        }
    }
    if (n_hist >= F) {
        for (int c = 0; c < C; ++c) {
            if (immune[c]) continue;
            int low = 1;
            for (int k = nh - F; k < nh && low; ++k)
                if (k >= 0 && !((double)last[k][c] / (double)E < p->fatigue_vote_share_threshold)) low = 0;
            mask[c] = (uint8_t)low;
        }
    }
}

/*
 * Batch of trials.  Layout and semantics of every argument as csim_run_args in
 * include/conclave_b200.h (host buffers).  Returns 0, or -1 on bad arguments / allocation failure.
 */
int co_run(int E, const double* x, const int32_t* region, const uint8_t* pap,
           const csim_params* params, int n_sets, int wiring,
           uint64_t sim_id_begin, uint64_t n_sims, uint64_t seed, const double* tape,
           int32_t* winner, int32_t* rounds_out, uint8_t* reason, int32_t* final_votes, int32_t* round_tallies,
           int64_t* winner_hist, int64_t* rounds_hist, int64_t* totals, int n_threads) {
    if (E < 0 || n_sets <= 0 || !params) return -1;
    const int C = E;
    const int R = params[0].max_rounds;
    for (int s = 0; s < n_sets; ++s) if (params[s].max_rounds != R) return -1;
    if (E == 0) {
        for (uint64_t t = 0; t < (uint64_t)n_sets * n_sims; ++t) {
            winner[t] = -1; rounds_out[t] = R >= 1 ? 1 : 0;
            if (reason) reason[t] = R >= 1 ? CSIM_REASON_NO_PROBS : CSIM_REASON_MAX_ROUNDS;
        }
        return 0;
    }
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    int rc = 0;
    for (int set = 0; set < n_sets; ++set) {
        const csim_params* p = &params[set];
        const int k = p->runoff_candidate_count < C ? p->runoff_candidate_count : C;
        const int need = co_need_votes(p->supermajority_threshold, E);
        double* betas = (double*)malloc(sizeof(double) * (R > 0 ? R : 1));
        co_beta_schedule(p, 1, R, betas);
        /* trial-invariant per-round tables */
        size_t EC = (size_t)E * C;
        double* W = (double*)malloc(sizeof(double) * EC * (size_t)(R > 0 ? R : 1));   /* base scores (model wiring) or P (ref_cli) */
        double* cdf_full = NULL; double* cdf_run = NULL;
        if (!W) { free(betas); return -1; }
        const int tables = (wiring == CSIM_WIRING_REF_CLI) && !g_faithful;
        if (tables) {
            cdf_full = (double*)malloc(sizeof(double) * EC * (size_t)R);
            cdf_run = (double*)malloc(sizeof(double) * (size_t)E * (k > 0 ? k : 1) * (size_t)R);
            if (!cdf_full || !cdf_run) { free(betas); free(W); free(cdf_full); free(cdf_run); return -1; }
#pragma omp parallel for schedule(dynamic, 1)
            for (int r = 0; r < R; ++r) {
                double* Wr = W + EC * r;
                double* bw = (double*)malloc(sizeof(double) * E);
                double* P = (double*)malloc(sizeof(double) * EC);
                base_scores(E, x, region, pap, p, betas[r], Wr);
                finish_matrix(E, x, p, Wr, NULL, NULL, NULL, NULL, P, bw);
                for (int e = 0; e < E; ++e) {
                    row_cdf(P + (size_t)e * C, C, 1, cdf_full + EC * r + (size_t)e * C);
                    if (k > 0) row_cdf(P + (size_t)e * C, k, 0, cdf_run + ((size_t)r * E + e) * k);
                }
                free(bw); free(P);
            }
        } else {
#pragma omp parallel for schedule(dynamic, 1)
            for (int r = 0; r < R; ++r) base_scores(E, x, region, pap, p, betas[r], W + EC * r);
        }
        const int F = p->fatigue_threshold_rounds > 0 ? p->fatigue_threshold_rounds : 0;

#pragma omp parallel
        {
            int32_t* tally = (int32_t*)malloc(sizeof(int32_t) * C);
            int32_t* votes = (int32_t*)malloc(sizeof(int32_t) * E);
            int32_t* eff = (int32_t*)malloc(sizeof(int32_t) * (k > 0 ? k : 1));
            uint8_t* used = (uint8_t*)malloc(C);
            double* u = (double*)malloc(sizeof(double) * (E + 1));
            int32_t* cur = (int32_t*)malloc(sizeof(int32_t) * E);
            double* P = NULL; double* bw = NULL; double* cdf = NULL; int32_t* hist = NULL;
            uint8_t* fat = NULL; uint8_t* immune = NULL; double* avg = NULL; double* shares = NULL;
            const int32_t** last = NULL;
            double* Wt = NULL;
            if (g_faithful) Wt = (double*)malloc(sizeof(double) * EC);
            if (!tables) {
                P = (double*)malloc(sizeof(double) * EC); bw = (double*)malloc(sizeof(double) * E);
                cdf = (double*)malloc(sizeof(double) * C);
                hist = (int32_t*)malloc(sizeof(int32_t) * (size_t)(F > 0 ? F : 1) * C);
                fat = (uint8_t*)malloc(C); immune = (uint8_t*)malloc(C); avg = (double*)malloc(sizeof(double) * C);
                shares = (double*)malloc(sizeof(double) * C);
                last = (const int32_t**)malloc(sizeof(int32_t*) * (F > 0 ? F : 1));
            }
            int64_t* wh = winner_hist ? (int64_t*)calloc(C + 1, sizeof(int64_t)) : NULL;
            int64_t* rh = rounds_hist ? (int64_t*)calloc(R + 1, sizeof(int64_t)) : NULL;
            int64_t tot_rounds = 0, tot_trials = 0;

#pragma omp for schedule(dynamic, 64)
            for (int64_t i = 0; i < (int64_t)n_sims; ++i) {
                const uint64_t t = (uint64_t)set * n_sims + (uint64_t)i;
                const uint64_t sim = sim_id_begin + (uint64_t)i;
                int have_eff = 0, win = -1, rr = 0, why = CSIM_REASON_MAX_ROUNDS, n_hist = 0, have_prev = 0;
                memset(tally, 0, sizeof(int32_t) * C);
                for (int r = 1; r <= R; ++r) {
                    rr = r;
                    if (tape) memcpy(u, tape + (t * (uint64_t)R + (uint64_t)(r - 1)) * (uint64_t)E, sizeof(double) * E);
                    else co_philox_uniforms(seed, sim, r, E, u);
                    if (!tables) {
                        const int live = wiring == CSIM_WIRING_MODEL;
                        const double* Wr = W + EC * (r - 1);
                        if (g_faithful) { base_scores(E, x, region, pap, p, betas[r - 1], Wt); Wr = Wt; }
                        int nh = n_hist < F ? n_hist : F;
                        for (int q = 0; q < nh; ++q) last[q] = hist + (size_t)((n_hist - nh + q) % (F > 0 ? F : 1)) * C;
                        fatigue_mask(p, r, n_hist, hist, last, nh, C, E, fat, immune, avg);
                        for (int c = 0; c < C; ++c) shares[c] = have_prev ? (double)tally[c] / (double)E : 0.0;
                        finish_matrix(E, x, p, Wr, (live && have_prev) ? votes : NULL, (live && have_prev) ? tally : NULL,
                                      live ? fat : NULL, live ? shares : NULL, P, bw);
                    }
                    /* all E draws of the round are made against round r-1's state; prev votes stay intact until then */
                    for (int e = 0; e < E; ++e) {
                        int j;
                        if (tables) {
                            if (!have_eff) j = upper_bound(cdf_full + EC * (r - 1) + (size_t)e * C, C, u[e]);
                            else j = eff[upper_bound(cdf_run + ((size_t)(r - 1) * E + e) * k, k, u[e])];
                        } else {
                            if (!have_eff) { row_cdf(P + (size_t)e * C, C, 1, cdf); j = upper_bound(cdf, C, u[e]); }
                            else { row_cdf(P + (size_t)e * C, k, 0, cdf); j = eff[upper_bound(cdf, k, u[e])]; }
                        }
                        cur[e] = j;
                    }
                    memset(tally, 0, sizeof(int32_t) * C);
                    for (int e = 0; e < E; ++e) { votes[e] = cur[e]; tally[cur[e]]++; }
                    have_prev = 1;
                    if (wiring == CSIM_WIRING_MODEL && F > 0) { memcpy(hist + (size_t)(n_hist % F) * C, tally, sizeof(int32_t) * C); }
                    n_hist++;
                    if (round_tallies) memcpy(round_tallies + (t * (uint64_t)R + (uint64_t)(r - 1)) * (uint64_t)C, tally, sizeof(int32_t) * C);
                    int mx = 0, arg = 0;
                    for (int c = 0; c < C; ++c) if (tally[c] > mx) { mx = tally[c]; arg = c; }
                    if (mx >= need) { win = arg; why = CSIM_REASON_WINNER; break; }
                    if (r >= p->runoff_threshold_rounds && k > 0) { top_k(tally, C, k, eff, used); have_eff = 1; }
                }
                winner[t] = win; rounds_out[t] = rr;
                if (reason) reason[t] = (uint8_t)why;
                if (final_votes) memcpy(final_votes + t * (uint64_t)C, tally, sizeof(int32_t) * C);
                if (round_tallies)
                    for (int r = rr; r < R; ++r)
                        for (int c = 0; c < C; ++c) round_tallies[(t * (uint64_t)R + (uint64_t)r) * (uint64_t)C + c] = -1;
                if (wh) wh[win >= 0 ? win : C]++;
                if (rh && win >= 0) rh[rr]++;
                tot_rounds += rr; tot_trials++;
            }
#pragma omp critical
            {
                if (wh) for (int c = 0; c <= C; ++c) winner_hist[(size_t)set * (C + 1) + c] += wh[c];
                if (rh) for (int r = 0; r <= R; ++r) rounds_hist[(size_t)set * (R + 1) + r] += rh[r];
                if (totals) { totals[2 * set] += tot_rounds; totals[2 * set + 1] += tot_trials; }
            }
            free(cur); free(tally); free(votes); free(eff); free(used); free(u); free(P); free(bw); free(cdf); free(hist);
            free(fat); free(immune); free(avg); free(shares); free((void*)last); free(wh); free(rh); free(Wt);
        }
        free(betas); free(W); free(cdf_full); free(cdf_run);
    }
    return rc;
}
