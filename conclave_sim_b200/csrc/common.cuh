// common.cuh — device-side structures and helpers shared by the conclave kernels.
//
// Reference semantics (file:line into the reference repo) are cited at each
// function; the arithmetic itself is restated from SURVEY.md Appendix A and the
// CPU oracle, never copied.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "conclave_b200.h"

#define CSIM_WARP 32
#define CSIM_FULL 0xffffffffu

// Per parameter-set constants, precomputed on the host (fp64 exactly as the
// reference computes them; fp32 copies for the fast path).
struct DevSet {
    double st1p;          // 1 + stickiness_factor                    (model.py:363)
    double bandwagon;     // bandwagon_strength                        (model.py:349)
    double bonus;         // regional_affinity_bonus                   (model.py:323)
    double pwf;           // papabile_weight_factor                    (model.py:319)
    double fat_thr;       // fatigue_vote_share_threshold              (simulate.py:111)
    double fat_pen;       // fatigue_penalty_factor                    (model.py:379)
    double stop_D;        // stop_candidate_threshold_unacceptable_distance (model.py:390)
    double stop_T;        // stop_candidate_threat_min_vote_share      (model.py:397)
    double stop_boost;    // stop_candidate_boost_factor               (model.py:414)
    float  f_st1p, f_bandwagon, f_bonus, f_pwf, f_fat_pen, f_stop_D, f_stop_boost, f_pad;
    int32_t need;         // smallest int v with (double)v >= thr*E    (simulate.py:175)
    int32_t max_rounds;
    int32_t rr;           // runoff_threshold_rounds                   (simulate.py:183)
    int32_t k;            // min(runoff_candidate_count, E)
    int32_t has_sticky;   // stickiness_factor > 0 and wiring == model
    int32_t has_bw;       // bandwagon_strength > 0 and wiring == model
    int32_t en_fat;       // candidate fatigue live (model wiring only)
    int32_t fat_rounds;
    int32_t fat_top_n;
    int32_t en_stop;      // stop-candidate live (model wiring only)
    int32_t fat_cnt;      // smallest vote count v with (double)v / E >= fatigue_vote_share_threshold: share < thr  <=>  votes < fat_cnt
    int32_t stop_cnt;     // smallest vote count v with (double)v / E >  stop_candidate_threat_min_vote_share: share > T <=> votes >= stop_cnt
};

struct RosterDev {
    int32_t E;            // electors == candidates
    int32_t Epad;         // E rounded up to a multiple of 32 (row pitch of [c][e] tables)
    int32_t G;            // number of distinct regions (dense codes 0..G-1)
    const double*  x64;   // [E]   ideology_score
    const float*   x32;   // [E]
    const int32_t* region;// [E]   dense region code
    const uint8_t* pap;   // [E]   is_papabile
};

struct TrialOut {
    int32_t* winner;        // [n]
    int32_t* rounds;        // [n]
    uint8_t* reason;        // [n] or null
    int32_t* final_votes;   // [n, E] or null
    int32_t* round_tallies; // [n, R, E] or null
    unsigned long long* winner_hist;  // [sets, E+1] or null
    unsigned long long* rounds_hist;  // [sets, R+1] or null
    unsigned long long* totals;       // [sets, 2] or null
};

struct BatchDev {
    const DevSet* sets;     // [n_sets]
    int32_t  n_sets;
    int32_t  R;             // max_rounds (same for every set of a batch)
    uint64_t n_sims;        // per set
    uint64_t sim_id_begin;
    uint32_t seed_lo, seed_hi;
    const double* tape;     // [n, R*E] or null
};

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw — SC'11).  Stream layout shared
// with the oracle: key = (seed_lo, seed_hi);
//   call A: counter = (sim_lo, sim_hi, round, e >> 2)               -> word (e & 3) = w_a(e)
//   call B: counter = (sim_lo, sim_hi, round, (e >> 2) | 0x80000000) -> word (e & 3) = w_b(e)
//   fp64 uniform  u = ((w_a >> 5) * 2^26 + (w_b >> 6)) / 2^53   (the construction numpy's
//                     MT19937 random_sample uses, so tape and native modes share one path)
//   fp32 uniform  u = (w_a >> 8) * 2^-24  (= the leading 24 bits of the fp64 uniform)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;   // IMAD.WIDE
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u53_from_words(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ float u24_from_word(uint32_t a) {
    return (float)(a >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------
// Warp-level helpers (one warp owns one trial).
// ---------------------------------------------------------------------------
// max tally and the LOWEST index attaining it (idxmax, simulate.py:174-176).
__device__ __forceinline__ void warp_argmax_first(const int32_t* tally, int C, int lane, int& mx, int& arg,
                                                  const uint8_t* excluded = nullptr) {
    int bm = -1, ba = 0x7fffffff;
    for (int c = lane; c < C; c += CSIM_WARP) {
        if (excluded && excluded[c]) continue;
        const int v = tally[c];
        if (v > bm) { bm = v; ba = c; }          // strided scan keeps the lowest c per lane on ties
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const int om = __shfl_xor_sync(CSIM_FULL, bm, off), oa = __shfl_xor_sync(CSIM_FULL, ba, off);
        if (om > bm || (om == bm && oa < ba)) { bm = om; ba = oa; }
    }
    mx = bm; arg = ba;
}

// First k of a stable descending sort of the tally (ties -> lowest index):
// Series.nlargest(k) as used at simulate.py:185.  `mark` is a [C] scratch of bytes.
__device__ __forceinline__ void warp_top_k(const int32_t* tally, int C, int k, int lane, int32_t* eff, uint8_t* mark) {
    for (int c = lane; c < C; c += CSIM_WARP) mark[c] = 0;
    __syncwarp();
    for (int j = 0; j < k; ++j) {
        int mx, arg;
        warp_argmax_first(tally, C, lane, mx, arg, mark);
        if (lane == 0) { eff[j] = arg; mark[arg] = 1; }
        __syncwarp();
    }
}

// Top two tallies by (count desc, index asc) as packed keys (count << 16 | 0xFFFF - index); needs C < 32768.
__device__ __forceinline__ void warp_top2(const int32_t* tally, int C, int lane, uint32_t& k1, uint32_t& k2) {
    uint32_t a = 0, b = 0;
    for (int c = lane; c < C; c += CSIM_WARP) {
        const uint32_t key = ((uint32_t)tally[c] << 16) | (uint32_t)(0xFFFF - c);
        if (key > a) { b = a; a = key; } else if (key > b) { b = key; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const uint32_t oa = __shfl_xor_sync(CSIM_FULL, a, off), ob = __shfl_xor_sync(CSIM_FULL, b, off);
        const uint32_t hi = max(a, oa), lo = min(a, oa);
        b = max(lo, max(b, ob));
        a = hi;
    }
    k1 = a; k2 = b;
}


// ---------------------------------------------------------------------------
// Bulk asynchronous copy global -> shared (the TMA engine's 1-D form, SASS UBLKCP) completing on an mbarrier:
// ONE thread issues the copy of a whole table, every thread of the CTA waits on the barrier's phase.
// Addresses are shared-window addresses; `bytes` must be a multiple of 16 and both ends 16-byte aligned.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // order earlier generic-proxy accesses of dst before the async write
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
}

// ---------------------------------------------------------------------------
// Dealing the electors of one round to the 32 lanes of the warp that owns the trial (PREFIX and FULL engines), each with
// its fp32 uniform: lane l owns whole Philox quads (electors 4q .. 4q+3, q = l + 32 qi: one Philox call, four uniforms) for
// the electors below e_quad_end, then the tail goes one elector per lane.  fn(e, u, slot) — slot = the elector's row in
// tables stored in lane-visiting order (prefix_slot) — must not contain warp collectives: lanes run out of electors
// at different times.
// ---------------------------------------------------------------------------
template <class Fn>
__device__ __forceinline__ void deal_electors(int lane, int E, int e_quad_end, int r, uint32_t s_lo, uint32_t s_hi, uint32_t seed_lo,
                                              uint32_t seed_hi, Fn&& fn) {
    const int n_qit = e_quad_end >> 7;                       // quad iterations (128 electors each)
    for (int qi = 0; qi < n_qit; ++qi) {
        const int q = lane + 32 * qi;
        if (4 * q < E) {
            uint32_t w[4];
            philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)q, seed_lo, seed_hi, w);
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
                const int e = 4 * q + i;
                const uint32_t wa = i == 0 ? w[0] : (i == 1 ? w[1] : (i == 2 ? w[2] : w[3]));
                if (e < E) fn(e, u24_from_word(wa), (qi << 7) + (i << 5) + lane);
            }
        }
    }
    for (int e = e_quad_end + lane; e < E; e += 32) {
        uint32_t w[4];
        philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)(e >> 2), seed_lo, seed_hi, w);
        const int sel = e & 3;
        const uint32_t wa = sel == 0 ? w[0] : (sel == 1 ? w[1] : (sel == 2 ? w[2] : w[3]));
        fn(e, u24_from_word(wa), e);
    }
}

// The same dealing as ONE loop over the lane's elector slots, so that the caller's body is instantiated once (deal_electors
// inlines it twice: the FULL engine's body is ~25 KB of code and two copies thrash the 32 KB instruction cache).
template <class Fn>
__device__ __forceinline__ void deal_electors_once(int lane, int E, int e_quad_end, int r, uint32_t s_lo, uint32_t s_hi, uint32_t seed_lo,
                                                   uint32_t seed_hi, Fn&& fn) {
    const int n_q = (e_quad_end >> 7) << 2;                  // quad slots of a lane
    const int n_tail = E > e_quad_end + lane ? (E - e_quad_end - lane + 31) >> 5 : 0;
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll 1
    for (int n = 0; n < n_q + n_tail; ++n) {
        int e, slot, sel;
        if (n < n_q) {
            const int qi = n >> 2, q = lane + 32 * qi;
            sel = n & 3;
            e = 4 * q + sel;
            slot = (qi << 7) + (sel << 5) + lane;
            if (sel == 0 && e < E) philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)q, seed_lo, seed_hi, w);
        } else {
            e = e_quad_end + lane + 32 * (n - n_q);
            slot = e; sel = e & 3;
            philox4x32_10(s_lo, s_hi, (uint32_t)r, (uint32_t)(e >> 2), seed_lo, seed_hi, w);
        }
        const uint32_t wa = sel == 0 ? w[0] : (sel == 1 ? w[1] : (sel == 2 ? w[2] : w[3]));
        if (e < E) fn(e, u24_from_word(wa), slot);
    }
}

// ---------------------------------------------------------------------------
// Dynamic trial scheduling of the warp-per-trial kernels.  Trials differ in length (6 ... max_rounds rounds), so a static
// split leaves most SMs idle while the unluckiest warps finish (12 % of the warp slots of the round-2 dense kernel were
// empty over its duration).  Instead: one 64-bit counter per parameter set in global memory (zeroed before the launch); a
// warp takes the next `chunk` trials of the set its CTA is on with ONE atomicAdd by lane 0; a CTA walks the sets in a
// rotated order (so that CTAs start on different sets), stages a set's tables once, and moves on when the set's counter has
// run out.  Results never depend on the schedule: a trial's random stream is keyed by its sim id.
// ---------------------------------------------------------------------------
struct WorkQueue {
    unsigned long long* next;   // [n_sets] next unclaimed trial of each set
    int chunk;                  // trials a warp claims per fetch
};

__device__ __forceinline__ uint64_t warp_claim_trials(const WorkQueue& wq, int set, int lane) {
    unsigned long long v = 0;
    if (lane == 0) v = atomicAdd(wq.next + set, (unsigned long long)wq.chunk);
    return __shfl_sync(CSIM_FULL, v, 0);
}

// the set a CTA visits in its si-th step, and whether that set still has unclaimed trials (thread 0 peeks; the caller
// publishes the answer through shared memory around its own __syncthreads)
__device__ __forceinline__ int cta_set_of_step(int si, int n_sets) {
    const int set0 = (int)(((uint64_t)blockIdx.x * (uint64_t)n_sets) / gridDim.x);
    const int set = set0 + si;
    return set >= n_sets ? set - n_sets : set;
}
__device__ __forceinline__ bool set_exhausted(const WorkQueue& wq, int set, uint64_t n_sims) {
    return *reinterpret_cast<volatile unsigned long long*>(wq.next + set) >= n_sims;
}

// The per-trial outputs every warp-per-trial kernel writes when its trial ends (winner / rounds / reason rows, the last
// tally, the -1 padding of the per-round tallies beyond `rounds`, the two histograms of simulate.py:356-387).
__device__ __forceinline__ void write_trial_outputs(const TrialOut& out, uint64_t t, int set, int C, int R, int lane, const int32_t* tally,
                                                    int winner, int rounds, int reason) {
    if (out.final_votes) for (int c = lane; c < C; c += 32) out.final_votes[t * (uint64_t)C + c] = tally[c];
    if (out.round_tallies)
        for (size_t i = (size_t)rounds * C + lane; i < (size_t)R * C; i += 32) out.round_tallies[t * (uint64_t)R * C + i] = -1;
    if (lane == 0) {
        out.winner[t] = winner; out.rounds[t] = rounds;
        if (out.reason) out.reason[t] = (uint8_t)reason;
        if (out.winner_hist) atomicAdd(&out.winner_hist[(size_t)set * (C + 1) + (winner >= 0 ? winner : C)], 1ULL);
        if (out.rounds_hist && winner >= 0) atomicAdd(&out.rounds_hist[(size_t)set * (R + 1) + rounds], 1ULL);
    }
}
