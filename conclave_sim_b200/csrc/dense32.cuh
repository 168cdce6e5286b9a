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
//
// Factorised evaluation (template flag FACT, the default whenever beta_r >= 0 and
// beta_r * half-range(x) <= 40): with x' = x - x_mid,
//     exp(-beta |x_e - x_c|) = min( e^{-beta x'_e} e^{+beta x'_c} ,  e^{+beta x'_e} e^{-beta x'_c} )
// so a pair costs two multiplies and a min instead of FADD + FMUL + MUFU.EX2, and the multiplies and
// the accumulation run as sm_100 packed-fp32 instructions (FMUL2 / FADD2, two pairs per issue slot):
// 3 issue slots per pair including the shared-memory operand loads.  The four per-round vectors
// A = e^{-beta x'}, A' = e^{+beta x'}, BP = A' * pw, BP' = A * pw come from k_round_tables32 (fp64 exp,
// rounded once).  FACT = false keeps the MUFU.EX2 form (any roster, any beta).
#pragma once
#include "common.cuh"

struct Dense32Args {
    RosterDev ro;
    BatchDev b;
    TrialOut out;
    const float* nbeta;     // [sets][R]  -beta_r * log2(e)
    const float* tab;       // [sets][R][4][NCH*L]  A, A', BP, BP' (FACT only)
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
    return ((CP * 4 /*xs*/ + CP * 4 /*reg*/ + (size_t)NCH * G * 4 * 2 /*regcnt, regmask*/) + 15) & ~(size_t)15;
}
template <int NCH>
__host__ __device__ inline size_t dense32_warp_smem(int E, int L, int kmax, bool model) {
    const size_t CP = (size_t)NCH * L;
    size_t n = CP * 4 * 2 /*pw | bp, bpp*/ + (size_t)E * 4 /*tally*/ + (size_t)kmax * 4 /*eff*/ + (((size_t)E + 3) & ~(size_t)3) /*mark*/;
    if (model) n += (size_t)E * 4 * 2 /*prev_tally, votes*/ + CP * 4 /*bw*/ + (size_t)((NCH + 3) & ~3) * 4 /*bwchunk*/;
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
__global__ void k_round_tables32(RosterDev ro, const DevSet* sets, const double* betas, int R, int CP, double xm, float* tab) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int sr = blockIdx.y;
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

// every term of one score, inline (run-off rounds and the generic re-scan)
template <bool MODEL, bool FACT>
__device__ __forceinline__ float full_score32(int e, int c, float xe, float ae, float ape, int ge, float nb, float bonus, float st1p,
                                              int prev, bool use_bw, const float* xs, const int32_t* reg, const float* pw,
                                              const float* bpp, const float* bw) {
    // FACT: pw = BP, bpp = BP';  else pw = per-candidate papabile weight
    float v = FACT ? fminf(ae * pw[c], ape * bpp[c]) : ex2_approx(fabsf(xe - xs[c]) * nb) * pw[c];
    if (reg[c] == ge) v += bonus;
    if (MODEL) {
        if (use_bw) v += bw[c];
        if (c == prev) v *= st1p;
    }
    return c == e ? 0.0f : v;
}

// Top two tallies by (count desc, index asc) as packed keys (count << 16 | 0xFFFF - index); needs C < 65536.
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

struct LaneCtx32 {              // per-warp / per-round constants of the draw
    const float* xs; const int32_t* reg; const float* regcnt; const uint32_t* regmask;
    const float* pw; const float* bpp; const float* bw; const float* bwchunk;   // FACT: pw = BP, bpp = BP'
    float nb, bonus, st1p;
    int C, G, L;
    bool use_bw;
};

// One full-field draw, chunk length 8 (E <= 136): phase 1 + chunk search + vectorised re-scan.
template <int NCH, bool MODEL, bool FACT>
__device__ __forceinline__ int draw_full_lc8(const LaneCtx32& k, int e, float u, int prev, float ae, float ape) {
    const float xe = FACT ? 0.f : k.xs[e];
    const int ge = k.reg[e];
    const float nb = k.nb;
    const int je = e >> 3;
    // the self score, removed from its chunk (model.py:418-424); same expression as phase 1 so it cancels
    float see = (FACT ? fminf(ae * k.pw[e], ape * k.bpp[e]) : k.pw[e]) + k.bonus;
    if (MODEL && k.use_bw) see += k.bw[e];
    int jp = -1; float sadd = 0.f;
    if (MODEL && prev >= 0 && prev != e) {               // stickiness: + st * s[e, prev] in prev's chunk (model.py:357-367)
        float sp = FACT ? fminf(ae * k.pw[prev], ape * k.bpp[prev]) : ex2_approx(fabsf(xe - k.xs[prev]) * nb) * k.pw[prev];
        if (k.reg[prev] == ge) sp += k.bonus;
        if (k.use_bw) sp += k.bw[prev];
        sadd = sp * (k.st1p - 1.0f);
        jp = prev >> 3;
    }
    // ---------------- phase 1: chunk sums -----------------------------------------------------------
    float cs[NCH];
    float S = 0.f;
    const unsigned long long a2 = f2_pack(ae, ae), ap2 = f2_pack(ape, ape);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        float acc0 = (j == je) ? -see : 0.f;
        float acc1 = (MODEL && j == jp) ? sadd : 0.f;
        if (FACT) {
            unsigned long long acc = f2_pack(acc0, acc1);
            const ulonglong2* b4 = reinterpret_cast<const ulonglong2*>(k.pw + j * 8);
            const ulonglong2* c4 = reinterpret_cast<const ulonglong2*>(k.bpp + j * 8);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const ulonglong2 b = b4[i], c = c4[i];
                unsigned long long t1 = f2_mul(a2, b.x), t2 = f2_mul(ap2, c.x);
                acc = f2_add(acc, f2_pack(fminf(f2_lo(t1), f2_lo(t2)), fminf(f2_hi(t1), f2_hi(t2))));
                t1 = f2_mul(a2, b.y); t2 = f2_mul(ap2, c.y);
                acc = f2_add(acc, f2_pack(fminf(f2_lo(t1), f2_lo(t2)), fminf(f2_hi(t1), f2_hi(t2))));
            }
            acc0 = f2_lo(acc); acc1 = f2_hi(acc);
        } else {
            const float4* x4 = reinterpret_cast<const float4*>(k.xs + j * 8);
            const float4* p4 = reinterpret_cast<const float4*>(k.pw + j * 8);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float4 xv = x4[i], pv = p4[i];
                acc0 = fmaf(ex2_approx(fabsf(xe - xv.x) * nb), pv.x, acc0);
                acc1 = fmaf(ex2_approx(fabsf(xe - xv.y) * nb), pv.y, acc1);
                acc0 = fmaf(ex2_approx(fabsf(xe - xv.z) * nb), pv.z, acc0);
                acc1 = fmaf(ex2_approx(fabsf(xe - xv.w) * nb), pv.w, acc1);
            }
        }
        float s = fmaf(k.bonus, k.regcnt[j * k.G + ge], acc0 + acc1);    // regional bonus of the chunk (model.py:322-323)
        if (MODEL && k.use_bw) s += k.bwchunk[j];                        // bandwagon of the chunk (model.py:354)
        cs[j] = s;
        S += s;
    }
    if (!(S > 0.0f)) {            // all scores vanished: the reference makes the row uniform INCLUDING self (model.py:447-449)
        const int j = (int)(u * (float)k.C);
        return j < k.C ? j : k.C - 1;
    }
    // ---------------- phase 2a: the chunk whose prefix crosses u * S --------------------------------
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
    const int last_chunk = (k.C - 1) >> 3;
    jstar = jstar > last_chunk ? last_chunk : jstar;
    // ---------------- phase 2b: re-scan that chunk, every term inline, roster order ------------------
    const int c0 = jstar * 8;
    // FACT: (xa, xb) hold BP', (pa, pb) hold BP;  else x and pw
    const float* src_x = FACT ? k.bpp : k.xs;
    const float4 xa = *reinterpret_cast<const float4*>(src_x + c0), xb = *reinterpret_cast<const float4*>(src_x + c0 + 4);
    const float4 pa = *reinterpret_cast<const float4*>(k.pw + c0), pb = *reinterpret_cast<const float4*>(k.pw + c0 + 4);
    const uint32_t rm = k.regmask[jstar * k.G + ge];                     // bit i: candidate c0+i is of e's region
    const uint32_t dm = (jstar == je) ? (1u << (e & 7)) : 0u;            // bit i: candidate c0+i is e itself
    uint32_t pm = 0;
    float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
    if (MODEL) {
        if (prev >= 0 && (prev >> 3) == jstar) pm = 1u << (prev & 7);
        if (k.use_bw) { ba = *reinterpret_cast<const float4*>(k.bw + c0); bb = *reinterpret_cast<const float4*>(k.bw + c0 + 4); }
    }
    float cum = below;
    int cnt = 0;
#define CSIM_RESCAN(i, xv, pv, bv)                                              \
    {                                                                           \
        const float ex = FACT ? fminf(ae * (pv), ape * (xv)) : ex2_approx(fabsf(xe - (xv)) * nb) * (pv); \
        float v = ex + (((rm >> (i)) & 1u) ? k.bonus : 0.0f);                   \
        if (MODEL) { v += (bv); if ((pm >> (i)) & 1u) v *= k.st1p; }            \
        if ((dm >> (i)) & 1u) v = 0.0f;                                         \
        cum += v;                                                               \
        cnt += (cum <= tgt) ? 1 : 0;                                            \
    }
    CSIM_RESCAN(0, xa.x, pa.x, ba.x) CSIM_RESCAN(1, xa.y, pa.y, ba.y) CSIM_RESCAN(2, xa.z, pa.z, ba.z) CSIM_RESCAN(3, xa.w, pa.w, ba.w)
    CSIM_RESCAN(4, xb.x, pb.x, bb.x) CSIM_RESCAN(5, xb.y, pb.y, bb.y) CSIM_RESCAN(6, xb.z, pb.z, bb.z) CSIM_RESCAN(7, xb.w, pb.w, bb.w)
#undef CSIM_RESCAN
    int vote = c0 + cnt;                                 // == first c with cum > tgt (searchsorted 'right', simulate.py:153)
    const int maxc = min(c0 + 7, k.C - 1);
    vote = vote > maxc ? maxc : vote;                    // rounding at the chunk's upper edge
    if (vote == e) vote = e > 0 ? e - 1 : (k.C > 1 ? 1 : 0);
    return vote;
}

// Generic chunk length (runtime L, multiple of 4): same algorithm, scalar re-scan.
template <int NCH, bool MODEL, bool FACT>
__device__ __forceinline__ int draw_full_generic(const LaneCtx32& k, int e, float u, int prev, float ae, float ape) {
    const int L = k.L, C = k.C;
    const float xe = FACT ? 0.f : k.xs[e];
    const int ge = k.reg[e];
    const float nb = k.nb;
    const int je = e / L;
    float see = (FACT ? fminf(ae * k.pw[e], ape * k.bpp[e]) : k.pw[e]) + k.bonus;
    if (MODEL && k.use_bw) see += k.bw[e];
    int jp = -1; float sadd = 0.f;
    if (MODEL && prev >= 0 && prev != e) {
        float sp = FACT ? fminf(ae * k.pw[prev], ape * k.bpp[prev]) : ex2_approx(fabsf(xe - k.xs[prev]) * nb) * k.pw[prev];
        if (k.reg[prev] == ge) sp += k.bonus;
        if (k.use_bw) sp += k.bw[prev];
        sadd = sp * (k.st1p - 1.0f);
        jp = prev / L;
    }
    float cs[NCH];
    float S = 0.f;
    const unsigned long long a2 = f2_pack(ae, ae), ap2 = f2_pack(ape, ape);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        float acc0 = (j == je) ? -see : 0.f;
        float acc1 = (MODEL && j == jp) ? sadd : 0.f;
        if (FACT) {
            unsigned long long acc = f2_pack(acc0, acc1);
            const ulonglong2* b4 = reinterpret_cast<const ulonglong2*>(k.pw + j * L);
            const ulonglong2* c4 = reinterpret_cast<const ulonglong2*>(k.bpp + j * L);
#pragma unroll 2
            for (int i = 0; i < (L >> 2); ++i) {
                const ulonglong2 b = b4[i], c = c4[i];
                unsigned long long t1 = f2_mul(a2, b.x), t2 = f2_mul(ap2, c.x);
                acc = f2_add(acc, f2_pack(fminf(f2_lo(t1), f2_lo(t2)), fminf(f2_hi(t1), f2_hi(t2))));
                t1 = f2_mul(a2, b.y); t2 = f2_mul(ap2, c.y);
                acc = f2_add(acc, f2_pack(fminf(f2_lo(t1), f2_lo(t2)), fminf(f2_hi(t1), f2_hi(t2))));
            }
            acc0 = f2_lo(acc); acc1 = f2_hi(acc);
        } else {
            const float4* x4 = reinterpret_cast<const float4*>(k.xs + j * L);
            const float4* p4 = reinterpret_cast<const float4*>(k.pw + j * L);
            for (int i = 0; i < (L >> 2); ++i) {
                const float4 xv = x4[i], pv = p4[i];
                acc0 = fmaf(ex2_approx(fabsf(xe - xv.x) * nb), pv.x, acc0);
                acc1 = fmaf(ex2_approx(fabsf(xe - xv.y) * nb), pv.y, acc1);
                acc0 = fmaf(ex2_approx(fabsf(xe - xv.z) * nb), pv.z, acc0);
                acc1 = fmaf(ex2_approx(fabsf(xe - xv.w) * nb), pv.w, acc1);
            }
        }
        float s = fmaf(k.bonus, k.regcnt[j * k.G + ge], acc0 + acc1);
        if (MODEL && k.use_bw) s += k.bwchunk[j];
        cs[j] = s;
        S += s;
    }
    if (!(S > 0.0f)) {
        const int j = (int)(u * (float)C);
        return j < C ? j : C - 1;
    }
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
    jstar = jstar > last_chunk ? last_chunk : jstar;
    const int c0 = jstar * L;
    float cum = below;
    int cnt = 0;
    for (int i = 0; i < L; ++i) {
        const int c = c0 + i;
        const float v = c < C ? full_score32<MODEL, FACT>(e, c, xe, ae, ape, ge, nb, k.bonus, k.st1p, prev, k.use_bw, k.xs, k.reg, k.pw, k.bpp, k.bw) : 0.0f;
        cum += v;
        cnt += (cum <= tgt) ? 1 : 0;
    }
    int vote = c0 + cnt;
    const int maxc = min(c0 + L - 1, C - 1);
    vote = vote > maxc ? maxc : vote;
    if (vote == e) vote = e > 0 ? e - 1 : (C > 1 ? 1 : 0);
    return vote;
}

template <int NCH, int LC, bool MODEL, bool FACT>
__global__ void __launch_bounds__(256, LC == 8 ? 3 : 1) k_trials_dense32(Dense32Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int E = a.ro.E, C = E, R = a.b.R, G = a.ro.G;
    const int L = LC > 0 ? LC : a.L;
    const int CP = NCH * L;
    // ---- CTA-shared roster tables ------------------------------------------------------------
    float* xs = (float*)smem_raw;
    int32_t* reg = (int32_t*)(xs + CP);
    float* regcnt = (float*)(reg + CP);
    uint32_t* regmask = (uint32_t*)(regcnt + NCH * G);
    for (int c = threadIdx.x; c < CP; c += blockDim.x) {
        xs[c] = c < C ? a.ro.x32[c] : 0.0f;
        reg[c] = c < C ? a.ro.region[c] : -1;
    }
    for (int i = threadIdx.x; i < NCH * G; i += blockDim.x) {
        const int j = i / G, g = i - j * G;
        int n = 0; uint32_t m = 0;
        for (int c = j * L; c < (j + 1) * L && c < C; ++c)
            if (a.ro.region[c] == g) { ++n; if (c - j * L < 32) m |= 1u << (c - j * L); }
        regcnt[i] = (float)n;
        regmask[i] = m;
    }
    // ---- per-warp state ----------------------------------------------------------------------
    unsigned char* base = smem_raw + a.cta_smem + (size_t)wib * a.warp_smem;
    // 16-byte aligned float arrays first (they are read with 128-bit loads), then the int arrays
    float* pw = (float*)base;                   base += sizeof(float) * CP;      // FACT: BP of the current round
    float* bpp = (float*)base;                  base += sizeof(float) * CP;      // FACT: BP'
    float* bw = nullptr; float* bwchunk = nullptr;
    if (MODEL) {
        bw = (float*)base;                      base += sizeof(float) * CP;
        bwchunk = (float*)base;                 base += sizeof(float) * ((NCH + 3) & ~3);
    }
    int32_t* tally = (int32_t*)base;            base += sizeof(int32_t) * E;
    int32_t* prev_tally = nullptr; int32_t* votes = nullptr;
    if (MODEL) {
        prev_tally = (int32_t*)base;            base += sizeof(int32_t) * E;
        votes = (int32_t*)base;                 base += sizeof(int32_t) * E;
    }
    int32_t* eff = (int32_t*)base;              base += sizeof(int32_t) * a.kmax;
    uint8_t* mark = (uint8_t*)base;
    __syncthreads();   // the only block barrier: roster tables are read-only from here on

    const uint64_t n_trials = (uint64_t)a.b.n_sets * a.b.n_sims;
    const uint64_t warp_global = (uint64_t)blockIdx.x * a.warps_per_cta + wib;
    const uint64_t n_warps = (uint64_t)gridDim.x * a.warps_per_cta;
    const int m_quads = E >> 7;                              // whole quads per lane
    const int n_main = 4 * m_quads;
    const int e_single0 = 128 * m_quads + lane;
    const int n_single = e_single0 < E ? (E - e_single0 + 31) >> 5 : 0;
    const int n_slots = n_main + n_single;
    const bool top2_ok = C >= 2 && C < 32768;

    LaneCtx32 kx;
    kx.xs = xs; kx.reg = reg; kx.regcnt = regcnt; kx.regmask = regmask; kx.pw = pw; kx.bpp = bpp; kx.bw = bw; kx.bwchunk = bwchunk;
    kx.C = C; kx.G = G; kx.L = L; kx.use_bw = false; kx.nb = 0.f; kx.bonus = 0.f; kx.st1p = 1.f;

    int cur_set = -1;
    float bw_strength = 0.f;
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
        const bool fast2 = (k == 2) && top2_ok;                  // two-candidate run-off: O(1) closing
        bool have_eff = false, have_prev = false;
        int winner = -1, rounds = 0, reason = CSIM_REASON_MAX_ROUNDS;
        int ef0 = 0, ef1 = 0;                                    // run-off pair when fast2

        for (int r = 1; r <= R; ++r) {
            rounds = r;
            kx.nb = __ldg(&nbeta[r - 1]);
            kx.use_bw = false;
            const float* tabA = nullptr;
            if (FACT) {                                          // stage this round's BP / BP' (trial-invariant vectors)
                tabA = a.tab + ((size_t)set * R + (r - 1)) * 4 * CP;
                __syncwarp();
                for (int c = lane; c < CP; c += 32) { pw[c] = __ldg(tabA + 2 * CP + c); bpp[c] = __ldg(tabA + 3 * CP + c); }
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
            int cnt0 = 0;                                        // run-off (fast2): this lane's votes for ef0
            for (int n = 0; n < n_slots; ++n) {
                int e, q; bool gen;
                if (n < n_main) { q = lane + 32 * (n >> 2); e = 4 * q + (n & 3); gen = (n & 3) == 0; }
                else { e = e_single0 + 32 * (n - n_main); q = e >> 2; gen = true; }
                if (gen) philox4x32_10((uint32_t)sim, (uint32_t)(sim >> 32), (uint32_t)r, (uint32_t)q, a.b.seed_lo, a.b.seed_hi, w);
                const int sel = e & 3;
                const uint32_t wa = sel == 0 ? w[0] : (sel == 1 ? w[1] : (sel == 2 ? w[2] : w[3]));
                const float u = u24_from_word(wa);
                const int prev = (MODEL && have_prev && has_sticky) ? votes[e] : -1;
                float ae = 0.f, ape = 0.f;
                if (FACT) { ae = __ldg(tabA + e); ape = __ldg(tabA + CP + e); }
                int vote;
                if (!have_eff) {
                    if (LC == 8) vote = draw_full_lc8<NCH, MODEL, FACT>(kx, e, u, prev, ae, ape);
                    else vote = draw_full_generic<NCH, MODEL, FACT>(kx, e, u, prev, ae, ape);
                    atomicAdd(&tally[vote], 1);
                } else if (fast2) {
                    // ---------------- two-candidate run-off: columns 0 and 1, renormalised (simulate.py:144-154)
                    const float xe = FACT ? 0.f : xs[e]; const int ge = reg[e];
                    const float v0 = full_score32<MODEL, FACT>(e, 0, xe, ae, ape, ge, kx.nb, kx.bonus, kx.st1p, prev, kx.use_bw, xs, reg, pw, bpp, bw);
                    const float v1 = full_score32<MODEL, FACT>(e, 1, xe, ae, ape, ge, kx.nb, kx.bonus, kx.st1p, prev, kx.use_bw, xs, reg, pw, bpp, bw);
                    const float sum = v0 + v1;
                    int j;
                    if (!(sum > 0.0f)) j = (u * 2.0f >= 1.0f) ? 1 : 0;       // simulate.py:151 uniform fallback
                    else j = (v0 > u * sum || !(v1 > 0.0f)) ? 0 : 1;
                    cnt0 += (j == 0);
                    vote = j == 0 ? ef0 : ef1;
                } else {
                    // ---------------- k-candidate run-off (generic) --------------------------------------------
                    const float xe = FACT ? 0.f : xs[e]; const int ge = reg[e];
                    float sum = 0.f;
                    for (int j = 0; j < k; ++j)
                        sum += full_score32<MODEL, FACT>(e, j, xe, ae, ape, ge, kx.nb, kx.bonus, kx.st1p, prev, kx.use_bw, xs, reg, pw, bpp, bw);
                    int jsel;
                    if (!(sum > 0.0f)) {
                        jsel = (int)(u * (float)k);
                    } else {
                        const float tgt = u * sum;
                        float cum = 0.f; jsel = -1; int lastpos = 0;
                        for (int j = 0; j < k; ++j) {
                            const float v = full_score32<MODEL, FACT>(e, j, xe, ae, ape, ge, kx.nb, kx.bonus, kx.st1p, prev, kx.use_bw, xs, reg, pw, bpp, bw);
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
                if (MODEL) votes[e] = vote;      // read only by this lane for its own electors: in-place update is safe
            }
            int n0 = 0, n1 = 0;
            if (have_eff && fast2) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) cnt0 += __shfl_xor_sync(CSIM_FULL, cnt0, off);
                n0 = cnt0; n1 = E - cnt0;
                if (lane == 0) { tally[ef0] = n0; tally[ef1] = n1; }
            }
            __syncwarp();
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
