// seqsum.cuh -- exact, parallel evaluation of a SEQUENTIAL float32 sum of non-negative terms.
//
// The reference's RMSNorm accumulates  ss = ((0 + x0*x0) + x1*x1) + ...  one float add at a
// time (InferenceCore.rmsnorm, InferenceCore.java:39-48 via FloatTensor.reduce,
// FloatTensor.java:110-116).  Float addition is not associative, so a tree reduction gives
// different bits, and a literal chain costs ~8-10 cycles per term on one thread (~20 us for
// dim 4096, more than the whole attention block).  This file reproduces the chain's result
// bit for bit with the work spread over 32 warps:
//
//   While the running sum s stays inside one binade [2^e, 2^(e+1)) its mantissa M is an integer
//   in units of u = 2^(e-23), and adding a term t rounds to  M + k + [f > 1/2]  with t/u = k + f
//   (on an exact tie, f == 1/2, round-half-even makes the increment depend on the parity of M; real
//   activations produce 3-30 ties per 4096 terms).  So inside a binade a step is the integer map
//   M -> M + a[M & 1], and such maps compose associatively.
//   The first quarter of the terms (where the sum crosses a binade every few terms) is summed
//   literally by one thread.  The rest is processed in groups of 32 (one warp, one term per lane).
//   Which binade a group starts and ends in is PREDICTED from a float prefix sum over the groups.
//   A tie-free group that stays in one binade is one integer (a warp reduction); a group with ties
//   is an ordered warp scan of parity pairs; a group containing a binade crossing is split at the
//   crossing steps, which are executed as real float adds.  Each warp owns a contiguous range of
//   groups and composes them, in order, into a handful of entries; one thread then walks the
//   ~20-40 entries and VERIFIES every prediction (exponent on entry, mantissa < 2^24 on exit).
//   A failed check falls back to the literal loop from that element on, so the result is exact
//   unconditionally; a misprediction (sum within ~1e-6 of a power of two) only costs time.
//
// This code runs as ONE CTA per launch, so its instruction footprint matters as much as its
// arithmetic: loops are not unrolled and the rare paths are out of line.
#pragma once
#include "common.cuh"

#define SEQSUM_SINGLE 0x7fff
#define SEQSUM_NONE (-100000)
#define SEQSUM_THREADS 1024 // 32 warps: every per-group step is a ~500-900 cycle dependent chain, so few groups per warp
#define SEQSUM_WL 16        // entries per warp list (more -> sequential fallback)
#define SEQSUM_HEAD_SHIFT 2 // literal head = the first quarter of the groups

struct SeqPair { // the step  M -> M + a[M & 1]
    unsigned a0, a1;
};
__device__ __forceinline__ SeqPair seq_compose(SeqPair L, SeqPair R) { // apply L, then R
    SeqPair o;
    o.a0 = L.a0 + ((L.a0 & 1u) ? R.a1 : R.a0);
    o.a1 = L.a1 + (((1u + L.a1) & 1u) ? R.a1 : R.a0);
    o.a0 = min(o.a0, 1u << 26); // saturate: anything >= 2^24 fails verification anyway
    o.a1 = min(o.a1, 1u << 26);
    return o;
}

struct SeqEntry { // cls: SEQSUM_SINGLE (a1 = float bits of the term) or the binade exponent of a composed run
    int cls;
    unsigned a0, a1;
    int start; // first element covered (for the fallback)
};

struct SeqSumScratch {
    float *gs;        // [ng+1] group sums, then predicted running sum before each group
    SeqEntry *wl;     // [NW][SEQSUM_WL] per-warp entry lists (each warp owns a contiguous range of groups)
    SeqEntry *flat;   // [NW*SEQSUM_WL] the same entries, compacted in order for the resolver
    int *wcnt;        // [NW] entries per warp
    float *result;    // [1]
    int *info;        // [4] diagnostics: {entries, first fallback element or -1, overflow flag}
};

__host__ __device__ inline size_t seqsum_scratch_bytes(int n) { // n = number of terms (multiple of 32)
    const size_t ng = (size_t)(n + 31) / 32 + 1;
    return ((ng * 4 + 15) & ~(size_t)15) + 2 * (SEQSUM_THREADS / 32) * SEQSUM_WL * sizeof(SeqEntry) + (SEQSUM_THREADS / 32) * 4 + 16 + 16;
}

__device__ __forceinline__ SeqSumScratch seqsum_carve(unsigned char *p, int n) { // p 16-byte aligned
    SeqSumScratch s;
    const size_t ng = (size_t)(n + 31) / 32 + 1;
    s.gs = reinterpret_cast<float *>(p); p += (ng * 4 + 15) & ~(size_t)15;
    s.wl = reinterpret_cast<SeqEntry *>(p); p += (SEQSUM_THREADS / 32) * SEQSUM_WL * sizeof(SeqEntry);
    s.flat = reinterpret_cast<SeqEntry *>(p); p += (SEQSUM_THREADS / 32) * SEQSUM_WL * sizeof(SeqEntry);
    s.wcnt = reinterpret_cast<int *>(p); p += (SEQSUM_THREADS / 32) * 4;
    s.result = reinterpret_cast<float *>(p); p += 16;
    s.info = reinterpret_cast<int *>(p);
    return s;
}

__device__ __forceinline__ int f32_exponent(float f) { return (int)((__float_as_uint(f) >> 23) & 0xffu) - 127; }

// Parity pair of adding t to a running sum in binade e (ulp 2^(e-23)).  Returns false when t >= 2^(e+1)
// (not a within-binade step).
__device__ __forceinline__ bool seq_pair(float t, int e, SeqPair &pr) {
    const unsigned tb = __float_as_uint(t);
    const int et = (int)(tb >> 23);
    pr.a0 = pr.a1 = 0u;
    if (et == 0) return true; // zero / denormal term: far below half an ulp (e >= -90)
    const unsigned m = (tb & 0x7fffffu) | 0x800000u;
    int shift = (e + 127) - et; // t / ulp = m * 2^-shift
    if (shift < 0) return false;
    if (shift > 25) shift = 25;
    const unsigned k = m >> shift;
    const unsigned rem = m & ((1u << shift) - 1u);
    const unsigned half = shift ? (1u << (shift - 1)) : 0u;
    if (shift && rem == half) { // exact tie: the result mantissa M + k + r must be even
        pr.a0 = k + (k & 1u);
        pr.a1 = k + ((k + 1u) & 1u);
    } else {
        pr.a0 = pr.a1 = k + ((shift && rem > half) ? 1u : 0u);
    }
    return true;
}

__device__ __forceinline__ unsigned warp_sum_u(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Ordered composition of the pairs of lanes [lo, hi); the result is returned to every lane.
__device__ __noinline__ SeqPair warp_compose_range(SeqPair p, int lo, int hi, int lane) {
    SeqPair v;
    v.a0 = (lane >= lo && lane < hi) ? p.a0 : 0u;
    v.a1 = (lane >= lo && lane < hi) ? p.a1 : 0u;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        SeqPair o;
        o.a0 = __shfl_up_sync(0xffffffffu, v.a0, d);
        o.a1 = __shfl_up_sync(0xffffffffu, v.a1, d);
        if (lane >= d) v = seq_compose(o, v);
    }
    SeqPair r;
    r.a0 = __shfl_sync(0xffffffffu, v.a0, hi - 1);
    r.a1 = __shfl_sync(0xffffffffu, v.a1, hi - 1);
    return r;
}

// Warp-uniform accumulator of the current run + this warp's entry list.
struct SeqAcc {
    int cls, start, n;
    SeqPair pr;
    bool overflow;
};
__device__ __forceinline__ void seq_flush(SeqAcc &A, SeqEntry *wl, int lane) {
    if (A.cls != SEQSUM_NONE) {
        if (A.n < SEQSUM_WL) { if (lane == 0) { wl[A.n].cls = A.cls; wl[A.n].a0 = A.pr.a0; wl[A.n].a1 = A.pr.a1; wl[A.n].start = A.start; } }
        else A.overflow = true;
        A.n++;
        A.cls = SEQSUM_NONE;
    }
}
__device__ __forceinline__ void seq_push_run(SeqAcc &A, SeqEntry *wl, int lane, int cls, SeqPair pr, int start) {
    if (A.cls == cls) A.pr = seq_compose(A.pr, pr);
    else { seq_flush(A, wl, lane); A.cls = cls; A.pr = pr; A.start = start; }
}
__device__ __forceinline__ void seq_push_single(SeqAcc &A, SeqEntry *wl, int lane, float t, int idx) {
    seq_flush(A, wl, lane);
    if (A.n < SEQSUM_WL) { if (lane == 0) { wl[A.n].cls = SEQSUM_SINGLE; wl[A.n].a0 = 0u; wl[A.n].a1 = __float_as_uint(t); wl[A.n].start = idx; } }
    else A.overflow = true;
    A.n++;
}

// A group that contains a binade crossing (or whose prediction is inconsistent): per-lane prediction,
// runs between the crossing steps composed in order, crossing steps kept as single float adds.  Rare
// (about one group per warp) and deliberately out of line.
__device__ __noinline__ void seq_cross_group(SeqAcc &A, SeqEntry *wl, int lane, float tt, float before_g, int base) {
    float incl = tt;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        float o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    const float bl = before_g + (incl - tt), al = before_g + incl;
    const int ebl = f32_exponent(bl), eal = f32_exponent(al);
    SeqPair pl;
    pl.a0 = pl.a1 = 0u;
    const bool plain = (eal == ebl) && ebl >= -90 && (tt < INFINITY) && seq_pair(tt, ebl, pl);
    unsigned singles = __ballot_sync(0xffffffffu, !plain);
    int start = 0;
    while (true) {
        const int L = singles ? (__ffs(singles) - 1) : 32;
        if (L > start) {
            const SeqPair rp = warp_compose_range(pl, start, L, lane);
            const int e = __shfl_sync(0xffffffffu, ebl, start);
            seq_push_run(A, wl, lane, e, rp, base + start);
        }
        if (L == 32) break;
        seq_push_single(A, wl, lane, __shfl_sync(0xffffffffu, tt, L), base + L);
        singles &= singles - 1;
        start = L + 1;
    }
}

// Called by all SEQSUM_THREADS threads of the block.  `sq` = the n (multiple of 32, <= 8192)
// non-negative terms in shared memory.  Returns the sequential float sum to every thread.
__device__ float block_seqsum_exact(const float *__restrict__ sq, int n, SeqSumScratch sc, long long *clk = nullptr) {
#define SEQ_CLK(k) if (clk && threadIdx.x == 0) clk[k] = clock64()
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = SEQSUM_THREADS / 32;   // warps 0..NW-2 compose groups, warp NW-1 walks the literal head
    constexpr int NC = NW - 1;
    const int ng = n >> 5;
    const int hg = max(1, ng >> SEQSUM_HEAD_SHIFT);     // head groups, summed literally
    const int gpw = (ng - hg + NC - 1) / NC;            // contiguous groups per composing warp
    const int ga = min(ng, hg + warp * gpw), gb = (warp < NC) ? min(ng, ga + gpw) : ga;
    SEQ_CLK(0);
    // ---- A: group sums (prediction only, any order); the head warp walks the literal head meanwhile
    if (warp == NC) {
        if (lane == 0) {
            float sHl = 0.0f;
#pragma unroll 8
            for (int i = 0; i < hg * 32; i++) sHl = __fadd_rn(sHl, sq[i]);
            *sc.result = sHl;
            sc.info[2] = 0;
        }
    } else {
#pragma unroll 2
        for (int g = ga; g < gb; g++) {
            const float v = warp_sum_f(sq[g * 32 + lane]);
            if (lane == 0) sc.gs[g] = v;
        }
    }
    __syncthreads();
    SEQ_CLK(1);
    const float sH = *sc.result;
    // ---- B: exclusive float prefix over the groups after the head, by warp 0
    if (warp == 0) {
        const int m = ng - hg;
        const int per = (m + 31) >> 5; // groups per lane, contiguous
        const int g0 = hg + lane * per, g1 = min(ng, g0 + per);
        float loc = 0.0f;
#pragma unroll 1
        for (int g = g0; g < g1; g++) loc += sc.gs[g];
        float inc = loc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            float o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        float run = sH + (inc - loc); // predicted running sum at this lane's first group
#pragma unroll 1
        for (int g = g0; g < g1; g++) {
            float v = sc.gs[g];
            sc.gs[g] = run; // gs[g] := predicted running sum BEFORE group g
            run += v;
        }
        if (lane == 31) sc.gs[ng] = run; // predicted total
    }
    __syncthreads();
    SEQ_CLK(2);
    const bool degenerate = !(sH > 0.0f) || !(sH < INFINITY) || ng <= hg;
    // ---- C: every composing warp folds its contiguous groups, in order, into a short entry list
    SeqEntry *wl = sc.wl + warp * SEQSUM_WL;
    SeqAcc A;
    A.cls = SEQSUM_NONE; A.start = 0; A.n = 0; A.pr.a0 = A.pr.a1 = 0u; A.overflow = false;
    if (!degenerate) {
#pragma unroll 1
        for (int g = ga; g < gb; g++) {
            const float t = sq[g * 32 + lane];
            const float before_g = sc.gs[g], after_g = sc.gs[g + 1];
            const int eb = f32_exponent(before_g);
            SeqPair pr;
            const bool ok = seq_pair(t, eb, pr) && (t < INFINITY);
            if ((f32_exponent(after_g) == eb) && eb >= -90 && __all_sync(0xffffffffu, ok)) {
                SeqPair gp;
                if (__any_sync(0xffffffffu, pr.a0 != pr.a1)) gp = warp_compose_range(pr, 0, 32, lane); // ties: ordered
                else gp.a0 = gp.a1 = min(warp_sum_u(pr.a0), 1u << 26);                                  // plain integer sum
                seq_push_run(A, wl, lane, eb, gp, g * 32);
            } else seq_cross_group(A, wl, lane, t, before_g, g * 32);
        }
        seq_flush(A, wl, lane);
    }
    SEQ_CLK(3);
    if (lane == 0) { sc.wcnt[warp] = A.n; if (A.overflow) sc.info[2] = 1; }
    __syncthreads();
    // compact the per-warp lists, in warp order, into one flat list
    int off = 0, total = 0;
#pragma unroll 1
    for (int w = 0; w < NW; w++) {
        const int c = min(sc.wcnt[w], SEQSUM_WL);
        if (w < warp) off += c;
        total += c;
    }
    if (lane < min(A.n, SEQSUM_WL)) sc.flat[off + lane] = wl[lane];
    __syncthreads();
    SEQ_CLK(4);
    // ---- E: sequential resolution + verification by one thread; state = (exponent, integer mantissa)
    if (tid == 0) {
        float s = sH;
        int fb_from = -1;
        if (degenerate || sc.info[2] != 0) fb_from = hg * 32;
        else {
            unsigned bits = __float_as_uint(s);
            int es = (int)((bits >> 23) & 0xff) - 127;
            unsigned M = (bits & 0x7fffffu) | 0x800000u;
            const int4 *fl = reinterpret_cast<const int4 *>(sc.flat);
            int4 nx = fl[0];
#pragma unroll 1
            for (int q = 0; q < total; q++) {
                const int cls = nx.x;
                const unsigned a0 = (unsigned)nx.y, a1 = (unsigned)nx.z;
                const int start = nx.w;
                if (q + 1 < total) nx = fl[q + 1]; // prefetch: independent of the (es, M) chain
                if (cls == SEQSUM_SINGLE) {
                    s = __uint_as_float(((unsigned)(es + 127) << 23) | (M & 0x7fffffu));
                    s = __fadd_rn(s, __uint_as_float(a1));
                    bits = __float_as_uint(s);
                    es = (int)((bits >> 23) & 0xff) - 127;
                    M = (bits & 0x7fffffu) | 0x800000u;
                    if (es < -100 || es > 126) { fb_from = start + 1; break; } // left the normal range: literal from the next element
                    continue;
                }
                const unsigned add = (M & 1u) ? a1 : a0;
                const unsigned M2 = M + add;
                if (es != cls || add >= (1u << 24) || M2 >= (1u << 24)) { fb_from = start; break; }
                M = M2;
            }
            s = __uint_as_float(((unsigned)(es + 127) << 23) | (M & 0x7fffffu));
        }
        if (fb_from >= 0) {
#pragma unroll 4
            for (int i = fb_from; i < n; i++) s = __fadd_rn(s, sq[i]);
        }
        *sc.result = s;
        sc.info[0] = total;
        sc.info[1] = fb_from;
    }
    SEQ_CLK(5);
    __syncthreads();
    const float r = *sc.result;
    __syncthreads();
    return r;
#undef SEQ_CLK
}
