// seqsum.cuh -- exact, parallel evaluation of a SEQUENTIAL float32 sum of non-negative terms.
//
// The reference's RMSNorm accumulates  ss = ((0 + x0*x0) + x1*x1) + ...  one float add at a
// time (InferenceCore.rmsnorm, InferenceCore.java:39-48 via FloatTensor.reduce,
// FloatTensor.java:110-116).  Float addition is not associative, so a tree reduction gives
// different bits; a literal chain is dim x 4 cycles (~9-12 us for dim 4096), which would be a
// third of the decode budget.  This file reproduces the chain's result bit for bit in ~1 us:
//
//   While the running sum stays inside one binade [2^e, 2^(e+1)) its mantissa M is an integer
//   in units of u = 2^(e-23) and one step is  M <- M + k + r  with  t/u = k + f,
//   r = [f > 1/2], or on a tie (f == 1/2) r = (M + k) & 1 (round half to even).  So a step is a
//   function  M -> M + a[M & 1]  described by two integers (a0, a1), and such functions compose
//   associatively:  (a ; b)[p] = a[p] + b[(p + a[p]) & 1].  A segmented parallel scan composes
//   all steps of a binade at once.
//   Which binade each step starts in is PREDICTED from a double-precision prefix sum; a step
//   that crosses into a higher binade is executed as a real float add.  One thread then walks
//   the ~10-20 resulting (segment | crossing step) entries and VERIFIES every prediction
//   (exponent at segment entry, mantissa < 2^24 at exit).  Any failed check falls back to the
//   literal sequential loop from that element on, so the result is exact unconditionally;
//   mispredictions (sum within ~1e-6 of a power of two) only cost time.
#pragma once
#include "common.cuh"

#define SEQSUM_HEAD 32      // first elements summed literally (the sum crosses binades quickly at first)
#define SEQSUM_MAXSEG 192   // capacity of the entry list; more -> sequential fallback
#define SEQSUM_CROSS 0x7fff

struct SeqPair {
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

struct SeqSeg { // one entry of the resolution list
    int cls;    // SEQSUM_CROSS or the (unbiased) binade exponent e of the segment
    int start;  // index of the first element of the entry (for the fallback)
    unsigned a0, a1;
    float t;    // the term of a crossing step
};

struct SeqSumScratch {
    short *cls;      // [n] per-element class
    SeqSeg *list;    // [SEQSUM_MAXSEG]
    double *wd;      // [32]
    unsigned *wi;    // [32*4]
    int *nseg;       // [1]
    float *result;   // [1]
};

__host__ __device__ inline size_t seqsum_scratch_bytes(int n) {
    return (size_t)((n * 2 + 15) & ~15) + SEQSUM_MAXSEG * sizeof(SeqSeg) + 32 * 8 + 32 * 4 * 4 + 16;
}

__device__ __forceinline__ SeqSumScratch seqsum_carve(unsigned char *base, int n) {
    SeqSumScratch s;
    unsigned char *p = base;
    s.wd = reinterpret_cast<double *>(p); p += 32 * 8;
    s.list = reinterpret_cast<SeqSeg *>(p); p += SEQSUM_MAXSEG * sizeof(SeqSeg);
    s.wi = reinterpret_cast<unsigned *>(p); p += 32 * 4 * 4;
    s.nseg = reinterpret_cast<int *>(p); p += 8;
    s.result = reinterpret_cast<float *>(p); p += 8;
    s.cls = reinterpret_cast<short *>(p);
    return s;
}

__device__ __forceinline__ int dbl_exponent(double d) { return ((__double2hiint(d) >> 20) & 0x7ff) - 1023; }

// All threads of the block must call this (blockDim.x multiple of 32, <= 1024).  `sq` holds the n
// non-negative terms in shared memory.  Returns the sequential float sum to every thread.
// E = max elements per thread (n <= E * blockDim.x).
template <int E>
__device__ float block_seqsum_exact(const float *__restrict__ sq, int n, SeqSumScratch sc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int H = n < SEQSUM_HEAD ? n : SEQSUM_HEAD;
    // ---- head: literal chain, computed redundantly by every thread (uniform result)
    float sH = 0.0f;
    for (int i = 0; i < H; i++) sH = __fadd_rn(sH, sq[i]);
    if (H == n) return sH;
    const bool degenerate = !(sH > 0.0f) || !(sH < INFINITY);
    if (degenerate) { // all-zero head or overflow: literal chain (uniform branch)
        if (tid == 0) {
            float s = sH;
            for (int i = H; i < n; i++) s = __fadd_rn(s, sq[i]);
            *sc.result = s;
        }
        __syncthreads();
        float r = *sc.result;
        __syncthreads();
        return r;
    }
    const int per = (n + blockDim.x - 1) / blockDim.x; // contiguous chunk per thread (<= E)
    const int i0 = tid * per;
    // ---- double-precision prefix sums (prediction only)
    float t[E];
    double loc = 0.0;
#pragma unroll
    for (int j = 0; j < E; j++) {
        int i = i0 + j;
        t[j] = (j < per && i >= H && i < n) ? sq[i] : 0.0f;
        loc += (double)t[j];
    }
    double inc = loc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) sc.wd[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        double w = lane < nwarps ? sc.wd[lane] : 0.0;
        double wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            double o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += o;
        }
        sc.wd[lane] = wi - w; // exclusive
    }
    __syncthreads();
    double before = (double)sH + sc.wd[warp] + (inc - loc);
    // ---- classify every step, build its (a0,a1)
    SeqPair pr[E];
    short cl[E];
#pragma unroll
    for (int j = 0; j < E; j++) {
        int i = i0 + j;
        pr[j].a0 = pr[j].a1 = 0;
        cl[j] = 0;
        if (j < per && i >= H && i < n) {
            double after = before + (double)t[j];
            int eb = dbl_exponent(before), ea = dbl_exponent(after);
            if (ea > eb || eb < -120 || !(t[j] < INFINITY)) {
                cl[j] = SEQSUM_CROSS;
            } else {
                double scale = __hiloint2double((1023 + 23 - eb) << 20, 0); // 2^(23-eb)
                double scaled = (double)t[j] * scale;                          // exact, < 2^24
                double kf = floor(scaled);
                unsigned k = (unsigned)kf;
                double f = scaled - kf;
                unsigned up = f > 0.5 ? 1u : 0u;
                bool tie = f == 0.5;
                pr[j].a0 = k + (tie ? (k & 1u) : up);
                pr[j].a1 = k + (tie ? ((k + 1u) & 1u) : up);
                cl[j] = (short)eb;
            }
            sc.cls[i] = cl[j];
            before = after;
        }
    }
    __syncthreads();
    // ---- head flags, per-thread aggregate since the last head
    unsigned headmask = 0;
    SeqPair agg = {0u, 0u};
    int nheads = 0;
#pragma unroll
    for (int j = 0; j < E; j++) {
        int i = i0 + j;
        if (j < per && i >= H && i < n) {
            bool head = (i == H) || cl[j] == SEQSUM_CROSS || sc.cls[i - 1] != cl[j];
            if (head) { headmask |= 1u << j; nheads++; agg = pr[j]; }
            else agg = seq_compose(agg, pr[j]);
        }
    }
    // ---- block inclusive scans: segmented composition (h, p) and head counts
    unsigned h = nheads > 0 ? 1u : 0u;
    SeqPair p = agg;
    int cnt = nheads;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned oh = __shfl_up_sync(0xffffffffu, h, d);
        unsigned o0 = __shfl_up_sync(0xffffffffu, p.a0, d), o1 = __shfl_up_sync(0xffffffffu, p.a1, d);
        int oc = __shfl_up_sync(0xffffffffu, cnt, d);
        if (lane >= d) {
            if (!h) { SeqPair L = {o0, o1}; p = seq_compose(L, p); h = oh; }
            cnt += oc;
        }
    }
    if (lane == 31) { sc.wi[warp * 4 + 0] = h; sc.wi[warp * 4 + 1] = p.a0; sc.wi[warp * 4 + 2] = p.a1; sc.wi[warp * 4 + 3] = (unsigned)cnt; }
    // exclusive value inside the warp
    unsigned xh = __shfl_up_sync(0xffffffffu, h, 1);
    unsigned x0 = __shfl_up_sync(0xffffffffu, p.a0, 1), x1 = __shfl_up_sync(0xffffffffu, p.a1, 1);
    int xc = __shfl_up_sync(0xffffffffu, cnt, 1);
    if (lane == 0) { xh = 0; x0 = 0; x1 = 0; xc = 0; }
    __syncthreads();
    if (warp == 0) {
        unsigned wh = lane < nwarps ? sc.wi[lane * 4 + 0] : 0u;
        SeqPair wp = {lane < nwarps ? sc.wi[lane * 4 + 1] : 0u, lane < nwarps ? sc.wi[lane * 4 + 2] : 0u};
        int wc = lane < nwarps ? (int)sc.wi[lane * 4 + 3] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned oh = __shfl_up_sync(0xffffffffu, wh, d);
            unsigned o0 = __shfl_up_sync(0xffffffffu, wp.a0, d), o1 = __shfl_up_sync(0xffffffffu, wp.a1, d);
            int oc = __shfl_up_sync(0xffffffffu, wc, d);
            if (lane >= d) {
                if (!wh) { SeqPair L = {o0, o1}; wp = seq_compose(L, wp); wh = oh; }
                wc += oc;
            }
        }
        __syncwarp();
        sc.wi[lane * 4 + 0] = wh; sc.wi[lane * 4 + 1] = wp.a0; sc.wi[lane * 4 + 2] = wp.a1; sc.wi[lane * 4 + 3] = (unsigned)wc;
        if (lane == 31) *sc.nseg = wc;
    }
    __syncthreads();
    if (warp > 0) { // fold in the inclusive aggregate of the preceding warps
        unsigned ch = sc.wi[(warp - 1) * 4 + 0];
        SeqPair cp = {sc.wi[(warp - 1) * 4 + 1], sc.wi[(warp - 1) * 4 + 2]};
        int cc = (int)sc.wi[(warp - 1) * 4 + 3];
        if (!xh) { SeqPair X = {x0, x1}; X = seq_compose(cp, X); x0 = X.a0; x1 = X.a1; xh = ch; }
        xc += cc;
    }
    const int nseg = *sc.nseg;
    // ---- emit one list entry per segment (at the element where the segment ends)
    if (nseg <= SEQSUM_MAXSEG) {
        SeqPair cur = {x0, x1};
        int rank = xc - 1; // index of the segment the carry-in belongs to
        int seg_start = -1;
#pragma unroll
        for (int j = 0; j < E; j++) {
            int i = i0 + j;
            if (j < per && i >= H && i < n) {
                if (headmask & (1u << j)) { cur = pr[j]; rank++; seg_start = i; }
                else cur = seq_compose(cur, pr[j]);
                bool last = (i == n - 1);
                bool next_head = !last && (sc.cls[i + 1] == SEQSUM_CROSS || sc.cls[i + 1] != cl[j]);
                if (last || next_head) {
                    SeqSeg e;
                    e.cls = cl[j];
                    e.a0 = cur.a0; e.a1 = cur.a1;
                    e.t = t[j];
                    e.start = seg_start; // -1 when the segment began in an earlier thread: patched below
                    sc.list[rank].cls = e.cls; sc.list[rank].a0 = e.a0; sc.list[rank].a1 = e.a1; sc.list[rank].t = e.t;
                }
                if (headmask & (1u << j)) sc.list[rank].start = i;
            }
        }
    }
    __syncthreads();
    // ---- sequential resolution + verification by one thread
    if (tid == 0) {
        float s = sH;
        int fb_from = -1;
        if (nseg > SEQSUM_MAXSEG) fb_from = H;
        else {
            for (int g = 0; g < nseg; g++) {
                const SeqSeg e = sc.list[g];
                if (e.cls == SEQSUM_CROSS) { s = __fadd_rn(s, e.t); continue; }
                unsigned bits = __float_as_uint(s);
                int es = (int)((bits >> 23) & 0xff) - 127;
                unsigned M = (bits & 0x7fffffu) | 0x800000u;
                unsigned M2 = M + ((M & 1u) ? e.a1 : e.a0);
                if (es != e.cls || M2 >= (1u << 24)) { fb_from = e.start; break; }
                s = __uint_as_float((bits & 0xff800000u) | (M2 & 0x7fffffu));
            }
        }
        if (fb_from >= 0)
            for (int i = fb_from; i < n; i++) s = __fadd_rn(s, sq[i]);
        *sc.result = s;
    }
    __syncthreads();
    float r = *sc.result;
    __syncthreads();
    return r;
}
