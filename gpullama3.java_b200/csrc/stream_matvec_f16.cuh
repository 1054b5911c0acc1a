// stream_matvec_f16.cuh -- FP16-weight matvec of the decode step on per-warp bulk-copy rings (FP16 plans, SURVEY 8(a) a4/a6/a9/a10,
// BASELINE configs 1 and 3).  Bit-exact with FP16FloatTensor.vectorDot (FP16FloatTensor.java:62-110) for an L-lane species, like
// k_matvec_f16 (decode_kernels.cuh) which stays as the fallback for odd shapes and the scalar (L = 0) order:
//   * per output row L independent chains  acc_c = fma(w[16j + c], x[16j + c], acc_c), j ascending   (FloatVector.fma: fused),
//   * reduceLanes(ADD) in ascending lane order starting from the identity,
//   * the FP16 -> FP32 widening flushes FP16 subnormals to zero (f16_bits_to_f32_daz, as the reference's vector conversion does).
// A row therefore offers only L-way parallelism of strictly sequential chains, and the arithmetic is trivial (11 elements per clock per
// SM saturate HBM): the kernel is all about keeping bytes in flight.  Layout: the weights stay ROW-MAJOR exactly as in the GGUF file
// (no repack pass for FP16 plans).  Every warp owns a private ring of `stages` shared-memory stages and its own mbarriers: lane 0 issues
// one cp.async.bulk per row segment (seg columns of each of the warp's 32/L rows), the warp waits on the stage's mbarrier, consumes it
// from shared memory, __syncwarp()s and lane 0 refills the stage -- no producer warp, no CTA-wide synchronisation after the activation
// has been staged.  The warp's rows of a stage are `SF_ROW_PAD` bytes apart modulo 128 so its half-warps hit different banks.
//   SF_GATEUP: the warp's row slots are ffn_gate row r and ffn_up row r; SwiGLU (InferenceCore.java:150-158) is applied in the epilogue,
//   so w1, w3 and the SwiGLU kernel of the round-1 FP16 graph collapse into one launch.
// PDL: the rings are filled before griddepcontrol.wait (weights are immutable); x, and out in SF_RESID, are touched only after it.
#pragma once
#include "stream_matvec.cuh"

#define SF_WARPS 8
#define SF_THREADS (SF_WARPS * 32)
#define SF_MAX_STAGES 8
#define SF_ROW_PAD 32

enum { SF_STORE = 0, SF_RESID = 1, SF_GATEUP = 2 };

struct SfArgs {
    const __half *w0, *w1; // [rows][cols] row-major; w1 = ffn_up (SF_GATEUP only)
    const float *x;        // activation, cols floats
    float *out;            // rows floats
    int rows, cols;        // of ONE matrix
    int seg, nseg;         // columns per stage, stages per row
    int stages;            // ring depth per warp
    TraceBuf tr;
};

struct SfLayout {
    int seg, nseg, stages, ctas_per_sm;
    size_t total;
    bool ok;
};

// Host: pick the segment, the ring depth and the CTAs per SM that maximise the bytes in flight per SM.
static inline SfLayout sf_layout(int rows, int cols, int lanes, bool gateup) {
    SfLayout o{};
    o.ok = false;
    if (lanes != 8 && lanes != 16) return o;
    const int rw = 32 / lanes;
    if (cols % 256 || cols < 256) return o;
    if (gateup ? rows % (rw / 2) : rows % rw) return o;
    o.seg = cols % 512 == 0 ? 512 : 256;
    o.nseg = cols / o.seg;
    const size_t stage = (size_t)rw * (o.seg * 2 + SF_ROW_PAD);
    const size_t fixed = (size_t)cols * 4 + SF_WARPS * SF_MAX_STAGES * 8 + 256;
    const size_t per_sm = 220 * 1024;
    int best = 0;
    for (int occ = 1; occ <= 4; occ++) {
        if (per_sm / occ <= fixed + 1024) continue;
        int s = (int)((per_sm / occ - 1024 - fixed) / (SF_WARPS * stage));
        if (s > SF_MAX_STAGES) s = SF_MAX_STAGES;
        if (s < 2) continue;
        const int score = occ * (s - 1);
        if (score >= best) { best = score; o.stages = s; o.ctas_per_sm = occ; }
    }
    if (!best) return o;
    o.total = fixed + (size_t)SF_WARPS * o.stages * stage;
    o.ok = true;
    return o;
}

template <int L, int MODE>
__global__ void __launch_bounds__(SF_THREADS) k_stream_matvec_f16(SfArgs a) {
    constexpr int RW = 32 / L;         // row slots per warp
    constexpr int MR = MODE == SF_GATEUP ? RW / 2 : RW; // matrix rows per warp step
    extern __shared__ __align__(128) unsigned char sf_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *sx = reinterpret_cast<float *>(sf_smem);
    const size_t off_bar = ((size_t)a.cols * 4 + 127) & ~(size_t)127;
    const unsigned row_b = (unsigned)(a.seg * 2 + SF_ROW_PAD), stage_b = RW * row_b;
    const unsigned bar0 = smem_u32(sf_smem + off_bar) + (unsigned)warp * SF_MAX_STAGES * 8;
    unsigned char *ring = sf_smem + off_bar + SF_WARPS * SF_MAX_STAGES * 8 + (size_t)warp * a.stages * stage_b;
    const unsigned ring_u = smem_u32(ring);
    const int S = a.stages, nseg = a.nseg, seg = a.seg;

    // warp-interleaved assignment: group g belongs to global warp (g mod G); consecutive groups go to different SMs
    const int G = gridDim.x * SF_WARPS, gw = warp * gridDim.x + blockIdx.x;
    const int ngroups = a.rows / MR;
    const int my_groups = gw < ngroups ? (ngroups - gw + G - 1) / G : 0;
    const int n_items = my_groups * nseg;
    const size_t row_bytes = (size_t)a.cols * 2;
    const unsigned long long pol = l2_policy_evict_first();

    auto issue = [&](int k) { // lane 0 only
        const int st = k % S, gi = k / nseg, s = k - gi * nseg;
        const size_t row0 = (size_t)(gw + gi * G) * MR;
        const unsigned bar = bar0 + 8u * st;
        mbar_expect_tx(bar, RW * (unsigned)seg * 2u);
#pragma unroll
        for (int r = 0; r < RW; r++) {
            const __half *m = (MODE == SF_GATEUP && r >= MR) ? a.w1 : a.w0;
            const size_t row = row0 + (MODE == SF_GATEUP ? r % MR : r);
            bulk_g2s_evict_first(ring_u + st * stage_b + r * row_b, reinterpret_cast<const unsigned char *>(m) + row * row_bytes + (size_t)s * seg * 2,
                                 (unsigned)seg * 2u, bar, pol);
        }
    };

    if (lane == 0) {
        for (int s = 0; s < S; s++) mbar_init(bar0 + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    trace_entry(a.tr);
    pdl_launch_dependents();
    if (lane == 0)
        for (int k = 0; k < S && k < n_items; k++) issue(k);
    pdl_wait(); // the activation (and out, SF_RESID) come from the previous kernel
    for (int i = tid * 4; i < a.cols; i += SF_THREADS * 4) *reinterpret_cast<float4 *>(sx + i) = *reinterpret_cast<const float4 *>(a.x + i);
    __syncthreads();
    trace_mark(a.tr, 2);

    const int r = lane / L, c = lane % L;
    float acc = 0.0f;
    for (int k = 0; k < n_items; k++) {
        const int st = k % S, gi = k / nseg, s = k - gi * nseg;
        mbar_wait(bar0 + 8u * st, (unsigned)(k / S) & 1u);
        const unsigned short *wr = reinterpret_cast<const unsigned short *>(ring + st * stage_b + r * row_b);
        const float *xs = sx + s * seg;
#pragma unroll 8
        for (int i = c; i < seg; i += L) acc = fmaf(f16_bits_to_f32_daz(wr[i]), xs[i], acc);
        __syncwarp(); // every lane has consumed the stage: lane 0 may overwrite it
        if (lane == 0 && k + S < n_items) issue(k + S);
        if (s == nseg - 1) {
            float result = 0.0f; // reduceLanes(ADD): ascending lanes from the identity
#pragma unroll
            for (int q = 0; q < L; q++) result = __fadd_rn(result, __shfl_sync(0xffffffffu, acc, r * L + q));
            const size_t row = (size_t)(gw + gi * G) * MR + (MODE == SF_GATEUP ? r % MR : r);
            if (MODE == SF_GATEUP) {
                const float up = __shfl_sync(0xffffffffu, result, ((r + MR) % RW) * L);
                if (r < MR && c == 0) a.out[row] = swiglu(result, up);
            } else if (c == 0) {
                a.out[row] = MODE == SF_RESID ? __fadd_rn(a.out[row], result) : result; // x[i] = x[i] + xb2[i] (InferenceCore.java:143,164)
            }
            acc = 0.0f;
        }
    }
    trace_mark(a.tr, 3);
}
