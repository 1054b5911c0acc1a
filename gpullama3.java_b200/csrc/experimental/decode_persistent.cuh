// decode_persistent.cuh -- ROUND-2 CANDIDATE: ONE persistent kernel per decoded token instead of
// 227 dependent launches.  Compile-checked only (opt-in build: B200_NVCC_DEFINES="B200_SEQSUM_V2 B200_PERSISTENT_DECODE"); the
// arithmetic of every phase is a copy of the validated round-1 kernels (k_stream_matvec_q8, k_rmsnorm_quant, k_attention,
// k_argmax_advance), what is NEW and untested is the orchestration:
//
//   * grid = one CTA per SM, 8 consumer warps + 1 producer warp, all resident for the whole token;
//   * the producer thread walks the tile-major weight stream of EVERY matrix of the token in consumption order (QKV, Wo,
//     gate/up, W2 per layer, then lm_head) through one shared-memory ring: weight addresses never depend on activations, so
//     HBM keeps streaming across what used to be kernel boundaries (the round-1 graph drained and refilled a 96 KB ring 225
//     times per token);
//   * phases are separated by device-scope epoch counters (monotone, never reset: target = (tick * layers + layer + 1) *
//     arrivers -- the scheme of the tensor-parallel flags, common.cuh) instead of kernel boundaries: 5 per layer
//     (QKV rows complete -> attention; attention heads complete -> Wo; x complete -> norm; hidden activation complete -> W2;
//     x complete -> next layer);
//   * RMSNorm is computed REDUNDANTLY by every CTA straight into its own shared-memory activation buffer (exact accumulator:
//     seqsum2.cuh with 256 threads), which removes two of the seven dependencies of a layer and the xq/xs round trip;
//   * attention heads run on the first n_heads CTAs with the consumer warps (same code as k_attention);
//   * the embedding row is read in place: layer 0's norm squares the embedding row and layer 0's Wo epilogue writes
//     x = emb + Wo*att, so no "x = embedding" pass and no grid sync before the first layer.
// Restrictions of this draft: Q8_0 streaming path, single GPU, every matrix with the same tile size (true for the Llama /
// Qwen3 shapes: segments of 2048 columns).
#pragma once
#include "../decode_kernels.cuh"
#include "../stream_matvec.cuh"
#include "seqsum2.cuh"

#define PD_CT (SMV_CONSUMER_WARPS * 32) // consumer threads
#define PD_MAX_STAGES 24

enum { PD_S_QKV = 0, PD_S_ATT = 1, PD_S_WO = 2, PD_S_GU = 3, PD_S_W2 = 4, PD_S_LM = 5, PD_S_TICK = 8, PD_S_LMTICK = 9, PD_S_WORDS = 16 };
// PD_S_TICK counts launches (epoch of the per-layer counters); PD_S_LMTICK counts launches that ran the lm_head (the prefill
// graph does not), which is the epoch of PD_S_LM.

struct PdLayer {
    TileMat qkv, wo, gu, w2;
    const float *attn_norm, *ffn_norm, *q_norm, *k_norm;
    float *kc, *vc; // this layer's FP32 KV cache
};

struct PdArgs {
    const PdLayer *layers; // device array [n_layers]
    int n_layers;
    TileMat lm_head;
    const float *out_norm;
    DevMat emb;
    int dim, hidden, qd, kvd, n_heads, n_kv_heads, head_size, arch, vocab, ctx;
    float eps, sqrt_hs;
    const float *rope_cr, *rope_ci;
    StepState *st;
    const int *seq_tokens;
    int *out_ids;
    float *x, *qkv, *hb, *logits;
    int8_t *attq;
    float *atts;
    int8_t *hq;
    float *hs;
    unsigned *blk_cnt;
    float *part_val;
    int *part_idx;
    unsigned *sync; // [PD_S_WORDS] epoch counters + tick
    int with_logits;
};

struct PdSmem {
    size_t off_bar, off_xq, off_xs, off_nbuf, off_xbuf, off_seq, off_terms, off_hvals, off_misc, off_ring, total;
    int stages, stage_bytes, nbs_pad, nbuf_floats;
};

__host__ __device__ inline PdSmem pd_layout(int dim, int qd, int hidden, int head_size, int ctx, int seg, size_t budget) {
    PdSmem L;
    const int unit = smv_unit_bytes(seg);
    L.stage_bytes = (4 * unit + 127) & ~127;
    L.nbs_pad = (seg / 32) | 1;
    int maxc = dim > qd ? dim : qd;
    if (hidden > maxc) maxc = hidden;
    const int dim_pad = (dim + PD_CT - 1) / PD_CT * PD_CT;
    const int att_floats = 3 * head_size + ctx;
    L.nbuf_floats = dim_pad > att_floats ? dim_pad : att_floats;
    size_t o = 0;
    L.off_bar = o; o += 2 * PD_MAX_STAGES * 8 + PD_MAX_STAGES * 4;
    o = (o + 15) & ~(size_t)15;
    L.off_xq = o; o += (size_t)maxc;
    L.off_xs = o; o += (size_t)(maxc / 32) * 4;
    o = (o + 15) & ~(size_t)15;
    L.off_nbuf = o; o += (size_t)L.nbuf_floats * 4; // squares of the norm | q,k,out,att of the attention (time-disjoint)
    o = (o + 15) & ~(size_t)15;
    L.off_xbuf = o; o += (size_t)dim * 4; // the residual stream, kept between the two passes of the norm
    o = (o + 15) & ~(size_t)15;
    L.off_seq = o; o += seqsum2_scratch_bytes();
    o = (o + 15) & ~(size_t)15;
    L.off_terms = o; o += (size_t)SMV_CONSUMER_WARPS * 4 * L.nbs_pad * 4;
    L.off_hvals = o; o += SMV_HVALS * 4;
    L.off_misc = o; o += 64 * 4; // red[8], s_val[2], scale, argmax merge scratch
    o = (o + 127) & ~(size_t)127;
    L.off_ring = o;
    long room = (long)budget - (long)o;
    int s = room > 0 ? (int)(room / L.stage_bytes) : 0;
    if (s > PD_MAX_STAGES) s = PD_MAX_STAGES;
    L.stages = s;
    L.total = o + (size_t)s * L.stage_bytes;
    return L;
}

__device__ __forceinline__ unsigned pd_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
struct PdConsumerSync {
    __device__ __forceinline__ void operator()() const { consumer_bar_sync(); }
};
// every consumer thread calls these; the CTA's stores (made before the barrier) are published by thread 0's fence + atomic
__device__ __forceinline__ void pd_arrive(unsigned *cnt, int tid) {
    consumer_bar_sync();
    if (tid == 0) { __threadfence(); atomicAdd(cnt, 1u); }
}
__device__ __forceinline__ void pd_wait(const unsigned *cnt, unsigned target, int tid) {
    if (tid == 0)
        while ((int)(pd_ld_acquire(cnt) - target) < 0) {}
    consumer_bar_sync();
}

// ---- producer: the whole token's weight stream, in consumption order ----------------------------------------------------
__device__ __forceinline__ void pd_produce_matrix(const TileMat &W, unsigned char *smem, const PdSmem &L, unsigned bar0, unsigned &seq) {
    const int S = L.stages, ngroups = W.rows >> 2;
    const int g0 = (int)(((long long)blockIdx.x * ngroups) / gridDim.x), g1 = (int)(((long long)(blockIdx.x + 1) * ngroups) / gridDim.x);
    const unsigned tile_bytes = 4u * (unsigned)W.unit_bytes;
    for (int gb = g0; gb < g1; gb += SMV_CONSUMER_WARPS) {
        const int nw = min(SMV_CONSUMER_WARPS, g1 - gb);
        for (int s = 0; s < W.nseg; s++)
            for (int w = 0; w < nw; w++, seq++) {
                const int st = seq % S;
                const unsigned ph = (seq / S) & 1u;
                mbar_wait(bar0 + 8 * (PD_MAX_STAGES + st), ph ^ 1u);
                const unsigned full = bar0 + 8 * st;
                mbar_expect_tx(full, tile_bytes);
                bulk_g2s(smem_u32(smem + L.off_ring + (size_t)st * L.stage_bytes), W.base + ((size_t)(gb + w) * W.nseg + s) * tile_bytes, tile_bytes, full);
            }
    }
}

// ---- consumers: one matrix (the loop of k_stream_matvec_q8, activation already in shared memory) --------------------------
// l0_emb: layer 0's Wo writes x = embedding + acc (the embedding row is never copied into x beforehand).
template <int MODE>
__device__ __forceinline__ void pd_consume_matrix(const TileMat &W, const PdArgs &a, unsigned char *smem, const PdSmem &L, unsigned bar0, volatile unsigned *rel,
                                                  unsigned &seq_base, float *out, bool argmax, bool l0_emb, int token, int tid) {
    const int lane = tid & 31, warp = tid >> 5, S = L.stages;
    const int ngroups = W.rows >> 2;
    const int g0 = (int)(((long long)blockIdx.x * ngroups) / gridDim.x), g1 = (int)(((long long)(blockIdx.x + 1) * ngroups) / gridDim.x);
    const int nseg = W.nseg, nbs = W.seg >> 5;
    float *terms = reinterpret_cast<float *>(smem + L.off_terms) + (size_t)warp * 4 * L.nbs_pad;
    const unsigned char *sact = smem + L.off_xq;
    const float *sxs = reinterpret_cast<const float *>(smem + L.off_xs);
    float *hvals = reinterpret_cast<float *>(smem + L.off_hvals);
    const int hsel = (lane >> 2) & 1;
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    for (int gb = g0; gb < g1; gb += SMV_CONSUMER_WARPS) {
        const int nw = min(SMV_CONSUMER_WARPS, g1 - gb);
        if (warp < nw) {
            const int G = gb + warp;
            float acc = 0.0f;
            for (int s = 0; s < nseg; s++) {
                const unsigned seq = seq_base + (unsigned)(s * nw + warp);
                const int st = seq % S;
                const unsigned lap = seq / S;
                if (lane == 0)
                    while (rel[st] != lap) {}
                __syncwarp();
                mbar_wait(bar0 + 8 * st, lap & 1u);
                const unsigned char *tile = smem + L.off_ring + (size_t)st * L.stage_bytes;
                for (int b = lane; b < nbs; b += 32) {
                    const unsigned char *ab = sact + ((size_t)(s * nbs + b) << 5);
                    const int4 a0 = *reinterpret_cast<const int4 *>(ab + 16 * hsel);
                    const int4 a1 = *reinterpret_cast<const int4 *>(ab + 16 * (hsel ^ 1));
                    const float as = sxs[s * nbs + b];
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const unsigned char *wb = tile + (size_t)r * W.unit_bytes + ((size_t)b << 5);
                        const int4 w0 = *reinterpret_cast<const int4 *>(wb + 16 * hsel);
                        const int4 w1 = *reinterpret_cast<const int4 *>(wb + 16 * (hsel ^ 1));
                        const __half sc = *reinterpret_cast<const __half *>(tile + (size_t)r * W.unit_bytes + W.seg + 2 * b);
                        int isum = __dp4a(w0.x, a0.x, 0);
                        isum = __dp4a(w0.y, a0.y, isum);
                        isum = __dp4a(w0.z, a0.z, isum);
                        isum = __dp4a(w0.w, a0.w, isum);
                        isum = __dp4a(w1.x, a1.x, isum);
                        isum = __dp4a(w1.y, a1.y, isum);
                        isum = __dp4a(w1.z, a1.z, isum);
                        isum = __dp4a(w1.w, a1.w, isum);
                        terms[r * L.nbs_pad + b] = __fmul_rn((float)isum, __fmul_rn(__half2float(sc), as));
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    rel[st] = lap + 1u;
                    mbar_arrive(bar0 + 8 * (PD_MAX_STAGES + st));
                }
                if (lane < 4) {
                    const float *t = terms + lane * L.nbs_pad;
                    for (int b = 0; b < nbs; b++) acc = __fadd_rn(acc, t[b]);
                }
                __syncwarp();
            }
            if (MODE == SMV_GATEUP) {
                const float up = __shfl_down_sync(0xffffffffu, acc, 2);
                if (lane < 2) {
                    const int unit = 2 * G + lane;
                    const float hval = swiglu_exact(acc, up);
                    out[unit] = hval;
                    hvals[unit - 2 * g0] = hval;
                }
            } else if (lane < 4) {
                const size_t row = (size_t)4 * G + lane;
                if (MODE == SMV_RESID) {
                    const float base = l0_emb ? emb_get(a.emb, token, (int)row) : out[row];
                    out[row] = __fadd_rn(base, acc);
                } else {
                    out[row] = acc;
                    if (acc > best) { best = acc; best_i = (int)row; }
                }
            }
        }
        seq_base += (unsigned)(nseg * nw);
    }
    if (MODE == SMV_GATEUP) { // Q8_0 quantisation of the hidden activation: copy of k_stream_matvec_q8's epilogue
        consumer_bar_sync();
        const int u0 = 2 * g0, u1 = 2 * g1;
        if (u1 > u0) {
            for (int blk = (u0 >> 5) + warp; blk <= ((u1 - 1) >> 5); blk += SMV_CONSUMER_WARPS) {
                const int lo = max(blk << 5, u0), hi = min((blk << 5) + 32, u1);
                float v = 0.0f;
                bool mine = true;
                if (hi - lo == 32) v = hvals[(blk << 5) + lane - u0];
                else {
                    unsigned old = 0;
                    if (lane == 0) {
                        __threadfence();
                        old = atomicAdd(&a.blk_cnt[blk], (unsigned)(hi - lo));
                    }
                    old = __shfl_sync(0xffffffffu, old, 0);
                    mine = (old + (unsigned)(hi - lo) == 32u);
                    if (mine) {
                        __threadfence();
                        v = ldcg_f32(out + (blk << 5) + lane);
                        if (lane == 0) a.blk_cnt[blk] = 0u;
                    }
                }
                if (mine) {
                    float as;
                    const int q = quant_block_lane(v, as);
                    a.hq[(blk << 5) + lane] = (int8_t)q;
                    if (lane == 0) a.hs[blk] = as;
                }
            }
        }
    } else if (MODE == SMV_STORE && argmax) {
        int *cand_i = reinterpret_cast<int *>(hvals + 64);
        consumer_bar_sync(); // hvals may still be read by a previous phase
        if (lane < 4) { hvals[warp * 4 + lane] = best; cand_i[warp * 4 + lane] = best_i; }
        consumer_bar_sync();
        if (tid == 0) {
            float bv = -INFINITY;
            int bi = 0x7fffffff;
            for (int k = 0; k < SMV_CONSUMER_WARPS * 4; k++) {
                const float v = hvals[k];
                const int ix = cand_i[k];
                if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
            }
            a.part_val[blockIdx.x] = bv;
            a.part_idx[blockIdx.x] = bi;
        }
    }
}

// ---- RMSNorm of the residual stream into THIS CTA's activation buffer (arithmetic of k_rmsnorm_quant) ----------------------
__device__ __forceinline__ float4 pd_ldcg128(const float *p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void pd_norm_to_smem(const PdArgs &a, const float *w, bool from_emb, int token, unsigned char *smem, const PdSmem &L, int tid) {
    const int lane = tid & 31, warp = tid >> 5, dim = a.dim;
    float *sq = reinterpret_cast<float *>(smem + L.off_nbuf);
    float *xb = reinterpret_cast<float *>(smem + L.off_xbuf);
    float *misc = reinterpret_cast<float *>(smem + L.off_misc);
    SeqSum2Scratch scratch = seqsum2_carve(smem + L.off_seq);
    const int dim_pad = (dim + PD_CT - 1) / PD_CT * PD_CT;
    if (from_emb) { // first layer: the embedding row (quantised table: element-wise)
        for (int i = tid; i < dim; i += PD_CT) {
            const float v = emb_get(a.emb, token, i);
            xb[i] = v;
            sq[i] = __fmul_rn(v, v);
        }
    } else { // x from L2: four independent 16-byte loads in flight per thread, one round trip for dim <= 4096
        const int n4 = dim >> 2;
        for (int base = 0; base < n4; base += 4 * PD_CT) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i4 = base + u * PD_CT + tid;
                v[u] = i4 < n4 ? pd_ldcg128(a.x + 4 * i4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i4 = base + u * PD_CT + tid;
                if (i4 < n4) {
                    reinterpret_cast<float4 *>(xb)[i4] = v[u];
                    reinterpret_cast<float4 *>(sq)[i4] = make_float4(__fmul_rn(v[u].x, v[u].x), __fmul_rn(v[u].y, v[u].y), __fmul_rn(v[u].z, v[u].z), __fmul_rn(v[u].w, v[u].w));
                }
            }
        }
    }
    for (int i = dim + tid; i < dim_pad; i += PD_CT) sq[i] = 0.0f;
    consumer_bar_sync();
    float ss = block_seqsum_exact_v2_t<PD_CT>(sq, dim, scratch, tid, PdConsumerSync());
    if (tid == 0) {
        ss = __fdiv_rn(ss, (float)dim);
        ss = __fadd_rn(ss, a.eps);
        misc[16] = (float)(1.0 / sqrt((double)ss));
    }
    consumer_bar_sync();
    ss = misc[16];
    int8_t *sxq = reinterpret_cast<int8_t *>(smem + L.off_xq);
    float *sxs = reinterpret_cast<float *>(smem + L.off_xs);
    const int nb = dim / 32;
#pragma unroll 1
    for (int b0 = warp; b0 < nb; b0 += 8 * SMV_CONSUMER_WARPS) { // 8 norm-weight loads in flight per lane; x comes from shared memory
        float wv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int b = b0 + u * SMV_CONSUMER_WARPS;
            wv[u] = b < nb ? __ldg(w + b * 32 + lane) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int b = b0 + u * SMV_CONSUMER_WARPS;
            if (b < nb) {
                const float v = __fmul_rn(wv[u], __fmul_rn(ss, xb[b * 32 + lane]));
                float as;
                const int q = quant_block_lane(v, as);
                sxq[b * 32 + lane] = (int8_t)q;
                if (lane == 0) sxs[b] = as;
            }
        }
    }
    consumer_bar_sync();
}

// a quantised activation vector produced by other CTAs (attention output, hidden activation) -> shared memory
__device__ __forceinline__ void pd_load_act(const int8_t *q, const float *s, int cols, unsigned char *smem, const PdSmem &L, int tid) {
    int4 *sxq = reinterpret_cast<int4 *>(smem + L.off_xq);
    float *sxs = reinterpret_cast<float *>(smem + L.off_xs);
    const int4 *src = reinterpret_cast<const int4 *>(q);
    for (int c = tid; c < cols / 16; c += PD_CT) sxq[c] = __ldcg(src + c);
    for (int b = tid; b < cols / 32; b += PD_CT) sxs[b] = __ldcg(s + b);
    consumer_bar_sync();
}

// ---- one attention head with the consumer warps: copy of k_attention's body (single GPU, Q8_0 output) -----------------------
template <int HS>
__device__ __forceinline__ void pd_attention_head(const PdArgs &a, const PdLayer &Ly, int h, int pos, unsigned char *smem, const PdSmem &L, int tid) {
    float *sm = reinterpret_cast<float *>(smem + L.off_nbuf);
    float *misc = reinterpret_cast<float *>(smem + L.off_misc);
    float *red = misc, *s_val = misc + 8;
    float *sq = sm, *sk = sm + HS, *so = sm + 2 * HS, *att = sm + 3 * HS;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int HALF = HS / 2;
    const int nt = pos + 1;
    const int kv_mul = a.n_heads / a.n_kv_heads, kvh = h / kv_mul;
    const int qd = a.n_heads * HS, kvd = a.n_kv_heads * HS;
    float *qkv = a.qkv, *kc = Ly.kc, *vc = Ly.vc;
    const float *qsrc = qkv + h * HS, *ksrc = qkv + qd + kvh * HS, *vsrc = qkv + qd + kvd + kvh * HS;
    if (tid < HS) {
        const bool is_q = tid < HALF;
        const int p = is_q ? tid : tid - HALF;
        const float *src = is_q ? qsrc : ksrc;
        int i0, i1;
        if (a.arch == 1) { i0 = p; i1 = p + HALF; } else { i0 = 2 * p; i1 = 2 * p + 1; }
        float v0 = ldcg_f32c(src + i0), v1 = ldcg_f32c(src + i1); // written by other CTAs in this kernel: bypass L1
        if (a.arch == 1) {
            float *sqr = is_q ? so : sk;
            sqr[i0] = __fmul_rn(v0, v0);
            sqr[i1] = __fmul_rn(v1, v1);
            asm volatile("bar.sync 3, %0;" ::"n"(HS) : "memory");
            if (p == 0) {
                float ss = 0.0f;
                for (int i = 0; i < HS; i++) ss = __fadd_rn(ss, sqr[i]);
                ss = __fdiv_rn(ss, (float)HS);
                ss = __fadd_rn(ss, a.eps);
                s_val[is_q ? 0 : 1] = (float)(1.0 / sqrt((double)ss));
            }
            asm volatile("bar.sync 3, %0;" ::"n"(HS) : "memory");
            const float ss = s_val[is_q ? 0 : 1];
            const float *nw = is_q ? Ly.q_norm : Ly.k_norm;
            v0 = __fmul_rn(nw[i0], __fmul_rn(ss, v0));
            v1 = __fmul_rn(nw[i1], __fmul_rn(ss, v1));
            asm volatile("bar.sync 3, %0;" ::"n"(HS) : "memory");
        }
        const float fcr = a.rope_cr[(size_t)pos * HALF + p], fci = a.rope_ci[(size_t)pos * HALF + p];
        const float r0 = __fsub_rn(__fmul_rn(v0, fcr), __fmul_rn(v1, fci));
        const float r1 = __fadd_rn(__fmul_rn(v0, fci), __fmul_rn(v1, fcr));
        float *dst = is_q ? sq : sk;
        dst[i0] = r0;
        dst[i1] = r1;
        if (h % kv_mul == 0 && !is_q) {
            const size_t o = (size_t)pos * kvd + kvh * HS;
            kc[o + i0] = r0;
            kc[o + i1] = r1;
            vc[o + i0] = ldcg_f32c(vsrc + i0);
            vc[o + i1] = ldcg_f32c(vsrc + i1);
        }
    }
    consumer_bar_sync();
    float lmax = -INFINITY;
    for (int t = tid; t < nt; t += PD_CT) {
        float acc = 0.0f;
        if (t == pos) {
#pragma unroll 8
            for (int j = 0; j < HS; j++) acc = __fadd_rn(acc, __fmul_rn(sq[j], sk[j]));
        } else {
            const float4 *k = reinterpret_cast<const float4 *>(kc + (size_t)t * kvd + kvh * HS);
#pragma unroll
            for (int j0 = 0; j0 < HS / 4; j0 += 8) {
                float4 kk[8];
#pragma unroll
                for (int u = 0; u < 8; u++) kk[u] = __ldcg(k + j0 + u); // rows of earlier tokens: written by earlier launches
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int j = 4 * (j0 + u);
                    acc = __fadd_rn(acc, __fmul_rn(sq[j + 0], kk[u].x));
                    acc = __fadd_rn(acc, __fmul_rn(sq[j + 1], kk[u].y));
                    acc = __fadd_rn(acc, __fmul_rn(sq[j + 2], kk[u].z));
                    acc = __fadd_rn(acc, __fmul_rn(sq[j + 3], kk[u].w));
                }
            }
        }
        const float s = __fdiv_rn(acc, a.sqrt_hs);
        att[t] = s;
        lmax = fmaxf(lmax, s);
    }
    lmax = warp_max_f(lmax);
    if (lane == 0) red[warp] = lmax;
    consumer_bar_sync();
    float mx = red[0];
#pragma unroll
    for (int w = 1; w < SMV_CONSUMER_WARPS; w++) mx = fmaxf(mx, red[w]);
    for (int t = tid; t < nt; t += PD_CT) att[t] = (float)exp((double)__fsub_rn(att[t], mx));
    consumer_bar_sync();
    if (tid == 0) {
        float sum = 0.0f;
        for (int t = 0; t < nt; t++) sum = __fadd_rn(sum, att[t]);
        s_val[0] = sum;
    }
    consumer_bar_sync();
    const float sum = s_val[0];
    for (int t = tid; t < nt; t += PD_CT) att[t] = __fdiv_rn(att[t], sum);
    consumer_bar_sync();
    if (tid < HS) {
        const float *v = vc + kvh * HS + tid;
        float acc = 0.0f;
        int t = 0;
        for (; t + 16 <= pos; t += 16) {
            float vv[16];
#pragma unroll
            for (int u = 0; u < 16; u++) vv[u] = __ldcg(v + (size_t)(t + u) * kvd);
#pragma unroll
            for (int u = 0; u < 16; u++) acc = __fadd_rn(__fmul_rn(att[t + u], vv[u]), acc);
        }
        for (; t < pos; t++) acc = __fadd_rn(__fmul_rn(att[t], __ldcg(v + (size_t)t * kvd)), acc);
        acc = __fadd_rn(__fmul_rn(att[pos], ldcg_f32c(vsrc + tid)), acc);
        so[tid] = acc;
    }
    consumer_bar_sync();
    for (int b = warp; b < HS / 32; b += SMV_CONSUMER_WARPS) {
        float as;
        const int q = quant_block_lane(so[b * 32 + lane], as);
        a.attq[h * HS + b * 32 + lane] = (int8_t)q;
        if (lane == 0) a.atts[(h * HS) / 32 + b] = as;
    }
}

// ---- the kernel -----------------------------------------------------------------------------------------------------------
template <int HS>
__global__ void __launch_bounds__(SMV_THREADS, 1) k_decode_persistent(PdArgs a, PdSmem L) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = L.stages;
    const unsigned bar0 = smem_u32(smem + L.off_bar);
    volatile unsigned *rel = reinterpret_cast<volatile unsigned *>(smem + L.off_bar + 2 * PD_MAX_STAGES * 8);
    if (tid == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(bar0 + 8 * s, 1);
            mbar_init(bar0 + 8 * (PD_MAX_STAGES + s), 1);
            rel[s] = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == SMV_CONSUMER_WARPS) { // ===== producer =====
        if (lane == 0) {
            unsigned seq = 0;
            for (int l = 0; l < a.n_layers; l++) {
                const PdLayer &Ly = a.layers[l];
                pd_produce_matrix(Ly.qkv, smem, L, bar0, seq);
                pd_produce_matrix(Ly.wo, smem, L, bar0, seq);
                pd_produce_matrix(Ly.gu, smem, L, bar0, seq);
                pd_produce_matrix(Ly.w2, smem, L, bar0, seq);
            }
            if (a.with_logits) pd_produce_matrix(a.lm_head, smem, L, bar0, seq);
        }
        return;
    }

    // ===== consumers =====
    const int token = a.st->token, pos = a.st->pos;
    const unsigned tick = *reinterpret_cast<volatile unsigned *>(a.sync + PD_S_TICK);
    const unsigned lmtick = *reinterpret_cast<volatile unsigned *>(a.sync + PD_S_LMTICK);
    const unsigned nC = gridDim.x, nL = (unsigned)a.n_layers;
    unsigned seq_base = 0;
    for (int l = 0; l < a.n_layers; l++) {
        const PdLayer &Ly = a.layers[l];
        const unsigned e = tick * nL + (unsigned)l + 1u; // this layer's epoch
        pd_norm_to_smem(a, Ly.attn_norm, l == 0, token, smem, L, tid);
        pd_consume_matrix<SMV_STORE>(Ly.qkv, a, smem, L, bar0, rel, seq_base, a.qkv, false, false, token, tid);
        pd_arrive(a.sync + PD_S_QKV, tid);
        if ((int)blockIdx.x < a.n_heads) { // attention: the first n_heads CTAs, one head each
            pd_wait(a.sync + PD_S_QKV, e * nC, tid);
            pd_attention_head<HS>(a, Ly, blockIdx.x, pos, smem, L, tid);
            pd_arrive(a.sync + PD_S_ATT, tid);
        }
        pd_wait(a.sync + PD_S_ATT, e * (unsigned)a.n_heads, tid);
        pd_load_act(a.attq, a.atts, a.qd, smem, L, tid);
        pd_consume_matrix<SMV_RESID>(Ly.wo, a, smem, L, bar0, rel, seq_base, a.x, false, l == 0, token, tid);
        pd_arrive(a.sync + PD_S_WO, tid);
        pd_wait(a.sync + PD_S_WO, e * nC, tid);
        pd_norm_to_smem(a, Ly.ffn_norm, false, token, smem, L, tid);
        pd_consume_matrix<SMV_GATEUP>(Ly.gu, a, smem, L, bar0, rel, seq_base, a.hb, false, false, token, tid);
        pd_arrive(a.sync + PD_S_GU, tid);
        pd_wait(a.sync + PD_S_GU, e * nC, tid);
        pd_load_act(a.hq, a.hs, a.hidden, smem, L, tid);
        pd_consume_matrix<SMV_RESID>(Ly.w2, a, smem, L, bar0, rel, seq_base, a.x, false, false, token, tid);
        pd_arrive(a.sync + PD_S_W2, tid);
        pd_wait(a.sync + PD_S_W2, e * nC, tid);
    }
    int best_i = 0;
    if (a.with_logits) {
        pd_norm_to_smem(a, a.out_norm, false, token, smem, L, tid);
        pd_consume_matrix<SMV_STORE>(a.lm_head, a, smem, L, bar0, rel, seq_base, a.logits, true, false, token, tid);
        pd_arrive(a.sync + PD_S_LM, tid);
        if (blockIdx.x != 0) return;
        pd_wait(a.sync + PD_S_LM, (lmtick + 1u) * nC, tid);
        // FloatTensor.argmax over the per-CTA (max, first index) pairs: copy of k_argmax_advance
        float best = -INFINITY;
        best_i = 0x7fffffff;
        for (int i = tid; i < (int)nC; i += PD_CT) argmax_merge(best, best_i, ldcg_f32c(a.part_val + i), __ldcg(a.part_idx + i));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            argmax_merge(best, best_i, ov, oi);
        }
        float *sv = reinterpret_cast<float *>(smem + L.off_misc) + 24;
        int *si = reinterpret_cast<int *>(sv + 8);
        if (lane == 0) { sv[warp] = best; si[warp] = best_i; }
        consumer_bar_sync();
        if (tid == 0) {
            best = sv[0]; best_i = si[0];
            for (int w = 1; w < SMV_CONSUMER_WARPS; w++) argmax_merge(best, best_i, sv[w], si[w]);
            if (best_i == 0x7fffffff) best_i = 0;
        }
    } else if (blockIdx.x != 0) {
        return;
    }
    if (tid == 0) { // step advance (k_argmax_advance's tail) + the next launch's epoch
        StepState *st = a.st;
        const int step = st->step;
        if (a.with_logits && a.out_ids) a.out_ids[step] = best_i;
        const int next = step + 1;
        if (st->feedback && a.with_logits) st->token = best_i;
        else if (next < st->n_seq) st->token = a.seq_tokens[next];
        st->step = next;
        st->pos = st->pos + 1;
        *reinterpret_cast<volatile unsigned *>(a.sync + PD_S_TICK) = tick + 1u;
        if (a.with_logits) *reinterpret_cast<volatile unsigned *>(a.sync + PD_S_LMTICK) = lmtick + 1u;
    }
}
