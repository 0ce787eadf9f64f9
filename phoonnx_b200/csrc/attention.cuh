// Tiled relative-position multi-head self-attention, fp32 (attentions.py:225-272; band restatement
// SURVEY.md A1).  Flash-style: one CTA owns 64 query rows of one (utterance, head), streams the
// utterance's keys/values through shared memory 64 at a time with an online softmax, and never
// materialises the T x T score matrix in HBM (the reference builds [B, heads, T, T] plus a
// [B, heads, T, 2T-1] skewed relative-logit tensor).
//
//   scores[i,j] = (q_i / sqrt(dk)) . k_j + [|j-i| <= w] (q_i / sqrt(dk)) . Ek[j-i+w]
//   out_i       = sum_j p_ij v_j + sum_{|j-i|<=w} p_ij Ev[j-i+w]
//
// Stays fp32 on CUDA cores: the text side feeds ceil(exp(logw)) and must not perturb durations
// (SURVEY.md A9); its FLOPs are ~1% of the path.  128 threads, 8x4 register tile for q.k^T and
// 8 x (dk/16) for p.v; q and k are staged transposed so the inner loops are broadcast /
// conflict-free shared-memory reads.
#pragma once
#include "common.cuh"

#define AT_QT 64
#define AT_KT 64
#define AT_LD 68          // padded leading dimension of the transposed tiles (16 B aligned rows)
#define AT_THREADS 128
#define AT_MAXREL 16      // 2*window+1 <= 16

static inline size_t attention_smem_bytes(int dk, int nrel) {
    return sizeof(float) * ((size_t)dk * AT_LD            // Qt
                            + (size_t)(dk > AT_KT ? dk : AT_KT) * AT_LD   // Kt [dk][AT_LD], aliased by Pt [AT_KT][AT_LD]
                            + (size_t)AT_KT * dk           // Vs
                            + (size_t)AT_QT * AT_MAXREL    // qe
                            + 2 * (size_t)nrel * dk);      // Ek, Ev
}

template <int NC>   // dk == 16 * NC
__global__ void __launch_bounds__(AT_THREADS) k_rel_attention_tiled(
    const float* __restrict__ qkv, const float* __restrict__ Ek, const float* __restrict__ Ev,
    float* __restrict__ out, const int* __restrict__ cu, const int* __restrict__ tile_cu, int B, int H, int window) {
    pdl_enter();
    constexpr int DK = 16 * NC;
    extern __shared__ __align__(16) float sm_att[];
    float* Qt = sm_att;                       // [DK][AT_LD]  q^T / sqrt(dk)
    float* Kt = Qt + DK * AT_LD;              // [DK][AT_LD]  k^T ; later Pt [AT_KT][AT_LD]
    float* Pt = Kt;
    float* Vs = Kt + (DK > AT_KT ? DK : AT_KT) * AT_LD;   // [AT_KT][DK]
    float* qe = Vs + AT_KT * DK;              // [AT_QT][AT_MAXREL]
    float* Eks = qe + AT_QT * AT_MAXREL;      // [nrel][DK]
    float* Evs = Eks + (2 * window + 1) * DK;

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nrel = 2 * window + 1;
    const int head = blockIdx.y;
    const int tile = blockIdx.x;
    const int b = find_segment(tile_cu, B, tile);
    const int q0 = (tile - __ldg(tile_cu + b)) * AT_QT;
    const int r0 = __ldg(cu + b), T = __ldg(cu + b + 1) - r0;
    const int ld = 3 * H;
    const float* qbase = qkv + (long)r0 * ld + head * DK;
    const float* kbase = qbase + H;
    const float* vbase = qbase + 2 * H;
    const float scale = 1.f / sqrtf((float)DK);      // attentions.py:232

    // ---- stage q^T (scaled), Ek, Ev
    for (int idx = tid; idx < AT_QT * (DK / 4); idx += AT_THREADS) {
        const int r = idx & (AT_QT - 1), d4 = idx >> 6;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < T) v = __ldg(reinterpret_cast<const float4*>(qbase + (long)(q0 + r) * ld) + d4);
        Qt[(4 * d4 + 0) * AT_LD + r] = v.x * scale; Qt[(4 * d4 + 1) * AT_LD + r] = v.y * scale;
        Qt[(4 * d4 + 2) * AT_LD + r] = v.z * scale; Qt[(4 * d4 + 3) * AT_LD + r] = v.w * scale;
    }
    for (int idx = tid; idx < nrel * DK; idx += AT_THREADS) { Eks[idx] = __ldg(Ek + idx); Evs[idx] = __ldg(Ev + idx); }
    __syncthreads();
    // relative-key logits of this q tile: qe[i][r] = q_i . Ek[r]
    for (int idx = tid; idx < AT_QT * nrel; idx += AT_THREADS) {
        const int i = idx & (AT_QT - 1), r = idx >> 6;
        float s = 0.f;
#pragma unroll 8
        for (int d = 0; d < DK; d++) s = fmaf(Qt[d * AT_LD + i], Eks[r * DK + d], s);
        qe[i * AT_MAXREL + r] = s;
    }

    float o[8][NC];
    float mrun[8], lrun[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        mrun[i] = -INFINITY; lrun[i] = 0.f;
#pragma unroll
        for (int c = 0; c < NC; c++) o[i][c] = 0.f;
    }

    for (int k0 = 0; k0 < T; k0 += AT_KT) {
        __syncthreads();                        // previous tile's Pt / Vs reads are done (also orders the qe writes)
        for (int idx = tid; idx < AT_KT * (DK / 4); idx += AT_THREADS) {
            const int r = idx & (AT_KT - 1), d4 = idx >> 6;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (k0 + r < T) {
                kv = __ldg(reinterpret_cast<const float4*>(kbase + (long)(k0 + r) * ld) + d4);
                vv = __ldg(reinterpret_cast<const float4*>(vbase + (long)(k0 + r) * ld) + d4);
            }
            Kt[(4 * d4 + 0) * AT_LD + r] = kv.x; Kt[(4 * d4 + 1) * AT_LD + r] = kv.y;
            Kt[(4 * d4 + 2) * AT_LD + r] = kv.z; Kt[(4 * d4 + 3) * AT_LD + r] = kv.w;
            *reinterpret_cast<float4*>(Vs + r * DK + 4 * d4) = vv;
        }
        __syncthreads();
        // ---- scores: rows 8*ty + i, keys tx + 16*c
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int c = 0; c < 4; c++) s[i][c] = 0.f;
#pragma unroll 4
        for (int d = 0; d < DK; d++) {
            const float4 qa = *reinterpret_cast<const float4*>(Qt + d * AT_LD + 8 * ty);
            const float4 qb = *reinterpret_cast<const float4*>(Qt + d * AT_LD + 8 * ty + 4);
            const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            float kv[4];
#pragma unroll
            for (int c = 0; c < 4; c++) kv[c] = Kt[d * AT_LD + tx + 16 * c];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int c = 0; c < 4; c++) s[i][c] = fmaf(qv[i], kv[c], s[i][c]);
        }
        const bool near_diag = (k0 <= q0 + AT_QT - 1 + window) && (k0 + AT_KT - 1 >= q0 - window);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int qi = q0 + 8 * ty + i;
            float tmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int j = k0 + tx + 16 * c;
                if (near_diag) {
                    const int rel = j - qi + window;
                    if (rel >= 0 && rel < nrel) s[i][c] += qe[(8 * ty + i) * AT_MAXREL + rel];
                }
                if (j >= T) s[i][c] = -INFINITY;
                tmax = fmaxf(tmax, s[i][c]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
            const float mnew = fmaxf(mrun[i], tmax);          // finite: every key tile holds >= 1 valid key
            const float corr = expf(mrun[i] - mnew);          // exp(-inf) = 0 on the first tile
            float psum = 0.f;
#pragma unroll
            for (int c = 0; c < 4; c++) { s[i][c] = expf(s[i][c] - mnew); psum += s[i][c]; }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
            lrun[i] = lrun[i] * corr + psum;
            mrun[i] = mnew;
#pragma unroll
            for (int c = 0; c < NC; c++) o[i][c] *= corr;
        }
        __syncthreads();                        // every thread is done reading Kt before Pt overwrites it
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float* dst = Pt + (tx + 16 * c) * AT_LD + 8 * ty;
            *reinterpret_cast<float4*>(dst) = make_float4(s[0][c], s[1][c], s[2][c], s[3][c]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(s[4][c], s[5][c], s[6][c], s[7][c]);
        }
        __syncthreads();
        // ---- out += p . v   (rows 8*ty + i, channels tx + 16*c)
        const int jn = min(AT_KT, T - k0);
#pragma unroll 2
        for (int j = 0; j < jn; j++) {
            const float4 pa = *reinterpret_cast<const float4*>(Pt + j * AT_LD + 8 * ty);
            const float4 pb = *reinterpret_cast<const float4*>(Pt + j * AT_LD + 8 * ty + 4);
            const float pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
            float vv[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) vv[c] = Vs[j * DK + tx + 16 * c];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int c = 0; c < NC; c++) o[i][c] = fmaf(pv[i], vv[c], o[i][c]);
        }
        // ---- relative-value band: out_i += p_ij Ev[j-i+w]
        if (near_diag) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int qi = q0 + 8 * ty + i;
                for (int r = 0; r < nrel; r++) {
                    const int jl = qi + r - window - k0;
                    if (jl < 0 || jl >= jn) continue;
                    const float p = Pt[jl * AT_LD + 8 * ty + i];
#pragma unroll
                    for (int c = 0; c < NC; c++) o[i][c] = fmaf(p, Evs[r * DK + tx + 16 * c], o[i][c]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int qi = q0 + 8 * ty + i;
        if (qi >= T) continue;
        const float inv = 1.f / lrun[i];
        float* orow = out + (long)(r0 + qi) * H + head * DK;
#pragma unroll
        for (int c = 0; c < NC; c++) orow[tx + 16 * c] = o[i][c] * inv;
    }
}

// returns false when the shape is outside this kernel (caller falls back to k_rel_attention)
static inline bool attention_tiled_launch(const float* qkv, const float* Ek, const float* Ev, float* out, const int* cu,
                                          const int* tile_cu64, int ntiles64, int B, int H, int n_heads, int dk, int window,
                                          cudaStream_t st, cudaError_t* err) {
    *err = cudaSuccess;
    if (dk % 16 || dk > 128 || 2 * window + 1 > AT_MAXREL || (3 * H) % 4 || dk % 4) return false;
    if (ntiles64 <= 0) return true;
    const size_t smem = attention_smem_bytes(dk, 2 * window + 1);
    const dim3 grid(ntiles64, n_heads);
#define AT_CASE(NC)                                                                                                   \
    case NC: {                                                                                                        \
        static bool attr_done[64] = {false};                                                                          \
        int dev = 0; cudaGetDevice(&dev);                                                                             \
        if (dev >= 0 && dev < 64 && !attr_done[dev]) {                                                                \
            *err = cudaFuncSetAttribute(k_rel_attention_tiled<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
            if (*err != cudaSuccess) return true;                                                                     \
            attr_done[dev] = true;                                                                                    \
        }                                                                                                             \
        launch_k(k_rel_attention_tiled<NC>, grid, AT_THREADS, smem, st, qkv, Ek, Ev, out, cu, tile_cu64, B, H, window);     \
        break;                                                                                                        \
    }
    switch (dk / 16) {
        AT_CASE(1) AT_CASE(2) AT_CASE(3) AT_CASE(4) AT_CASE(6) AT_CASE(8)
        default: return false;
    }
#undef AT_CASE
    *err = cudaGetLastError();
    return true;
}

// ------------------------------------------------------------------------------------------------------------------
// Tensor-core variant (bf16 mode): the same flash-style schedule with both contractions on mma.sync.m16n8k16 as an
// fp32-faithful bf16x3 product (the text side's precision policy):
//     q.k^T ~= qh.kh + qh.kl + ql.kh          p.v ~= ph.vh + ph.vl + pl.vh          (xh = bf16(x), xl = bf16(x - xh))
// with fp32 accumulation, fp32 softmax (expf), and the 9-wide relative-position band (q.Ek logits, p.Ev values) on CUDA
// cores in fp32 exactly as above.  One CTA = 64 query rows of one (utterance, head); 4 warps x 16 rows; keys / values
// stream through shared memory 64 at a time, split into hi / lo planes while being staged.  The accumulator fragment
// of the score MMA is the A fragment of the value MMA, so P never leaves registers.
// r01d: the fp32 CUDA-core kernel above ran at 13 TFLOP/s and was 9.5 % of the whole path.
// ------------------------------------------------------------------------------------------------------------------
#define AT2_THREADS 128
#define AT2_PBP 12        // floats per row of the per-warp band buffer

namespace at2 {
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x, y) -> packed bf16 hi pair and the packed bf16 of the remainders
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h); lo = *reinterpret_cast<const uint32_t*>(&l);
}
// stage a [64 x DK] fp32 tile (rows r0.. of a [*, ld] matrix, zero beyond nvalid rows) as bf16 hi / lo planes [64][DK + 8]
template <int DK>
__device__ __forceinline__ void stage_split(const float* __restrict__ base, long ld, int nvalid, float scale,
                                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int tid) {
    constexpr int P = DK + 8, D4 = DK / 4;
    constexpr int NIT = 64 * D4 / AT2_THREADS;          // exact for DK in {32, 48, 64, 96}
    static_assert(NIT * AT2_THREADS == 64 * D4, "tile does not divide over the CTA");
    // every load of the tile is in flight before the first conversion (one memory round trip per tile: with load -> split ->
    // store per element the r01e profile had > 50 % of all stall samples on these loads)
    float4 v[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int idx = tid + it * AT2_THREADS;
        const int r = idx / D4, d4 = idx - r * D4;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nvalid) v[it] = __ldg(reinterpret_cast<const float4*>(base + (long)r * ld) + d4);
    }
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int idx = tid + it * AT2_THREADS;
        const int r = idx / D4, d4 = idx - r * D4;
        uint2 h, l;
        split2(v[it].x * scale, v[it].y * scale, h.x, l.x); split2(v[it].z * scale, v[it].w * scale, h.y, l.y);
        *reinterpret_cast<uint2*>(hi + r * P + 4 * d4) = h;
        *reinterpret_cast<uint2*>(lo + r * P + 4 * d4) = l;
    }
}
// two tiles (K and V) with every load of both in flight before the first conversion: one memory round trip per key tile
template <int DK>
__device__ __forceinline__ void stage_split2(const float* __restrict__ base_a, const float* __restrict__ base_b, long ld, int nvalid,
                                             __nv_bfloat16* __restrict__ hi_a, __nv_bfloat16* __restrict__ lo_a,
                                             __nv_bfloat16* __restrict__ hi_b, __nv_bfloat16* __restrict__ lo_b, int tid) {
    constexpr int P = DK + 8, D4 = DK / 4;
    constexpr int NIT = 64 * D4 / AT2_THREADS;
    float4 va[NIT], vb[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int idx = tid + it * AT2_THREADS;
        const int r = idx / D4, d4 = idx - r * D4;
        va[it] = make_float4(0.f, 0.f, 0.f, 0.f); vb[it] = va[it];
        if (r < nvalid) {
            va[it] = __ldg(reinterpret_cast<const float4*>(base_a + (long)r * ld) + d4);
            vb[it] = __ldg(reinterpret_cast<const float4*>(base_b + (long)r * ld) + d4);
        }
    }
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int idx = tid + it * AT2_THREADS;
        const int r = idx / D4, d4 = idx - r * D4;
        uint2 h, l;
        split2(va[it].x, va[it].y, h.x, l.x); split2(va[it].z, va[it].w, h.y, l.y);
        *reinterpret_cast<uint2*>(hi_a + r * P + 4 * d4) = h;
        *reinterpret_cast<uint2*>(lo_a + r * P + 4 * d4) = l;
        split2(vb[it].x, vb[it].y, h.x, l.x); split2(vb[it].z, vb[it].w, h.y, l.y);
        *reinterpret_cast<uint2*>(hi_b + r * P + 4 * d4) = h;
        *reinterpret_cast<uint2*>(lo_b + r * P + 4 * d4) = l;
    }
}
}  // namespace at2

static inline size_t attention_mma_smem_bytes(int dk, int nrel) {
    return (size_t)(6 * 64 + 2 * 16) * (dk + 8) * 2 + sizeof(float) * ((size_t)64 * AT_MAXREL + (size_t)nrel * dk + 4 * 16 * AT2_PBP);
}

template <int DK>
__global__ void __launch_bounds__(AT2_THREADS, 2) k_rel_attention_mma(
    const float* __restrict__ qkv, const float* __restrict__ Ek, const float* __restrict__ Ev,
    float* __restrict__ out, const int4* __restrict__ tdesc64, int H, int window) {
    pdl_enter();
    constexpr int P = DK + 8;                 // bf16 elements per staged row: 16-byte segments of 8 consecutive rows hit 8 distinct bank groups
    constexpr int NTO = DK / 8;               // output n-tiles (8 channels each)
    extern __shared__ __align__(16) uint8_t sm_at2[];
    __nv_bfloat16* sQh = reinterpret_cast<__nv_bfloat16*>(sm_at2);
    __nv_bfloat16* sQl = sQh + 64 * P;
    __nv_bfloat16* sKh = sQl + 64 * P;
    __nv_bfloat16* sKl = sKh + 64 * P;
    __nv_bfloat16* sVh = sKl + 64 * P;
    __nv_bfloat16* sVl = sVh + 64 * P;
    __nv_bfloat16* sEh = sVl + 64 * P;                       // relative-key embeddings, bf16 hi / lo planes [16][P] (rows >= nrel zero)
    __nv_bfloat16* sEl = sEh + 16 * P;
    float* qe = reinterpret_cast<float*>(sEl + 16 * P);      // [64][AT_MAXREL]
    const int nrel = 2 * window + 1;
    float* Evs = qe + 64 * AT_MAXREL;                        // [nrel][DK]
    float* sPB = Evs + nrel * DK;                            // [4 warps][16][AT2_PBP]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int head = blockIdx.y, tile = blockIdx.x;
    const int4 dsc = __ldg(tdesc64 + tile);          // {utterance's first row, its rows, tile's first row in it, utterance} (k_tile_desc)
    const int q0 = dsc.z, r0 = dsc.x, T = dsc.y;
    const int ld = 3 * H;
    const float* qbase = qkv + (long)r0 * ld + head * DK;
    const float* kbase = qbase + H;
    const float* vbase = qbase + 2 * H;
    const float scale = 1.f / sqrtf((float)DK);      // attentions.py:232

    at2::stage_split<DK>(qbase + (long)q0 * ld, ld, T - q0, scale, sQh, sQl, tid);
    for (int idx = tid; idx < nrel * DK; idx += AT2_THREADS) Evs[idx] = __ldg(Ev + idx);
    for (int idx = tid; idx < 16 * (DK / 2); idx += AT2_THREADS) {
        const int r = idx / (DK / 2), d2 = idx - r * (DK / 2);
        float2 e = make_float2(0.f, 0.f);
        if (r < nrel) e = __ldg(reinterpret_cast<const float2*>(Ek + r * DK) + d2);
        uint32_t eh, el;
        at2::split2(e.x, e.y, eh, el);
        *reinterpret_cast<uint32_t*>(sEh + r * P + 2 * d2) = eh;
        *reinterpret_cast<uint32_t*>(sEl + r * P + 2 * d2) = el;
    }
    __syncthreads();

    float o[NTO][4];
#pragma unroll
    for (int n = 0; n < NTO; n++) { o[n][0] = 0.f; o[n][1] = 0.f; o[n][2] = 0.f; o[n][3] = 0.f; }
    float mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.f, 0.f};
    const int mi = lane >> 3, l7 = lane & 7;
    // ldmatrix lane addresses (bytes, relative to the plane): A = this warp's 16 query rows; B(K) = key row l7 of n-tile (mi >> 1),
    // k half (mi & 1); B(V, transposed) = key row (mi & 1) * 8 + l7, channel block (mi >> 1)
    const uint32_t a_off = (uint32_t)(((warp * 16 + l7 + (mi & 1) * 8) * P + (mi >> 1) * 8) * 2);
    const uint32_t k_off = (uint32_t)((((mi >> 1) * 8 + l7) * P + (mi & 1) * 8) * 2);
    const uint32_t v_off = (uint32_t)((((mi & 1) * 8 + l7) * P + (mi >> 1) * 8) * 2);
    const uint32_t uQh = at2::s_u32(sQh), uQl = at2::s_u32(sQl), uKh = at2::s_u32(sKh), uKl = at2::s_u32(sKl),
                   uVh = at2::s_u32(sVh), uVl = at2::s_u32(sVl);
    float* pb = sPB + warp * 16 * AT2_PBP;
    // relative-key logits of this warp's 16 query rows: qe[i][r] = (q_i / sqrt(dk)) . Ek[r] -- one more bf16x3 product (the 16 padded
    // "keys" are the relative embeddings); rows of qe are only ever read by the warp that wrote them
    {
        float c2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const uint32_t uEh = at2::s_u32(sEh), uEl = at2::s_u32(sEl);
#pragma unroll
        for (int ks = 0; ks < DK / 16; ks++) {
            uint32_t ah[4], al[4], bh[4], bl[4];
            at2::ldsm_x4(uQh + a_off + ks * 32, ah[0], ah[1], ah[2], ah[3]);
            at2::ldsm_x4(uQl + a_off + ks * 32, al[0], al[1], al[2], al[3]);
            at2::ldsm_x4(uEh + k_off + ks * 32, bh[0], bh[1], bh[2], bh[3]);
            at2::ldsm_x4(uEl + k_off + ks * 32, bl[0], bl[1], bl[2], bl[3]);
            at2::mma16816(c2[0], al, bh[0], bh[1]); at2::mma16816(c2[0], ah, bl[0], bl[1]); at2::mma16816(c2[0], ah, bh[0], bh[1]);
            at2::mma16816(c2[1], al, bh[2], bh[3]); at2::mma16816(c2[1], ah, bl[2], bl[3]); at2::mma16816(c2[1], ah, bh[2], bh[3]);
        }
#pragma unroll
        for (int n = 0; n < 2; n++)
#pragma unroll
            for (int h = 0; h < 2; h++)
                *reinterpret_cast<float2*>(qe + (warp * 16 + g + 8 * h) * AT_MAXREL + n * 8 + 2 * t) = make_float2(c2[n][2 * h], c2[n][2 * h + 1]);
        __syncwarp();
    }

    for (int k0 = 0; k0 < T; k0 += 64) {
        __syncthreads();                        // the previous tile's K / V reads are done (first pass: qe is complete)
        at2::stage_split2<DK>(kbase + (long)k0 * ld, vbase + (long)k0 * ld, ld, T - k0, sKh, sKl, sVh, sVl, tid);
        __syncthreads();
        // ---- scores: 16 rows x 64 keys per warp
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; n++) { s[n][0] = 0.f; s[n][1] = 0.f; s[n][2] = 0.f; s[n][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < DK / 16; ks++) {
            uint32_t ah[4], al[4];
            at2::ldsm_x4(uQh + a_off + ks * 32, ah[0], ah[1], ah[2], ah[3]);
            at2::ldsm_x4(uQl + a_off + ks * 32, al[0], al[1], al[2], al[3]);
#pragma unroll
            for (int np = 0; np < 4; np++) {
                uint32_t bh[4], bl[4];
                const uint32_t off = k_off + (uint32_t)(np * 16 * P * 2 + ks * 32);
                at2::ldsm_x4(uKh + off, bh[0], bh[1], bh[2], bh[3]);
                at2::ldsm_x4(uKl + off, bl[0], bl[1], bl[2], bl[3]);
                at2::mma16816(s[2 * np], al, bh[0], bh[1]);
                at2::mma16816(s[2 * np], ah, bl[0], bl[1]);
                at2::mma16816(s[2 * np], ah, bh[0], bh[1]);
                at2::mma16816(s[2 * np + 1], al, bh[2], bh[3]);
                at2::mma16816(s[2 * np + 1], ah, bl[2], bl[3]);
                at2::mma16816(s[2 * np + 1], ah, bh[2], bh[3]);
            }
        }
        // ---- relative-key band, key mask, online softmax (fp32); thread: rows g and g + 8, keys nt * 8 + 2t + {0, 1}
        const bool near_diag = (k0 <= q0 + 63 + window) && (k0 + 63 >= q0 - window);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int il = warp * 16 + g + 8 * h, qi = q0 + il;
            float tmax = -INFINITY;
#pragma unroll
            for (int n = 0; n < 8; n++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int j = k0 + n * 8 + 2 * t + e;
                    float v = s[n][2 * h + e];
                    if (near_diag) {
                        const int rel = j - qi + window;
                        if (rel >= 0 && rel < nrel) v += qe[il * AT_MAXREL + rel];
                    }
                    if (j >= T) v = -INFINITY;
                    s[n][2 * h + e] = v;
                    tmax = fmaxf(tmax, v);
                }
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
            const float mnew = fmaxf(mrun[h], tmax);          // finite: every key tile holds >= 1 valid key
            const float corr = expf(mrun[h] - mnew);          // exp(-inf) = 0 on the first tile
            float psum = 0.f;
#pragma unroll
            for (int n = 0; n < 8; n++)
#pragma unroll
                for (int e = 0; e < 2; e++) { const float p = expf(s[n][2 * h + e] - mnew); s[n][2 * h + e] = p; psum += p; }
            psum += __shfl_xor_sync(0xffffffffu, psum, 1);
            psum += __shfl_xor_sync(0xffffffffu, psum, 2);
            lrun[h] = lrun[h] * corr + psum;
            mrun[h] = mnew;
#pragma unroll
            for (int n = 0; n < NTO; n++) { o[n][2 * h] *= corr; o[n][2 * h + 1] *= corr; }
        }
        // ---- relative-value band: out_i += sum_{|j-i|<=w} p_ij Ev[j-i+w]; the band entries change hands through a per-warp buffer
        if (near_diag) {
            for (int idx = lane; idx < 16 * AT2_PBP; idx += 32) pb[idx] = 0.f;
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int qi = q0 + warp * 16 + g + 8 * h;
#pragma unroll
                for (int n = 0; n < 8; n++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int rel = k0 + n * 8 + 2 * t + e - qi + window;
                        if (rel >= 0 && rel < nrel) pb[(g + 8 * h) * AT2_PBP + rel] = s[n][2 * h + e];
                    }
            }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; h++) {
                for (int r = 0; r < nrel; r++) {
                    const float p = pb[(g + 8 * h) * AT2_PBP + r];
                    if (p != 0.f) {
#pragma unroll
                        for (int n = 0; n < NTO; n++) {
                            const float2 ev = *reinterpret_cast<const float2*>(Evs + r * DK + n * 8 + 2 * t);
                            o[n][2 * h] = fmaf(p, ev.x, o[n][2 * h]);
                            o[n][2 * h + 1] = fmaf(p, ev.y, o[n][2 * h + 1]);
                        }
                    }
                }
            }
            __syncwarp();
        }
        // ---- out += p . v: the score accumulators of key n-tiles (2kk, 2kk + 1) are the A fragment of key step kk
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            uint32_t ph[4], pl[4];
            at2::split2(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
            at2::split2(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
            at2::split2(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
            at2::split2(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int np = 0; np < NTO / 2; np++) {
                uint32_t vh[4], vl[4];
                const uint32_t off = v_off + (uint32_t)(kk * 16 * P * 2 + np * 32);
                at2::ldsm_x4_t(uVh + off, vh[0], vh[1], vh[2], vh[3]);
                at2::ldsm_x4_t(uVl + off, vl[0], vl[1], vl[2], vl[3]);
                at2::mma16816(o[2 * np], pl, vh[0], vh[1]);
                at2::mma16816(o[2 * np], ph, vl[0], vl[1]);
                at2::mma16816(o[2 * np], ph, vh[0], vh[1]);
                at2::mma16816(o[2 * np + 1], pl, vh[2], vh[3]);
                at2::mma16816(o[2 * np + 1], ph, vl[2], vl[3]);
                at2::mma16816(o[2 * np + 1], ph, vh[2], vh[3]);
            }
        }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int qi = q0 + warp * 16 + g + 8 * h;
        if (qi >= T) continue;
        const float inv = 1.f / lrun[h];
        float* orow = out + (long)(r0 + qi) * H + head * DK + 2 * t;
#pragma unroll
        for (int n = 0; n < NTO; n++) *reinterpret_cast<float2*>(orow + n * 8) = make_float2(o[n][2 * h] * inv, o[n][2 * h + 1] * inv);
    }
}

// returns false when the shape is outside this kernel (caller falls back to the fp32 kernels)
static inline bool attention_mma_launch(const float* qkv, const float* Ek, const float* Ev, float* out, const int4* tdesc64,
                                        int ntiles64, int H, int n_heads, int dk, int window, cudaStream_t st, cudaError_t* err) {
    *err = cudaSuccess;
    if ((dk != 96 && dk != 48 && dk != 64 && dk != 32) || 2 * window + 1 > AT2_PBP || 2 * window + 1 > AT_MAXREL || (3 * H) % 4 || H % 2) return false;
    if (ntiles64 <= 0) return true;
    const size_t smem = attention_mma_smem_bytes(dk, 2 * window + 1);
    const dim3 grid(ntiles64, n_heads);
#define AT2_CASE(DKV)                                                                                                 \
    case DKV: {                                                                                                       \
        static bool attr_done[64] = {false};                                                                          \
        int dev = 0; cudaGetDevice(&dev);                                                                             \
        if (dev >= 0 && dev < 64 && !attr_done[dev]) {                                                                \
            *err = cudaFuncSetAttribute(k_rel_attention_mma<DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); \
            if (*err != cudaSuccess) return true;                                                                     \
            attr_done[dev] = true;                                                                                    \
        }                                                                                                             \
        launch_k(k_rel_attention_mma<DKV>, grid, AT2_THREADS, smem, st, qkv, Ek, Ev, out, tdesc64, H, window);              \
        break;                                                                                                        \
    }
    switch (dk) {
        AT2_CASE(96) AT2_CASE(48) AT2_CASE(64) AT2_CASE(32)
        default: return false;
    }
#undef AT2_CASE
    *err = cudaGetLastError();
    return true;
}
