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
        k_rel_attention_tiled<NC><<<grid, AT_THREADS, smem, st>>>(qkv, Ek, Ev, out, cu, tile_cu64, B, H, window);     \
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
