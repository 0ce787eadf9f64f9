// fp32 CUDA-core kernels of the VITS hot path (parity mode + every non-contraction stage).
// Layout everywhere: packed varlen, channel-last  [sum_rows, C]  (C contiguous).
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// Generic implicit-GEMM conv1d, fp32, 64x64 tile, 256 threads, 4x4 micro-tile.
// Covers every Conv1d / polyphase ConvTranspose1d on the path (SURVEY.md 8a rows 3,5,7,10-12,
// 15,18-22).  HBM/L2-bound for the narrow layers, FFMA-bound for the wide ones.
// ------------------------------------------------------------------------------------------
#define CF_TM 64
#define CF_TN 64
#define CF_KC 16

__device__ __forceinline__ void conv_epilogue_store(const ConvArgs& a, int b, long row, int n, float v) {
    // v already holds acc; applies bias / per-utterance bias / residual / activation / store
    if (a.bias) v += __ldg(a.bias + n);
    if (a.utab) v += __ldg(a.utab + (long)__ldg(a.uidx + b) * a.utab_ld + n);
    if (a.epi == EPI_SUBFROM) {
        float r = a.res[row * a.ldres + a.rescol + n];
        a.out[row * a.ldo + a.ocol + n] = r - v;
        return;
    }
    if (a.res) v += a.res[row * a.ldres + a.rescol + n];
    if (a.out_act == ACT_RELU) v = fmaxf(v, 0.f);
    float* dst;
    int acc;
    if (a.epi == EPI_SPLIT && n >= a.split) {
        dst = a.out2 + row * a.ldo2 + a.ocol2 + (n - a.split);
        acc = a.accumulate2;
    } else {
        dst = a.out + row * a.ldo + a.ocol + n;
        acc = a.accumulate;
    }
    if (acc) v += *dst;
    if (a.out_div != 1.f) v = v / a.out_div;
    if (a.out_act == ACT_TANH) v = tanhf(v);
    *dst = v;
}

__global__ void __launch_bounds__(256) k_conv_f32(const ConvArgs a) {
    pdl_enter();
    __shared__ __align__(16) float As[CF_KC][CF_TM + 4];
    __shared__ __align__(16) float Bs[CF_KC][CF_TN];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int b = find_segment(a.tile_cu, a.B, tile);
    const int t0 = (tile - __ldg(a.tile_cu + b)) * CF_TM;
    const int c0b = __ldg(a.cu + b), c1b = __ldg(a.cu + b + 1);
    const long row0 = (long)c0b * a.rate;
    const int len = (c1b - c0b) * a.rate;
    const int n0 = blockIdx.y * CF_TN;
    const int tx = tid & 15, ty = tid >> 4;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    const int ar = tid >> 2, akq = (tid & 3) * 4;      // A loader: row, k-quad
    const int bk = tid >> 4, bnq = (tid & 15) * 4;     // B loader: k, n-quad

    for (int tap = 0; tap < a.ntaps; tap++) {
        const int t = t0 + ar + a.toff[tap];
        const bool rowok = (t >= 0) && (t < len);
        const float* xrow = a.x + (row0 + t) * a.ldx + a.xcol;
        const float* wtap = a.w + (long)tap * a.cin * a.npad;
        for (int c0 = 0; c0 < a.cin; c0 += CF_KC) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rowok && (c0 + akq) < a.cin) v = *reinterpret_cast<const float4*>(xrow + c0 + akq);
            if (a.in_act) {
                v.x = leaky(v.x, a.in_slope); v.y = leaky(v.y, a.in_slope);
                v.z = leaky(v.z, a.in_slope); v.w = leaky(v.w, a.in_slope);
            }
            float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((c0 + bk) < a.cin && (n0 + bnq) < a.npad)
                wv = __ldg(reinterpret_cast<const float4*>(wtap + (long)(c0 + bk) * a.npad + n0 + bnq));
            __syncthreads();   // previous iteration's compute done before overwrite
            As[akq + 0][ar] = v.x; As[akq + 1][ar] = v.y; As[akq + 2][ar] = v.z; As[akq + 3][ar] = v.w;
            *reinterpret_cast<float4*>(&Bs[bk][bnq]) = wv;
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < CF_KC; kk++) {
                const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                const float aa[4] = {av.x, av.y, av.z, av.w};
                const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
            }
        }
    }
    // ---------------- epilogue
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int t = t0 + ty * 4 + i;
        if (t >= len) continue;
        const long row = row0 + t;
        if (a.epi == EPI_GATE) {
            // columns are interleaved (tanh-arg, sigmoid-arg) pairs (packing.py); commons.py:99-106
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                const int n = n0 + tx * 4 + j;
                if (n + 1 >= a.n) continue;
                float va = acc[i][j], vb = acc[i][j + 1];
                if (a.bias) { va += __ldg(a.bias + n); vb += __ldg(a.bias + n + 1); }
                if (a.utab) {
                    const float* ur = a.utab + (long)__ldg(a.uidx + b) * a.utab_ld;
                    va += __ldg(ur + n); vb += __ldg(ur + n + 1);
                }
                const float g = tanhf(va) * (1.f / (1.f + expf(-vb)));
                a.out[row * a.ldo + a.ocol + (n >> 1)] = g;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int n = n0 + tx * 4 + j;
                if (n < a.n) conv_epilogue_store(a, b, row, n, acc[i][j]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Embedding gather * sqrt(H)   (models.py:199)
// ------------------------------------------------------------------------------------------
__global__ void k_embed(const int* __restrict__ ids, const float* __restrict__ emb, float* __restrict__ x,
                        int rows, int H, float scale) {
    pdl_enter();
    const int q = H >> 2;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)rows * q) return;
    const int r = (int)(i / q), c = (int)(i % q) * 4;
    float4 v = __ldg(reinterpret_cast<const float4*>(emb + (long)__ldg(ids + r) * H + c));
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    *reinterpret_cast<float4*>(x + (long)r * H + c) = v;
}

// ------------------------------------------------------------------------------------------
// Relative-position multi-head self-attention (attentions.py:225-272, restated SURVEY.md A1).
// One warp per (query row, head); keys streamed 32 at a time with an online softmax;
// the +-window relative terms are a 9-wide band added in-register.
// qkv: [rows, 3H] (q | k | v).  out: [rows, H].
// ------------------------------------------------------------------------------------------
#define ATT_MAX_DKM 4   // d_k <= 128
__global__ void __launch_bounds__(128) k_rel_attention(const float* __restrict__ qkv, const float* __restrict__ Ek,
                                                       const float* __restrict__ Ev, float* __restrict__ out,
                                                       const int* __restrict__ cu, int B, int rows, int H,
                                                       int n_heads, int dk, int window) {
    pdl_enter();
    extern __shared__ float sm[];   // per warp: q[dk] + qe[2w+1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nrel = 2 * window + 1;
    float* qs = sm + warp * (dk + nrel);
    float* qe = qs + dk;
    const int row = blockIdx.x * 4 + warp;
    const int head = blockIdx.y;
    if (row >= rows) return;
    const int b = find_segment(cu, B, row);
    const int r0 = __ldg(cu + b), T = __ldg(cu + b + 1) - r0;
    const int i = row - r0;
    const int ld = 3 * H;
    const float* qrow = qkv + (long)row * ld + head * dk;
    for (int d = lane; d < dk; d += 32) qs[d] = qrow[d] / sqrtf((float)dk);   // query / sqrt(d_k), attentions.py:232
    __syncwarp();
    if (lane < nrel) {
        float s = 0.f;
        const float* e = Ek + lane * dk;
        for (int d = 0; d < dk; d++) s = fmaf(qs[d], __ldg(e + d), s);
        qe[lane] = s;
    }
    __syncwarp();
    float o[ATT_MAX_DKM];
#pragma unroll
    for (int m = 0; m < ATT_MAX_DKM; m++) o[m] = 0.f;
    float mrun = -INFINITY, lrun = 0.f;
    const float* kbase = qkv + (long)r0 * ld + H + head * dk;
    const float* vbase = qkv + (long)r0 * ld + 2 * H + head * dk;
    for (int j0 = 0; j0 < T; j0 += 32) {
        const int j = j0 + lane;
        float s = -INFINITY;
        if (j < T) {
            const float4* kr = reinterpret_cast<const float4*>(kbase + (long)j * ld);
            float acc = 0.f;
            for (int d4 = 0; d4 < (dk >> 2); d4++) {
                const float4 kv = kr[d4];
                acc = fmaf(qs[d4 * 4 + 0], kv.x, acc); acc = fmaf(qs[d4 * 4 + 1], kv.y, acc);
                acc = fmaf(qs[d4 * 4 + 2], kv.z, acc); acc = fmaf(qs[d4 * 4 + 3], kv.w, acc);
            }
            const int rel = j - i + window;
            if (rel >= 0 && rel < nrel) acc += qe[rel];
            s = acc;
        }
        const float cmax = warp_max(s);
        const float mnew = fmaxf(mrun, cmax);
        const float corr = expf(mrun - mnew);       // exp(-inf) = 0 on the first chunk
        const float p = (j < T) ? expf(s - mnew) : 0.f;
        lrun = lrun * corr + warp_sum(p);
#pragma unroll
        for (int m = 0; m < ATT_MAX_DKM; m++) o[m] *= corr;
        const int jn = min(32, T - j0);
        for (int jj = 0; jj < jn; jj++) {
            const float pj = __shfl_sync(0xffffffffu, p, jj);
            const int jg = j0 + jj;
            const float* vr = vbase + (long)jg * ld;
            const int rel = jg - i + window;
            const bool band = (rel >= 0 && rel < nrel);
#pragma unroll
            for (int m = 0; m < ATT_MAX_DKM; m++) {
                const int d = lane + 32 * m;
                if (d < dk) {
                    float vv = vr[d];
                    if (band) vv += __ldg(Ev + rel * dk + d);
                    o[m] = fmaf(pj, vv, o[m]);
                }
            }
        }
        mrun = mnew;
    }
    const float inv = 1.f / lrun;
#pragma unroll
    for (int m = 0; m < ATT_MAX_DKM; m++) {
        const int d = lane + 32 * m;
        if (d < dk) out[(long)row * H + head * dk + d] = o[m] * inv;
    }
}

// ------------------------------------------------------------------------------------------
// Channel LayerNorm family (modules.py:14-26), one warp per row, C % 32 == 0, C <= 256.
//   mode 0: out = LN(in)                                   (encoder post-LN, DP)
//   mode 1: out = GELU(LN(depthwise_k3_dilated(in)))       (DDSConv first half, modules.py:121-123)
//   mode 2: out += GELU(LN(in))                            (DDSConv second half, modules.py:125-128)
// ------------------------------------------------------------------------------------------
#define LN_MAXV 8
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

// per row of the packed batch: {position inside its utterance, utterance length}; built once per vits_prepare so that the
// per-row kernels below need no binary search over the utterance table (10 dependent L2 loads per warp, r01e profile)
__global__ void k_row_pos(const int* __restrict__ cu, int B, int rows, int2* __restrict__ out) {
    pdl_enter();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const int b = find_segment(cu, B, row);
    const int r0 = __ldg(cu + b);
    out[row] = make_int2(row - r0, __ldg(cu + b + 1) - r0);
}

// LN_RPW rows per warp: that many independent load -> reduce -> store chains in flight
template <int LN_RPW>
__global__ void __launch_bounds__(128, 10) k_layernorm(const float* __restrict__ in, float* __restrict__ out,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   int rows, int C, int mode,
                                                   const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                   int dw_k, int dw_dil, const int2* __restrict__ rowpos) {
    pdl_enter();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * 4 + warp) * LN_RPW;
    if (row0 >= rows) return;
    const int nv = C >> 5;
    float v[LN_RPW][LN_MAXV], prev[LN_RPW][LN_MAXV];
    if (mode == 1 && dw_k == 3) {
        // k = 3 (the exported voices): every load of the three taps is issued before the first FMA -- with the loads inside the
        // predicated tap loop they were serialised behind each other (r01g: 197 us per 131k rows against a 33 us HBM floor)
#pragma unroll
        for (int r = 0; r < LN_RPW; r++) {
            const int row = min(row0 + r, rows - 1);
            const int2 rp = __ldg(rowpos + row);
            const int t = rp.x, T = rp.y;
            const bool vl = t - dw_dil >= 0, vr = t + dw_dil < T;
            const float* pl = in + (long)(vl ? row - dw_dil : row) * C + lane;
            const float* pc = in + (long)row * C + lane;
            const float* pr = in + (long)(vr ? row + dw_dil : row) * C + lane;
            float xl[LN_MAXV], xc[LN_MAXV], xr[LN_MAXV];
#pragma unroll
            for (int m = 0; m < LN_MAXV; m++) if (m < nv) { xl[m] = pl[32 * m]; xc[m] = pc[32 * m]; xr[m] = pr[32 * m]; }
#pragma unroll
            for (int m = 0; m < LN_MAXV; m++) {
                if (m < nv) {
                    const int c = lane + 32 * m;
                    float acc = __ldg(dw_b + c);
                    if (vl) acc = fmaf(__ldg(dw_w + c), xl[m], acc);
                    acc = fmaf(__ldg(dw_w + C + c), xc[m], acc);
                    if (vr) acc = fmaf(__ldg(dw_w + 2 * C + c), xr[m], acc);
                    v[r][m] = acc;
                }
            }
        }
    } else if (mode == 1) {
#pragma unroll
        for (int r = 0; r < LN_RPW; r++) {
            const int row = min(row0 + r, rows - 1);
            const int2 rp = __ldg(rowpos + row);
            const int t = rp.x, T = rp.y;
#pragma unroll
            for (int m = 0; m < LN_MAXV; m++) {
                if (m < nv) {
                    const int c = lane + 32 * m;
                    float acc = __ldg(dw_b + c);
                    for (int k = 0; k < dw_k; k++) {
                        const int off = (k - dw_k / 2) * dw_dil, tt = t + off;
                        if (tt >= 0 && tt < T) acc = fmaf(__ldg(dw_w + k * C + c), in[(long)(row + off) * C + c], acc);
                    }
                    v[r][m] = acc;
                }
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < LN_RPW; r++) {
            const int row = min(row0 + r, rows - 1);
#pragma unroll
            for (int m = 0; m < LN_MAXV; m++)
                if (m < nv) {
                    v[r][m] = in[(long)row * C + lane + 32 * m];
                    if (mode == 2) prev[r][m] = out[(long)row * C + lane + 32 * m];
                }
        }
    }
    float mean[LN_RPW], rstd[LN_RPW];
#pragma unroll
    for (int r = 0; r < LN_RPW; r++) {
        float s = 0.f;
#pragma unroll
        for (int m = 0; m < LN_MAXV; m++) if (m < nv) s += v[r][m];
        mean[r] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < LN_RPW; r++) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
#pragma unroll
    for (int r = 0; r < LN_RPW; r++) {
        mean[r] /= (float)C;
        float q = 0.f;
#pragma unroll
        for (int m = 0; m < LN_MAXV; m++) if (m < nv) { const float d = v[r][m] - mean[r]; q = fmaf(d, d, q); }
        rstd[r] = q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < LN_RPW; r++) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
#pragma unroll
    for (int r = 0; r < LN_RPW; r++) rstd[r] = 1.f / sqrtf(rstd[r] / (float)C + 1e-5f);
#pragma unroll
    for (int m = 0; m < LN_MAXV; m++) {
        if (m < nv) {
            const int c = lane + 32 * m;
            const float gm = __ldg(gamma + c), bt = __ldg(beta + c);
#pragma unroll
            for (int r = 0; r < LN_RPW; r++) {
                if (row0 + r < rows) {
                    float y = (v[r][m] - mean[r]) * rstd[r] * gm + bt;
                    if (mode != 0) y = gelu_erf(y);
                    if (mode == 2) y += prev[r][m];
                    out[(long)(row0 + r) * C + c] = y;
                }
            }
        }
    }
}

// 64-bit accesses (round 2): C = 64 * NV2, lane l owns channels 64 m + 2 l + {0, 1}.  The ncu pass over the benched shape put the
// scalar kernel above at 2.8 TB/s = 0.43 of the measured HBM peak (profiles/r02f_membound.txt) -- half the load / store instructions
// for the same bytes.  Same arithmetic per element; the reduction order over lanes differs (fp32 re-association only).
template <int NV2>
__global__ void __launch_bounds__(128, 8) k_layernorm_v2(const float* __restrict__ in, float* __restrict__ out,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        int rows, int mode, const float* __restrict__ dw_w,
                                                        const float* __restrict__ dw_b, int dw_dil, const int2* __restrict__ rowpos) {
    pdl_enter();
    constexpr int C = 64 * NV2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 4 + warp;
    if (row >= rows) return;
    float2 v[NV2], prev[NV2];
    const float2* pc = reinterpret_cast<const float2*>(in + (long)row * C) + lane;
    if (mode == 1) {
        // depthwise k = 3 (modules.py:121-123) with zero padding at utterance edges; all three taps' loads before the first FMA
        const int2 rp = __ldg(rowpos + row);
        const bool vl = rp.x - dw_dil >= 0, vr = rp.x + dw_dil < rp.y;
        const float2* pl = reinterpret_cast<const float2*>(in + (long)(vl ? row - dw_dil : row) * C) + lane;
        const float2* pr = reinterpret_cast<const float2*>(in + (long)(vr ? row + dw_dil : row) * C) + lane;
        float2 xl[NV2], xc[NV2], xr[NV2];
#pragma unroll
        for (int m = 0; m < NV2; m++) { xl[m] = pl[32 * m]; xc[m] = pc[32 * m]; xr[m] = pr[32 * m]; }
#pragma unroll
        for (int m = 0; m < NV2; m++) {
            const int c2 = 32 * m + lane;
            const float2 w0 = __ldg(reinterpret_cast<const float2*>(dw_w) + c2), w1 = __ldg(reinterpret_cast<const float2*>(dw_w + C) + c2),
                         w2 = __ldg(reinterpret_cast<const float2*>(dw_w + 2 * C) + c2);
            float2 acc = __ldg(reinterpret_cast<const float2*>(dw_b) + c2);
            if (vl) { acc.x = fmaf(w0.x, xl[m].x, acc.x); acc.y = fmaf(w0.y, xl[m].y, acc.y); }
            acc.x = fmaf(w1.x, xc[m].x, acc.x); acc.y = fmaf(w1.y, xc[m].y, acc.y);
            if (vr) { acc.x = fmaf(w2.x, xr[m].x, acc.x); acc.y = fmaf(w2.y, xr[m].y, acc.y); }
            v[m] = acc;
        }
    } else {
#pragma unroll
        for (int m = 0; m < NV2; m++) {
            v[m] = pc[32 * m];
            if (mode == 2) prev[m] = reinterpret_cast<const float2*>(out + (long)row * C)[32 * m + lane];
        }
    }
    float s = 0.f;
#pragma unroll
    for (int m = 0; m < NV2; m++) s += v[m].x + v[m].y;
    s = warp_sum(s);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int m = 0; m < NV2; m++) { const float dx = v[m].x - mean, dy = v[m].y - mean; q = fmaf(dx, dx, q); q = fmaf(dy, dy, q); }
    q = warp_sum(q);
    const float rstd = 1.f / sqrtf(q / (float)C + 1e-5f);
#pragma unroll
    for (int m = 0; m < NV2; m++) {
        const int c2 = 32 * m + lane;
        const float2 gm = __ldg(reinterpret_cast<const float2*>(gamma) + c2), bt = __ldg(reinterpret_cast<const float2*>(beta) + c2);
        float2 y = make_float2((v[m].x - mean) * rstd * gm.x + bt.x, (v[m].y - mean) * rstd * gm.y + bt.y);
        if (mode != 0) { y.x = gelu_erf(y.x); y.y = gelu_erf(y.y); }
        if (mode == 2) { y.x += prev[m].x; y.y += prev[m].y; }
        reinterpret_cast<float2*>(out + (long)row * C)[c2] = y;
    }
}

// x[t, c] += tab[idx[b]][c]   (DurationPredictor speaker conditioning, models.py:153-155)
__global__ void k_add_rowbias(float* __restrict__ x, const float* __restrict__ tab, const int* __restrict__ idx,
                              const int* __restrict__ cu, int B, int rows, int C) {
    pdl_enter();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)rows * C) return;
    const int r = (int)(i / C), c = (int)(i % C);
    const int b = find_segment(cu, B, r);
    x[i] += __ldg(tab + (long)__ldg(idx + b) * C + c);
}

// ------------------------------------------------------------------------------------------
// Counter-based N(0,1): Philox4x32-10 + Box-Muller.  Stream = (seed, domain, element index).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = ((float)(a >> 8) + 0.5f) * (1.f / 16777216.f);
    const float u2 = ((float)(b >> 8) + 0.5f) * (1.f / 16777216.f);
    const float r = sqrtf(-2.f * logf(u1));
    float s, c;
    sincospif(2.f * u2, &s, &c);
    return make_float2(r * c, r * s);
}
__device__ __forceinline__ float normal_at(uint64_t seed, uint32_t domain, uint64_t idx) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)(idx >> 1), (uint32_t)(idx >> 33), domain, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float2 n = box_muller(r.x, r.y);
    return (idx & 1) ? n.y : n.x;
}

// z[ch][t] = noise * noise_w    (models.py:111); injected layout [B][2][stride]
__global__ void k_noise_dp(float* __restrict__ z0, float* __restrict__ z1, const float* __restrict__ inj,
                           long stride, const int* __restrict__ cu, int B, int rows, float noise_w,
                           uint64_t seed, uint64_t utt_base) {
    pdl_enter();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int b = find_segment(cu, B, r);
    const int t = r - __ldg(cu + b);
    float a, c;
    if (inj) {
        a = inj[((long)b * 2 + 0) * stride + t];
        c = inj[((long)b * 2 + 1) * stride + t];
    } else {
        a = normal_at(seed, 1u, ((utt_base + b) << 24) + 2ull * t);
        c = normal_at(seed, 1u, ((utt_base + b) << 24) + 2ull * t + 1);
    }
    z0[r] = a * noise_w;
    z1[r] = c * noise_w;
}

// ConvFlow.pre (Conv1d(1, F, 1)) + DDSConv conditioning add (modules.py:498, 118-119):
// h[t, c] = w[c] * x0[t] + b[c] + g[t, c]
__global__ void k_cf_pre(const float* __restrict__ x0, const float* __restrict__ w, const float* __restrict__ bias,
                         const float* __restrict__ g, float* __restrict__ h, int rows, int C) {
    pdl_enter();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)rows * C) return;
    const int r = (int)(i / C), c = (int)(i % C);
    h[i] = fmaf(__ldg(w + c), x0[r], __ldg(bias + c)) + g[i];
}

// ------------------------------------------------------------------------------------------
// Inverse rational-quadratic spline with linear tails, one thread per position
// (transforms.py:50-191).  hp: [rows, ldh] projection (3K-1 used), x1 in/out.
// ------------------------------------------------------------------------------------------
#define SPL_K 10
__global__ void k_spline_inverse(const float* __restrict__ hp, int ldh, float* __restrict__ x1, int rows,
                                 float inv_sqrt_fc) {
    pdl_enter();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float y = x1[r];
    const float Bd = 5.0f;
    if (!(y >= -Bd && y <= Bd)) return;          // identity outside the tails
    const float* h = hp + (long)r * ldh;
    float uw[SPL_K], uh[SPL_K];
    float mw = -INFINITY, mh = -INFINITY;
#pragma unroll
    for (int k = 0; k < SPL_K; k++) {
        uw[k] = h[k] * inv_sqrt_fc; uh[k] = h[SPL_K + k] * inv_sqrt_fc;
        mw = fmaxf(mw, uw[k]); mh = fmaxf(mh, uh[k]);
    }
    float sw = 0.f, sh = 0.f;
#pragma unroll
    for (int k = 0; k < SPL_K; k++) { uw[k] = expf(uw[k] - mw); uh[k] = expf(uh[k] - mh); sw += uw[k]; sh += uh[k]; }
    const float eps = 1e-3f;
    float cw[SPL_K + 1], ch[SPL_K + 1];
    cw[0] = -Bd; ch[0] = -Bd;
    float aw = 0.f, ah = 0.f;
#pragma unroll
    for (int k = 0; k < SPL_K; k++) {
        aw += eps + (1.f - eps * SPL_K) * (uw[k] / sw);
        ah += eps + (1.f - eps * SPL_K) * (uh[k] / sh);
        cw[k + 1] = 2.f * Bd * aw - Bd;
        ch[k + 1] = 2.f * Bd * ah - Bd;
    }
    cw[SPL_K] = Bd; ch[SPL_K] = Bd;
    int bin = -1;
#pragma unroll
    for (int k = 0; k <= SPL_K; k++) {
        const float loc = (k == SPL_K) ? (ch[k] + 1e-6f) : ch[k];
        bin += (y >= loc) ? 1 : 0;
    }
    bin = min(max(bin, 0), SPL_K - 1);
    float in_cw = 0.f, in_w = 0.f, in_ch = 0.f, in_h = 0.f;
#pragma unroll
    for (int k = 0; k < SPL_K; k++) {
        if (k == bin) { in_cw = cw[k]; in_w = cw[k + 1] - cw[k]; in_ch = ch[k]; in_h = ch[k + 1] - ch[k]; }
    }
    const float cst = 0.5397424172369522f;      // log(exp(1 - 1e-3) - 1), transforms.py:70
    const float ud0 = (bin == 0) ? cst : h[2 * SPL_K + bin - 1];
    const float ud1 = (bin == SPL_K - 1) ? cst : h[2 * SPL_K + bin];
    // softplus as torch (threshold 20)
    const float d0 = eps + ((ud0 > 20.f) ? ud0 : log1pf(expf(ud0)));
    const float d1 = eps + ((ud1 > 20.f) ? ud1 : log1pf(expf(ud1)));
    const float delta = in_h / in_w;
    const float u = y - in_ch;
    const float t2 = d0 + d1 - 2.f * delta;
    const float qa = u * t2 + in_h * (delta - d0);
    const float qb = in_h * d0 - u * t2;
    const float qc = -delta * u;
    const float disc = qb * qb - 4.f * qa * qc;
    const float root = (2.f * qc) / (-qb - sqrtf(disc));
    x1[r] = root * in_w + in_cw;
}

// ElementwiseAffine reverse on the logw channel (modules.py:408): logw = (z - m) * exp(-logs)
__global__ void k_ea_logw(const float* __restrict__ z, float m, float neg_logs_exp, float* __restrict__ logw, int rows) {
    pdl_enter();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) logw[r] = (z[r] - m) * neg_logs_exp;
}

// ------------------------------------------------------------------------------------------
// Length regulation, integer path (models.py:702-704, commons.py:116-129):
//   w = exp(logw) * mask * length_scale ; dur = ceil(w) ; cum = inclusive scan ; y_len = max(sum, 1)
// One block per utterance, 256 threads, chunked scan with carry.  int32 exact (fp32 cumsum of
// the reference is exact below 2^24).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_durations(const float* __restrict__ logw, float length_scale,
                                                   const int* __restrict__ cu, int* __restrict__ dur,
                                                   int* __restrict__ cum, int* __restrict__ y_len) {
    pdl_enter();
    __shared__ int wsum[8];
    __shared__ long long carry_s;      // 64-bit: 2^20 ids x 10^6 frames would wrap an int; cum / y_len saturate at INT_MAX and the host rejects
    const int b = blockIdx.x;
    const int r0 = cu[b], T = cu[b + 1] - r0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < T; base += 256) {
        const int t = base + tid;
        int d = 0;
        if (t < T) {
            const float w = (expf(logw[r0 + t]) * 1.0f) * length_scale;
            const float c = ceilf(w);
            d = (int)fminf(fmaxf(c, 0.f), 1.0e6f);
        }
        int v = d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        int woff = 0;
        for (int k = 0; k < warp; k++) woff += wsum[k];
        const long long carry = carry_s;
        const long long incl = carry + (long long)woff + (long long)v;      // a 256-id block sums to <= 2.56e8: woff + v fits an int
        if (t < T) { dur[r0 + t] = d; cum[r0 + t] = (int)min(incl, (long long)INT_MAX); }
        __syncthreads();
        if (tid == 255) carry_s = incl;
        __syncthreads();
    }
    if (tid == 0) y_len[b] = (int)min(max(carry_s, 1ll), (long long)INT_MAX);
}

// frame j of utterance b -> phoneme index: first t with cum[t] > j (searchsorted right), -1 if none
__global__ void k_frame_index(const int* __restrict__ cum, const int* __restrict__ cu_t, const int* __restrict__ cu_y,
                              int b_lo, int nB, int frames, int* __restrict__ fidx, int2* __restrict__ fpos) {
    pdl_enter();
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    const int lb = find_segment(cu_y, nB, f);
    const int j = f - __ldg(cu_y + lb);
    const int b = b_lo + lb;
    const int r0 = __ldg(cu_t + b), T = __ldg(cu_t + b + 1) - r0;
    int lo = 0, hi = T;              // first index with cum > j
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cum + r0 + mid) > j) hi = mid; else lo = mid + 1;
    }
    fidx[f] = (lo < T) ? (r0 + lo) : -1;
    if (fpos) fpos[f] = make_int2(b, j);          // {utterance, frame inside it}: consumers need no search of their own
}

// z_p[f, c] = m_p[idx, c] + eps * exp(logs_p[idx, c]) * noise_scale   (models.py:711-718)
// stats: [sumT, 2C] (m | logs).  Injected noise layout [B][C][stride].
// One thread = 4 consecutive channels of one frame (128-bit loads / stores); the frame's utterance and position come from the
// table k_frame_index wrote (the first version searched the utterance table per ELEMENT and ran one Philox per element: 410 us per
// 262144-frame chunk against a 34 us HBM floor).  The noise stream is unchanged: element idx draws from Philox counter idx >> 1.
__global__ void k_expand_sample(const float* __restrict__ stats, const int* __restrict__ fidx, const int2* __restrict__ fpos,
                                const float* __restrict__ inj, long stride, float noise_scale, uint64_t seed,
                                uint64_t utt_base, float* __restrict__ zp, int frames, int C) {
    pdl_enter();
    const int c4n = C >> 2;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)frames * c4n) return;
    const int f = (int)(i / c4n), c = (int)(i - (long)f * c4n) * 4;
    const int src = __ldg(fidx + f);
    const int2 bj = __ldg(fpos + f);
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f), lg = m;
    if (src >= 0 && stats) {          // stats == nullptr: test hook, z_p = eps * noise_scale
        m = __ldg(reinterpret_cast<const float4*>(stats + (long)src * 2 * C + c));
        lg = __ldg(reinterpret_cast<const float4*>(stats + (long)src * 2 * C + C + c));
    }
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inj) {
        const float* q = inj + ((long)bj.x * C + c) * stride + bj.y;
        e = make_float4(q[0], q[stride], q[2 * stride], q[3 * stride]);
    } else if (noise_scale != 0.f) {
        const uint64_t idx = ((utt_base + (uint64_t)bj.x) << 32) + (uint64_t)bj.y * C + c;      // even (C and c are multiples of 4)
        const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        const uint4 r0 = philox4x32_10(make_uint4((uint32_t)(idx >> 1), (uint32_t)(idx >> 33), 2u, 0u), key);
        const uint4 r1 = philox4x32_10(make_uint4((uint32_t)((idx + 2) >> 1), (uint32_t)((idx + 2) >> 33), 2u, 0u), key);
        const float2 n0 = box_muller(r0.x, r0.y), n1 = box_muller(r1.x, r1.y);
        e = make_float4(n0.x, n0.y, n1.x, n1.y);
    }
    float4 o;
    o.x = m.x + e.x * expf(lg.x) * noise_scale; o.y = m.y + e.y * expf(lg.y) * noise_scale;
    o.z = m.z + e.z * expf(lg.z) * noise_scale; o.w = m.w + e.w * expf(lg.w) * noise_scale;
    *reinterpret_cast<float4*>(zp + (long)f * C + c) = o;
}

// ------------------------------------------------------------------------------------------
// conv_post (C -> 1, k7, no bias) on leaky_relu(x, 0.01), then tanh (models.py:364-366).
// Memory-bound: reads C floats per sample once (smem tile with halo), writes 1 float.
// ------------------------------------------------------------------------------------------
#define CP_TILE 256
__global__ void __launch_bounds__(256) k_conv_post(const float* __restrict__ x, int C, const float* __restrict__ w /*[7][C]*/,
                                                   const int* __restrict__ cu, const int* __restrict__ tile_cu, int B,
                                                   int rate, float slope, float* __restrict__ audio) {
    pdl_enter();
    extern __shared__ float sx[];      // [(CP_TILE + 6)][C + 1]
    __shared__ float sw[7 * 64];
    const int tile = blockIdx.x;
    const int b = find_segment(tile_cu, B, tile);
    const int t0 = (tile - __ldg(tile_cu + b)) * CP_TILE;
    const long row0 = (long)__ldg(cu + b) * rate;
    const int len = (__ldg(cu + b + 1) - __ldg(cu + b)) * rate;
    const int ldc = C + 1;
    for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) sw[i] = __ldg(w + i);
    const int total = (CP_TILE + 6) * C;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int rr = i / C, c = i % C;
        const int t = t0 + rr - 3;
        float v = 0.f;
        if (t >= 0 && t < len) v = leaky(x[(row0 + t) * C + c], slope);
        sx[rr * ldc + c] = v;
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t < len) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const float* xr = sx + (threadIdx.x + k) * ldc;
            const float* wr = sw + k * C;
            for (int c = 0; c < C; c++) acc = fmaf(xr[c], wr[c], acc);
        }
        audio[row0 + t] = tanhf(acc);
    }
}

// ------------------------------------------------------------------------------------------
// Caller-side post-processing moved on device (voice.py:271-282, 88-91):
// per-utterance peak normalise, volume, clip, x32767 -> int16.
// ------------------------------------------------------------------------------------------
__global__ void k_absmax(const float* __restrict__ audio, const int* __restrict__ cu, int B, int rate,
                         unsigned int* __restrict__ peak_bits) {
    pdl_enter();
    // grid: (chunks, min(B, 65535)); utterances beyond the grid's y extent are covered by the loop
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const long s0 = (long)cu[b] * rate, n = (long)(cu[b + 1] - cu[b]) * rate;
        float m = 0.f;
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
            m = fmaxf(m, fabsf(audio[s0 + i]));
        m = warp_max(m);
        if ((threadIdx.x & 31) == 0) atomicMax(peak_bits + b, __float_as_uint(m));   // non-negative floats order as uints
    }
}
__global__ void k_to_int16(const float* __restrict__ audio, const int* __restrict__ cu, int B, int rate,
                           const unsigned int* __restrict__ peak_bits, int normalize, float volume,
                           int16_t* __restrict__ out) {
    pdl_enter();
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const long s0 = (long)cu[b] * rate, n = (long)(cu[b + 1] - cu[b]) * rate;
        const float peak = __uint_as_float(peak_bits[b]);
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
            float v = audio[s0 + i];
            if (normalize) v = (peak < 1e-8f) ? 0.f : (v / peak);
            if (volume != 1.f) v = v * volume;
            v = fminf(fmaxf(v, -1.f), 1.f);
            float s = v * 32767.0f;
            s = fminf(fmaxf(s, -32767.f), 32767.f);
            out[s0 + i] = (int16_t)s;      // numpy astype(int16) truncates toward zero, as does this cast
        }
    }
}
