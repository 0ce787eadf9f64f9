// Fused multi-receptive-field stage of the HiFi-GAN generator for sm_100a (ResBlock2 family):
//
//     out = ( sum_r  rb_r(x) ) / n_r ,   rb_r(x) = x1 + conv_{k_r, d_r2}(lrelu(x1)) ,  x1 = x + conv_{k_r, d_r1}(lrelu(x))
//
// (models.py:356-363 + modules.py:355-364) in ONE kernel per stage: the upsampled activation tile
// is read from HBM once, every intermediate (lrelu'd bf16 operands, x1, the per-resblock sums)
// stays in shared memory / TMEM / registers, and the stage output is written once.  The unfused
// path moves ~20 tensor passes through HBM per stage; this one moves ~2.4.
//
// Tile geometry: a CTA owns a window of S = 128*NB rows; both convs of every resblock are computed
// on the same NB M=128 blocks (so a thread's TMEM lane holds the same time row in conv1, conv2 and
// for every resblock -> x1 and the running sum never leave registers, and conv2 accumulates across
// resblocks directly in TMEM).  Only the central T_out = S - 2*HMAX rows are stored
// (HMAX = largest conv2 half-receptive-field); the rim rows are the halo recompute.
//
// Warp roles (320 threads): warps 0-7 load / epilogue (lane quadrant = warp % 4, block = warp / 4),
// warp 8 lane 0 = weight producer (cp.async.bulk ring or resident), warp 9 lane 0 = MMA issuer.
#pragma once
#include "conv_tc.cuh"

#define MRF_THREADS 320
#define MRF_EPI_THREADS 256
#define MRF_MAX_RB 3
#define MRF_MAX_STAGES 32

struct MrfArgs {
    const float* x;  float* out;  int C;
    int nrb;  int k[MRF_MAX_RB];  int d1[MRF_MAX_RB];  int d2[MRF_MAX_RB];
    const __nv_bfloat16* w[MRF_MAX_RB][2];  const float* b[MRF_MAX_RB][2];
    const int* cu;  const int* tile_cu;  int B;  int rate;  int ntiles;
    float out_div;  float slope;
};

struct MrfCfg {
    int nb;          // M=128 blocks per window
    int span;        // 128 * nb
    int hmax;        // max conv2 half receptive field
    int h1max;       // max conv1 half receptive field
    int t_out;       // span - 2*hmax
    int rx, rx1;     // rows of the two operand tiles (odd)
    int xf_bytes, x_bytes, x1_bytes;
    int slot_bytes, nstages, resident, npieces;
    int tmem_cols;
    int bias_off;
    int smem_bytes;
};

template <int C, int OWN>
__global__ void __launch_bounds__(MRF_THREADS, (C == 32 && OWN == 1) ? 2 : 1) k_mrf_tc(const MrfArgs a, const MrfCfg c) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* sXf = reinterpret_cast<float*>(smem);             // raw fp32 rows of x (bulk-copied; also the conv1 residual)
    uint8_t* sX = smem + c.xf_bytes;
    uint8_t* sX1 = sX + c.x_bytes;
    uint8_t* sW = sX1 + c.x1_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)c.nstages * c.slot_bytes);
    const uint32_t bar_full0 = tc::smem_u32(bars);
    const uint32_t bar_empty0 = bar_full0 + 8u * c.nstages;
    const uint32_t bar_x = bar_empty0 + 8u * c.nstages;     // X tile staged            (256 arrivals / tile)
    const uint32_t bar_x1 = bar_x + 8u;                      // x1 operand tile staged   (256 arrivals / resblock)
    const uint32_t bar_c1 = bar_x1 + 8u;                     // conv1 accumulators ready (commit / resblock)
    const uint32_t bar_c2 = bar_c1 + 8u;                     // conv2 accumulators ready (commit / tile)
    const uint32_t bar_xf = bar_c2 + 8u;                     // raw x rows landed (bulk copy tx / tile)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * c.nstages + 5);
    float* sB = reinterpret_cast<float*>(smem + c.bias_off);   // bias1 of every resblock, then sum_r bias2_r

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < c.nstages; s++) { tc::mbar_init(bar_full0 + 8u * s, 1); tc::mbar_init(bar_empty0 + 8u * s, 1); }
        tc::mbar_init(bar_x, MRF_EPI_THREADS);
        tc::mbar_init(bar_x1, MRF_EPI_THREADS);
        tc::mbar_init(bar_c1, 1);
        tc::mbar_init(bar_c2, 1);
        tc::mbar_init(bar_xf, 1);
        tc::fence_mbar_init();
    }
    for (int i = tid; i < C; i += MRF_THREADS) {
        float sum = 0.f;
        for (int r = 0; r < a.nrb; r++) { sB[r * C + i] = __ldg(a.b[r][0] + i); sum += __ldg(a.b[r][1] + i); }
        sB[a.nrb * C + i] = sum;
    }
    if (warp == 8) tc::tmem_alloc(tc::smem_u32(tmem_slot), (uint32_t)c.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    constexpr int KC = C / 8;                       // 16-byte chunks per row
    const uint32_t lbo_x = (uint32_t)c.rx * 16u, lbo_x1 = (uint32_t)c.rx1 * 16u, lbo_w = (uint32_t)C * 16u;
    const uint32_t acc2_col = (uint32_t)(c.nb * C);

    if (warp < 8) {
        // ===================== loader + epilogues (256 threads) =====================
        const int q = warp & 3, hb = warp >> 2;
        uint32_t ph_c1 = 0, ph_c2 = 0, ph_xf = 0;
        // raw rows [max(tbase,0), min(tbase+rx,len)) of the tile's utterance: one contiguous bulk copy
        auto issue_x_copy = [&](int tile) {
            const int b = find_segment(a.tile_cu, a.B, tile);
            const int o0 = (tile - __ldg(a.tile_cu + b)) * c.t_out;
            const int cb0 = __ldg(a.cu + b), cb1 = __ldg(a.cu + b + 1);
            const long row0 = (long)cb0 * a.rate;
            const int len = (cb1 - cb0) * a.rate;
            const int tbase = o0 - c.hmax - c.h1max;
            const int ts = tbase > 0 ? tbase : 0;
            const int te = (tbase + c.rx < len) ? (tbase + c.rx) : len;
            const uint32_t bytes = (uint32_t)(te - ts) * (uint32_t)(C * 4);
            tc::mbar_expect_tx(bar_xf, bytes);
            tc::bulk_g2s(tc::smem_u32(sXf), a.x + (row0 + ts) * C, bytes, bar_xf);
        };
        if (tid == 0 && (int)blockIdx.x < a.ntiles) issue_x_copy(blockIdx.x);
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            const int b = find_segment(a.tile_cu, a.B, tile);
            const int o0 = (tile - __ldg(a.tile_cu + b)) * c.t_out;
            const int cb0 = __ldg(a.cu + b), cb1 = __ldg(a.cu + b + 1);
            const long row0 = (long)cb0 * a.rate;
            const int len = (cb1 - cb0) * a.rate;
            const int tbase = o0 - c.hmax - c.h1max;
            const int tstart = tbase > 0 ? tbase : 0;          // time row held at sXf row 0
            // ---- smem -> smem: lrelu(x) as bf16 K-major chunks; rows outside the utterance are zero padding
            tc::mbar_wait(bar_xf, ph_xf & 1); ph_xf++;
            {
                const int items = c.rx * KC;
                for (int i = tid; i < items; i += MRF_EPI_THREADS) {
                    const int r = i / KC, kc = i - r * KC;
                    const int t = tbase + r;
                    uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                    if (t >= 0 && t < len) {
                        const float4* src = reinterpret_cast<const float4*>(sXf + (size_t)(t - tstart) * C + kc * 8);
                        const float4 a0 = src[0], a1 = src[1];
                        pk.x = tc::pack_bf16(leaky(a0.x, a.slope), leaky(a0.y, a.slope)); pk.y = tc::pack_bf16(leaky(a0.z, a.slope), leaky(a0.w, a.slope));
                        pk.z = tc::pack_bf16(leaky(a1.x, a.slope), leaky(a1.y, a.slope)); pk.w = tc::pack_bf16(leaky(a1.z, a.slope), leaky(a1.w, a.slope));
                    }
                    *reinterpret_cast<uint4*>(sX + ((size_t)kc * c.rx + r) * 16) = pk;
                }
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(bar_x);

            // this thread's rows: block bb (= hb, hb+2, ...), window row wr = 128*bb + 32*q + lane.
            // The fp32 residual row is pulled into registers once per tile (it is the same for every resblock),
            // after which the raw staging buffer is free and the NEXT tile's rows are prefetched under this tile.
            float xrow[OWN][C];
            bool inr[OWN];
            float xacc[OWN][C];                     // sum_r (x1_r + bias2_r) for the OWN blocks this thread owns
#pragma unroll
            for (int o = 0; o < OWN; o++) {
                const int bb = hb + 2 * o;
                const int tm = o0 - c.hmax + 128 * bb + 32 * q + lane;
                inr[o] = (bb < c.nb) && (tm >= 0 && tm < len);
                const float4* xr = reinterpret_cast<const float4*>(sXf + (size_t)(inr[o] ? (tm - tstart) : 0) * C);
#pragma unroll
                for (int j = 0; j < C / 4; j++) {
                    const float4 xv = inr[o] ? xr[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                    xrow[o][4 * j] = xv.x; xrow[o][4 * j + 1] = xv.y; xrow[o][4 * j + 2] = xv.z; xrow[o][4 * j + 3] = xv.w;
                    const float4 bs = *reinterpret_cast<const float4*>(sB + a.nrb * C + 4 * j);      // sum_r bias2_r
                    xacc[o][4 * j] = bs.x; xacc[o][4 * j + 1] = bs.y; xacc[o][4 * j + 2] = bs.z; xacc[o][4 * j + 3] = bs.w;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(MRF_EPI_THREADS) : "memory");
            if (tid == 0 && tile + (int)gridDim.x < a.ntiles) { tc::fence_proxy_async(); issue_x_copy(tile + gridDim.x); }

            for (int r = 0; r < a.nrb; r++) {
                tc::mbar_wait(bar_c1, ph_c1 & 1); ph_c1++;
                tc::tc_fence_after();
                const float* b1 = sB + r * C;
#pragma unroll
                for (int o = 0; o < OWN; o++) {
                    const int bb = hb + 2 * o;
                    if (bb >= c.nb) break;
                    const int wr = 128 * bb + 32 * q + lane;
                    const uint32_t trow = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(bb * C);
#pragma unroll
                    for (int n0 = 0; n0 < C; n0 += 16) {
                        float v[16];
                        tc::tmem_ld16(trow + (uint32_t)n0, v);
                        uint32_t pk[8];
#pragma unroll
                        for (int qd = 0; qd < 4; qd++) {
                            const float4 bv = *reinterpret_cast<const float4*>(b1 + n0 + 4 * qd);      // smem broadcast
                            float x1[4];
                            x1[0] = v[4 * qd + 0] + bv.x + xrow[o][n0 + 4 * qd + 0]; x1[1] = v[4 * qd + 1] + bv.y + xrow[o][n0 + 4 * qd + 1];
                            x1[2] = v[4 * qd + 2] + bv.z + xrow[o][n0 + 4 * qd + 2]; x1[3] = v[4 * qd + 3] + bv.w + xrow[o][n0 + 4 * qd + 3];
                            if (!inr[o]) { x1[0] = x1[1] = x1[2] = x1[3] = 0.f; }   // conv2 zero-pads x1 beyond the utterance
                            xacc[o][n0 + 4 * qd + 0] += x1[0]; xacc[o][n0 + 4 * qd + 1] += x1[1];
                            xacc[o][n0 + 4 * qd + 2] += x1[2]; xacc[o][n0 + 4 * qd + 3] += x1[3];
                            pk[2 * qd + 0] = tc::pack_bf16(leaky(x1[0], a.slope), leaky(x1[1], a.slope));
                            pk[2 * qd + 1] = tc::pack_bf16(leaky(x1[2], a.slope), leaky(x1[3], a.slope));
                        }
                        const int row1 = wr + c.hmax;
                        uint8_t* dst = sX1 + ((size_t)(n0 >> 3) * c.rx1 + row1) * 16;
                        *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(dst + (size_t)c.rx1 * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
                tc::fence_proxy_async();
                tc::tc_fence_before();
                tc::mbar_arrive(bar_x1);
            }
            // ---- final epilogue: out = (acc2 + sum_r(x1_r + b2_r)) / n_r on the central rows
            tc::mbar_wait(bar_c2, ph_c2 & 1); ph_c2++;
            tc::tc_fence_after();
#pragma unroll
            for (int o = 0; o < OWN; o++) {
                const int bb = hb + 2 * o;
                if (bb >= c.nb) break;
                const int wr = 128 * bb + 32 * q + lane;
                const int tm = o0 - c.hmax + wr;
                const bool st = (wr >= c.hmax) && (wr < c.hmax + c.t_out) && (tm < len);
                float* orow = a.out + (row0 + tm) * C;
                const uint32_t trow = tmem_base + ((uint32_t)(32 * q) << 16) + acc2_col + (uint32_t)(bb * C);
#pragma unroll
                for (int n0 = 0; n0 < C; n0 += 16) {
                    float v[16];
                    tc::tmem_ld16(trow + (uint32_t)n0, v);
                    if (st) {
#pragma unroll
                        for (int qd = 0; qd < 4; qd++) {
                            float4 ov;
                            ov.x = (v[4 * qd + 0] + xacc[o][n0 + 4 * qd + 0]) / a.out_div;
                            ov.y = (v[4 * qd + 1] + xacc[o][n0 + 4 * qd + 1]) / a.out_div;
                            ov.z = (v[4 * qd + 2] + xacc[o][n0 + 4 * qd + 2]) / a.out_div;
                            ov.w = (v[4 * qd + 3] + xacc[o][n0 + 4 * qd + 3]) / a.out_div;
                            *(reinterpret_cast<float4*>(orow + n0) + qd) = ov;
                        }
                    }
                }
            }
            tc::tc_fence_before();     // TMEM reads retire before the next tile's bar_x arrive releases the MMA warp
        }
    } else if (warp == 8) {
        // ===================== weight producer =====================
        if (lane == 0) {
            uint32_t s = 0, ph = 1;                           // ring slot and the parity to wait for on its "empty" barrier
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                if (c.resident && tile != (int)blockIdx.x) break;
                for (int r = 0; r < a.nrb; r++)
                    for (int cv = 0; cv < 2; cv++) {
                        const __nv_bfloat16* wsrc = a.w[r][cv];
                        for (int tap = 0; tap < a.k[r]; tap++) {
                            if (!c.resident) tc::mbar_wait(bar_empty0 + 8u * s, ph);
                            const uint32_t fb = bar_full0 + 8u * s;
                            tc::mbar_expect_tx(fb, (uint32_t)c.slot_bytes);
                            tc::bulk_g2s(tc::smem_u32(sW) + s * (uint32_t)c.slot_bytes, wsrc, (uint32_t)c.slot_bytes, fb);
                            wsrc += C * C;
                            if (++s == (uint32_t)c.nstages) { s = 0; ph ^= 1u; }
                        }
                    }
            }
        }
    } else {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc(128, C);
            const uint32_t sX_u = tc::smem_u32(sX), sX1_u = tc::smem_u32(sX1), sW_u = tc::smem_u32(sW);
            const uint64_t dhi_x = tc::make_desc(0, lbo_x, 128u), dhi_x1 = tc::make_desc(0, lbo_x1, 128u), dhi_w = tc::make_desc(0, lbo_w, 128u);
            const uint64_t bd_step = (uint64_t)((2u * lbo_w) >> 4);
            const uint64_t ad_step_x = (uint64_t)((2u * lbo_x) >> 4), ad_step_x1 = (uint64_t)((2u * lbo_x1) >> 4);
            const uint32_t x16 = (sX_u >> 4) + (uint32_t)c.h1max, x116 = (sX1_u >> 4) + (uint32_t)c.hmax;   // 16-byte units == rows
            uint32_t s = 0, ph = 0, it = 0, ph_x1 = 0;          // ring slot / parity of its "full" barrier
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
                tc::mbar_wait(bar_x, it & 1);
                tc::tc_fence_after();
                for (int r = 0; r < a.nrb; r++) {
                    const int kr = a.k[r];
#pragma unroll
                    for (int cv = 0; cv < 2; cv++) {
                        if (cv == 1) { tc::mbar_wait(bar_x1, ph_x1 & 1); ph_x1++; tc::tc_fence_after(); }
                        const int dil = cv ? a.d2[r] : a.d1[r];
                        const uint64_t dhi = cv ? dhi_x1 : dhi_x;
                        const uint64_t ad_step = cv ? ad_step_x1 : ad_step_x;
                        const uint32_t dcol0 = tmem_base + (cv ? acc2_col : 0u);
                        uint32_t arow16 = (cv ? x116 : x16) - (uint32_t)(((kr - 1) >> 1) * dil);     // tap 0
                        for (int tap = 0; tap < kr; tap++, arow16 += (uint32_t)dil) {
                            if (!c.resident || it == 0) { tc::mbar_wait(bar_full0 + 8u * s, ph); tc::tc_fence_after(); }
                            const uint64_t bd0 = dhi_w | (uint64_t)(((sW_u + s * (uint32_t)c.slot_bytes) >> 4) & 0x3FFF);
                            // conv1: fresh accumulator per resblock; conv2 accumulates across resblocks (and taps)
                            const uint32_t acc0 = (tap > 0 || (cv && r > 0)) ? 1u : 0u;
                            for (int bb = 0; bb < c.nb; bb++) {
                                uint64_t ad = dhi | (uint64_t)((arow16 + 128u * (uint32_t)bb) & 0x3FFF);
                                uint64_t bd = bd0;
                                const uint32_t dcol = dcol0 + (uint32_t)(bb * C);
#pragma unroll
                                for (int k16 = 0; k16 < C / 16; k16++) {
                                    tc::umma_bf16(dcol, ad, bd, idesc, k16 ? 1u : acc0);
                                    ad += ad_step; bd += bd_step;
                                }
                            }
                            if (!c.resident) tc::umma_commit(bar_empty0 + 8u * s);
                            if (++s == (uint32_t)c.nstages) { s = 0; ph ^= 1u; }
                        }
                        if (cv == 0) tc::umma_commit(bar_c1);
                    }
                }
                if (c.resident) { s = 0; }
                tc::umma_commit(bar_c2);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem_base, (uint32_t)c.tmem_cols);
}

// ------------------------------------------------------------------------------------------ host side
static inline bool mrf_tc_plan(const MrfArgs& a, MrfCfg& c, int nb_pref) {
    if (a.C != 32 && a.C != 64) return false;
    if (a.nrb < 1 || a.nrb > MRF_MAX_RB) return false;
    int hmax = 0, h1max = 0, npieces = 0;
    for (int r = 0; r < a.nrb; r++) {
        if (a.k[r] % 2 == 0 || !a.w[r][0] || !a.w[r][1]) return false;
        const int h1 = a.d1[r] * (a.k[r] - 1) / 2, h2 = a.d2[r] * (a.k[r] - 1) / 2;
        hmax = h2 > hmax ? h2 : hmax; h1max = h1 > h1max ? h1 : h1max;
        npieces += 2 * a.k[r];
    }
    c.hmax = hmax; c.h1max = h1max; c.npieces = npieces;
    c.slot_bytes = a.C * a.C * 2;
    for (int nb = nb_pref; nb >= 1; nb--) {
        if (nb == 3) continue;
        if (2 * nb * a.C > 512) continue;
        c.nb = nb; c.span = 128 * nb; c.t_out = c.span - 2 * hmax;
        if (c.t_out < 32) continue;
        c.rx = ((c.span + 2 * h1max + 7) / 8) * 8 + 1;
        c.rx1 = ((c.span + 2 * hmax + 7) / 8) * 8 + 1;
        c.xf_bytes = (c.rx * a.C * 4 + 127) / 128 * 128;
        c.x_bytes = ((a.C / 8) * c.rx * 16 + 127) / 128 * 128;
        c.x1_bytes = ((a.C / 8) * c.rx1 * 16 + 127) / 128 * 128;
        const long res_bytes = (long)npieces * c.slot_bytes;
        // weights stay resident only when that does not cost a second CTA per SM
        const long fixed = (long)c.xf_bytes + c.x_bytes + c.x1_bytes + 2048;
        if (npieces <= MRF_MAX_STAGES && fixed + res_bytes <= 112 * 1024) { c.resident = 1; c.nstages = npieces; }
        else { c.resident = 0; c.nstages = (c.slot_bytes <= 4096) ? 8 : 6; }
        c.bias_off = (c.xf_bytes + c.x_bytes + c.x1_bytes + c.nstages * c.slot_bytes + (2 * c.nstages + 5) * 8 + 16 + 15) / 16 * 16;
        c.smem_bytes = c.bias_off + (MRF_MAX_RB + 1) * a.C * 4;
        if (c.smem_bytes > 225 * 1024) continue;
        int cols = 32; while (cols < 2 * nb * a.C) cols <<= 1;
        c.tmem_cols = cols;
        return true;
    }
    return false;
}

template <int C, int OWN>
static inline cudaError_t mrf_tc_launch_t(const MrfArgs& a, const MrfCfg& c, int num_sms, cudaStream_t st) {
    static bool attr_set[64] = {false};
    static int regs = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_mrf_tc<C, OWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_mrf_tc<C, OWN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    if (!regs) {
        cudaFuncAttributes fa;
        regs = (cudaFuncGetAttributes(&fa, k_mrf_tc<C, OWN>) == cudaSuccess && fa.numRegs > 0) ? fa.numRegs : 128;
    }
    int occ = (228 * 1024) / (c.smem_bytes + 1024);
    const int reg_occ = 65536 / (((regs + 7) / 8 * 8) * MRF_THREADS);
    if (occ > reg_occ) occ = reg_occ;
    const int tmem_occ = 512 / c.tmem_cols;
    if (occ > tmem_occ) occ = tmem_occ;
    if (occ < 1) occ = 1;
    int gx = num_sms * occ;
    if (gx > a.ntiles) gx = a.ntiles;
    if (gx < 1) return cudaSuccess;
    k_mrf_tc<C, OWN><<<gx, MRF_THREADS, c.smem_bytes, st>>>(a, c);
    return cudaGetLastError();
}

static inline cudaError_t mrf_tc_launch(const MrfArgs& a, const MrfCfg& c, int num_sms, cudaStream_t st) {
    if (c.nb > 4) return cudaErrorInvalidConfiguration;
    if (a.C == 32) return c.nb <= 2 ? mrf_tc_launch_t<32, 1>(a, c, num_sms, st) : mrf_tc_launch_t<32, 2>(a, c, num_sms, st);
    if (a.C == 64) return c.nb <= 2 ? mrf_tc_launch_t<64, 1>(a, c, num_sms, st) : mrf_tc_launch_t<64, 2>(a, c, num_sms, st);
    return cudaErrorInvalidConfiguration;
}
