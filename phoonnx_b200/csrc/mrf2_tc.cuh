// Fused multi-receptive-field stage of the HiFi-GAN generator, v2 (sm_100a, ResBlock2 family):
//
//     out = ( sum_r  rb_r(x) ) / n_r ,   rb_r(x) = x1 + conv_{k_r, d_r2}(lrelu(x1)) ,  x1 = x + conv_{k_r, d_r1}(lrelu(x))
//     [last stage only]  audio = tanh( conv_post( lrelu_{0.01}(out) ) )
//
// (models.py:356-366 + modules.py:355-364) in ONE kernel per stage.  Same tile geometry as mrf_tc.cuh (a CTA owns
// a window of 128*NB rows, all convs of all resblocks run on the same M=128 blocks, only the central rows are
// stored); what changed is the schedule, after ncu showed the v1 tensor pipe 11-16 % busy with every warp parked
// on an mbarrier (profiles/r01_mrf_v1_ncu.md):
//
//   * the MMA warp issues conv1 of resblock r+1 BEFORE it waits for the x1 operand of resblock r (conv1
//     accumulators are double-buffered in TMEM), so the tensor pipe always has queued work while the epilogue
//     warps turn accumulator r into the conv2 operand;
//   * weights stay resident in shared memory whenever they fit (C = 32: 61 KB), otherwise the ring is as deep as
//     shared memory allows -- v1's 2 KB-per-tap ring was bound by L2 latency, not bandwidth;
//   * raw fp32 rows are staged by a dedicated loader warp with cp.async (LDGSTS, 16 B per lane, completion on an
//     mbarrier) into a PADDED buffer (row pitch C+4 floats), so the thread-per-row residual reads are
//     bank-conflict-free (v1: 8-way), and the next tile's rows are prefetched under the current tile;
//   * the epilogues are spread over 16 warps, each thread owning 32 channels of one row (the first v2 cut, 8 warps
//     x 64 values, ran at IPC 0.9 and was epilogue-bound: profiles/r01_mrf_ncu.md);
//   * in the last stage lrelu -> conv_post -> tanh is one more tensor-core pass over the stage output, which is
//     written as a bf16 operand tile into the (idle) x1 buffer instead of going to HBM; its 128 x 16 accumulator is
//     drained one tile later, under the next tile's MMAs.
//
// Warp roles (608 threads): warps 0-15 operand conversion + epilogues (TMEM lane quadrant = warp % 4, work item =
// warp / 4 = (M block, 32-channel group)), warp 16 lane 0 = weight producer (+ TMEM allocation), warp 17 lane 0 =
// MMA issuer, warp 18 = raw-row loader.
#pragma once
#include "conv_tc.cuh"

#define MRF2_THREADS 608
#define MRF2_EPI_THREADS 512
#define MRF2_EPI_WARPS 16
#define MRF2_MAX_RB 3
#define MRF2_MAX_STAGES 32
#define MRF2_POST_K 7

struct Mrf2Args {
    const float* x;  float* out;  int C;
    int nrb;  int k[MRF2_MAX_RB];  int d1[MRF2_MAX_RB];  int d2[MRF2_MAX_RB];
    const __nv_bfloat16* w[MRF2_MAX_RB][2];  const float* b[MRF2_MAX_RB][2];
    const int* cu;  const int* tile_cu;  int B;  int rate;  int ntiles;
    const int4* tdesc;                                           // per tile {first row of the utterance, its rows, o0, -} (k_mrf2_tiles)
    float out_div;  float slope;
    const float* post_w;  float post_slope;  float* audio;      // fused conv_post (last stage) or null
    unsigned long long* dbg;                                     // test-only phase timeline of CTA 0 ([tile][48] clock64 stamps) or null
};

struct Mrf2Cfg {
    int nb;          // M=128 blocks per window
    int span;        // 128 * nb
    int hmax;        // max conv2 half receptive field
    int h1max;       // max conv1 half receptive field
    int t_out;       // span - 2*hmax: rows of stage output produced per tile
    int post_halo;   // 3 when conv_post is fused, else 0
    int t_step;      // t_out - 2*post_halo: rows a tile advances (== rows of final output per tile)
    int rx, rx1;     // rows of the two operand tiles (odd)
    int xf_pitch;    // floats per raw row (C + 4)
    int xf_bytes, x_bytes, x1_bytes;
    int slot_bytes, nstages, resident, npieces;
    int tmem_cols;
    int bias_off;
    int postw_off;   // bf16 conv_post operand [7][C/8][16][8] (column 0 = the filter)
    int smem_bytes;
};

namespace tc {
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// the mbarrier receives one arrival when all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return __uint_as_float(r);
}
}  // namespace tc

// leaky-relu for 0 < slope < 1 in two instructions
__device__ __forceinline__ float lrelu_max(float v, float slope) { return fmaxf(v, v * slope); }

#define MRF2_DBG_TILES 24
#define MRF2_STAMP(it_, slot_) do { if (dbg_on && (it_) < MRF2_DBG_TILES) a.dbg[(it_) * 48 + (slot_)] = (unsigned long long)clock64(); } while (0)

template <int C>
__global__ void __launch_bounds__(MRF2_THREADS, 1) k_mrf2_tc(const Mrf2Args a, const Mrf2Cfg c) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* sXf = reinterpret_cast<float*>(smem);             // raw fp32 rows of x, pitch C+4 (cp.async)
    uint8_t* sX = smem + c.xf_bytes;                          // lrelu(x)  bf16 K-major chunks [C/8][rx][8]
    uint8_t* sX1 = sX + c.x_bytes;                            // lrelu(x1) bf16 K-major chunks [C/8][rx1][8]; then the conv_post operand
    uint8_t* sW = sX1 + c.x1_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)c.nstages * c.slot_bytes);
    const uint32_t bar_full0 = tc::smem_u32(bars);
    const uint32_t bar_empty0 = bar_full0 + 8u * c.nstages;
    const uint32_t bar_x = bar_empty0 + 8u * c.nstages;      // X operand staged                   (512 arrivals / tile)
    const uint32_t bar_x1 = bar_x + 8u;                       // x1 operand staged, acc1 drained    (512 / resblock)
    const uint32_t bar_c1 = bar_x1 + 8u;                      // conv1 accumulators ready, 2 buffers (commit / resblock)
    const uint32_t bar_c2 = bar_c1 + 16u;                     // conv2 MMAs of one resblock complete (commit / resblock)
    const uint32_t bar_xf = bar_c2 + 8u;                      // raw rows landed                    (32 cp.async arrivals / tile)
    const uint32_t bar_xf_free = bar_xf + 8u;                 // raw rows consumed                  (512 / tile)
    const uint32_t bar_acc2_free = bar_xf_free + 8u;          // conv2 accumulators drained         (512 / tile)
    const uint32_t bar_post_rdy = bar_acc2_free + 8u;         // conv_post operand staged           (512 / tile)
    const uint32_t bar_post_done = bar_post_rdy + 8u;         // conv_post accumulators ready       (commit / tile)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * c.nstages + 10);
    float* sB = reinterpret_cast<float*>(smem + c.bias_off);  // bias1 of every resblock, then sum_r bias2_r
    uint8_t* sWp = smem + c.postw_off;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool post = a.post_w != nullptr;
    if (tid == 0) {
        for (int s = 0; s < c.nstages; s++) { tc::mbar_init(bar_full0 + 8u * s, 1); tc::mbar_init(bar_empty0 + 8u * s, 1); }
        tc::mbar_init(bar_x, MRF2_EPI_THREADS);
        tc::mbar_init(bar_x1, MRF2_EPI_THREADS);
        tc::mbar_init(bar_c1, 1);
        tc::mbar_init(bar_c1 + 8u, 1);
        tc::mbar_init(bar_c2, 1);
        tc::mbar_init(bar_xf, 32);
        tc::mbar_init(bar_xf_free, MRF2_EPI_THREADS);
        tc::mbar_init(bar_acc2_free, MRF2_EPI_THREADS);
        tc::mbar_init(bar_post_rdy, MRF2_EPI_THREADS);
        tc::mbar_init(bar_post_done, 1);
        tc::fence_mbar_init();
    }
    for (int i = tid; i < C; i += MRF2_THREADS) {
        float sum = 0.f;
        for (int r = 0; r < a.nrb; r++) { sB[r * C + i] = __ldg(a.b[r][0] + i); sum += __ldg(a.b[r][1] + i); }
        sB[a.nrb * C + i] = sum;
    }
    if (post) {
        // conv_post filter as a K-major bf16 B operand [tap][C/8][16][8]: output column 0 holds w[tap][:], columns 1-15 zero
        __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(sWp);
        for (int i = tid; i < MRF2_POST_K * C * 16; i += MRF2_THREADS) {
            const int e = i & 7, n = (i >> 3) & 15, kc = (i >> 7) % (C / 8), tap = i / (C * 16);
            wp[i] = __float2bfloat16_rn(n == 0 ? __ldg(a.post_w + tap * C + kc * 8 + e) : 0.f);
        }
        tc::fence_proxy_async();
    }
    if (warp == MRF2_EPI_WARPS) tc::tmem_alloc(tc::smem_u32(tmem_slot), (uint32_t)c.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    constexpr int KC = C / 8;                       // 16-byte chunks per row
    constexpr int NG = C / 32;                      // 32-channel groups per row
    const uint32_t lbo_x = (uint32_t)c.rx * 16u, lbo_x1 = (uint32_t)c.rx1 * 16u, lbo_w = (uint32_t)C * 16u;
    const uint32_t acc1_cols = (uint32_t)(c.nb * C);            // per conv1 buffer
    const uint32_t acc2_col = 2u * acc1_cols;
    const uint32_t accp_col = 3u * acc1_cols;                   // conv_post accumulators: 16 columns per M block
    const int lead = c.hmax + c.h1max;              // window row 0 of the X tile sits `lead` rows before the first stored row

    // tile -> (utterance rows, first stage-output row), precomputed by k_mrf2_tiles: one 16-byte load instead of a
    // binary search with ~8 dependent loads per thread per tile.  o0 may be negative by post_halo.
    auto tile_geom = [&](int tile, long& row0, int& len, int& o0) {
        const int4 d = __ldg(a.tdesc + tile);
        row0 = (long)d.x; len = d.y; o0 = d.z;
    };

    if (warp < MRF2_EPI_WARPS) {
        // ===================== operand conversion + epilogues (512 threads) =====================
        // work item of this warp: rows 128*bb + 32*q + lane, channels [32*cg, 32*cg + 32)
        const int q = warp & 3, item = warp >> 2;
        const int bb = item / NG, cg = item - bb * NG;
        const bool active = bb < c.nb;
        const int wr = 128 * bb + 32 * q + lane;                  // window row of this thread
        const float inv_div = 1.f / a.out_div;
        uint32_t n_c1[2] = {0, 0}, n_c2 = 0, n_xf = 0, n_post = 0;
        long p_row0 = 0; int p_len = 0, p_o0 = 0; bool have_prev = false;
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
        int it = 0;
        // conv_post epilogue of the previous tile: acc_post column 0 -> tanh -> audio (thread per row, 32-channel group 0)
        auto post_epilogue = [&]() {
            tc::mbar_wait(bar_post_done, n_post & 1); n_post++;
            tc::tc_fence_after();
            if (active && cg == 0) {
                const float v = tc::tmem_ld1(tmem_base + ((uint32_t)(32 * q) << 16) + accp_col + (uint32_t)(bb * 16));
                const int t = p_o0 - c.hmax + wr;                 // stage row == audio sample of this thread
                if (wr >= c.hmax + c.post_halo && wr < c.hmax + c.t_out - c.post_halo && t < p_len) a.audio[p_row0 + t] = tanhf(v);
            }
            tc::tc_fence_before();
        };
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
            long row0; int len, o0;
            tile_geom(tile, row0, len, o0);
            const int tbase = o0 - lead;
            const int tstart = tbase > 0 ? tbase : 0;          // time row held at sXf row 0
            MRF2_STAMP(it, 0);
            // ---- smem -> smem: lrelu(x) as bf16 K-major chunks; rows outside the utterance are zero padding
            tc::mbar_wait(bar_xf, n_xf & 1); n_xf++;
            MRF2_STAMP(it, 1);
            {
                const int items = c.rx * KC;
                for (int i = tid; i < items; i += MRF2_EPI_THREADS) {
                    const int r = i / KC, kc = i - r * KC;
                    const int t = tbase + r;
                    uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                    if (t >= 0 && t < len) {
                        const float4* src = reinterpret_cast<const float4*>(sXf + (size_t)(t - tstart) * c.xf_pitch + kc * 8);
                        const float4 a0 = src[0], a1 = src[1];
                        pk.x = tc::pack_bf16(lrelu_max(a0.x, a.slope), lrelu_max(a0.y, a.slope)); pk.y = tc::pack_bf16(lrelu_max(a0.z, a.slope), lrelu_max(a0.w, a.slope));
                        pk.z = tc::pack_bf16(lrelu_max(a1.x, a.slope), lrelu_max(a1.y, a.slope)); pk.w = tc::pack_bf16(lrelu_max(a1.z, a.slope), lrelu_max(a1.w, a.slope));
                    }
                    *reinterpret_cast<uint4*>(sX + ((size_t)kc * c.rx + r) * 16) = pk;
                }
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(bar_x);
            MRF2_STAMP(it, 2);
            // the previous tile's conv_post accumulators drain here, under this tile's first conv1 MMAs; this also
            // guarantees its MMAs no longer read sX1 before E1(0) below overwrites it
            if (post && have_prev) post_epilogue();
            MRF2_STAMP(it, 3);

            const int tm = o0 - c.hmax + wr;
            const bool inr = active && (tm >= 0 && tm < len);
            const float* xrow = sXf + (size_t)(inr ? (tm - tstart) : 0) * c.xf_pitch + 32 * cg;
            float xacc[32];                         // sum_r (x1_r + bias2_r) of this thread's 32 channels
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 bs = *reinterpret_cast<const float4*>(sB + a.nrb * C + 32 * cg + 4 * j);      // sum_r bias2_r
                xacc[4 * j] = bs.x; xacc[4 * j + 1] = bs.y; xacc[4 * j + 2] = bs.z; xacc[4 * j + 3] = bs.w;
            }
            const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(bb * C + 32 * cg);

            for (int r = 0; r < a.nrb; r++) {
                const uint32_t buf = (uint32_t)r & 1u;
                tc::mbar_wait(bar_c1 + 8u * buf, n_c1[buf] & 1); n_c1[buf]++;
                MRF2_STAMP(it, 4 + 4 * r);
                tc::tc_fence_after();
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; j++) pk[j] = 0u;
                if (active) {                       // warp-uniform: tcgen05.ld is .sync.aligned
                    const float* b1 = sB + r * C + 32 * cg;
#pragma unroll
                    for (int n0 = 0; n0 < 32; n0 += 16) {
                        float v[16];
                        tc::tmem_ld16(tlane + buf * acc1_cols + (uint32_t)n0, v);
#pragma unroll
                        for (int qd = 0; qd < 4; qd++) {
                            const float4 bv = *reinterpret_cast<const float4*>(b1 + n0 + 4 * qd);        // smem broadcast
                            const float4 xv = *reinterpret_cast<const float4*>(xrow + n0 + 4 * qd);      // conflict-free (pitch C+4)
                            const float x10 = v[4 * qd + 0] + bv.x + xv.x, x11 = v[4 * qd + 1] + bv.y + xv.y;
                            const float x12 = v[4 * qd + 2] + bv.z + xv.z, x13 = v[4 * qd + 3] + bv.w + xv.w;
                            xacc[n0 + 4 * qd + 0] += x10; xacc[n0 + 4 * qd + 1] += x11;
                            xacc[n0 + 4 * qd + 2] += x12; xacc[n0 + 4 * qd + 3] += x13;
                            // conv2 zero-pads x1 beyond the utterance
                            pk[(n0 >> 1) + 2 * qd + 0] = inr ? tc::pack_bf16(lrelu_max(x10, a.slope), lrelu_max(x11, a.slope)) : 0u;
                            pk[(n0 >> 1) + 2 * qd + 1] = inr ? tc::pack_bf16(lrelu_max(x12, a.slope), lrelu_max(x13, a.slope)) : 0u;
                        }
                    }
                }
                MRF2_STAMP(it, 5 + 4 * r);
                if (r == a.nrb - 1) tc::mbar_arrive(bar_xf_free);        // last read of the raw rows: the loader may prefetch
                if (r > 0) { tc::mbar_wait(bar_c2, n_c2 & 1); n_c2++; }  // conv2 of resblock r-1 no longer reads sX1
                MRF2_STAMP(it, 6 + 4 * r);
                if (active) {
                    const int row1 = wr + c.hmax;
#pragma unroll
                    for (int n8 = 0; n8 < 4; n8++)
                        *reinterpret_cast<uint4*>(sX1 + ((size_t)(4 * cg + n8) * c.rx1 + row1) * 16) =
                            make_uint4(pk[4 * n8], pk[4 * n8 + 1], pk[4 * n8 + 2], pk[4 * n8 + 3]);
                }
                tc::fence_proxy_async();
                tc::tc_fence_before();
                tc::mbar_arrive(bar_x1);
                MRF2_STAMP(it, 7 + 4 * r);
            }
            // ---- final epilogue: out = (acc2 + sum_r(x1_r + b2_r)) / n_r on the central rows
            tc::mbar_wait(bar_c2, n_c2 & 1); n_c2++;
            MRF2_STAMP(it, 16);
            tc::tc_fence_after();
            if (active) {
                const bool central = (wr >= c.hmax) && (wr < c.hmax + c.t_out);
                const bool st = central && inr;
                if (post) {
                    // lrelu_{0.01}(out) as the bf16 operand of the conv_post pass, same placement as x1 (zero outside the utterance)
                    uint32_t pk[16];
#pragma unroll
                    for (int n0 = 0; n0 < 32; n0 += 16) {
                        float v[16];
                        tc::tmem_ld16(tlane + acc2_col + (uint32_t)n0, v);
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const float o0v = (v[j] + xacc[n0 + j]) * inv_div, o1v = (v[j + 1] + xacc[n0 + j + 1]) * inv_div;
                            pk[(n0 + j) >> 1] = st ? tc::pack_bf16(lrelu_max(o0v, a.post_slope), lrelu_max(o1v, a.post_slope)) : 0u;
                        }
                    }
                    const int row1 = wr + c.hmax;
#pragma unroll
                    for (int n8 = 0; n8 < 4; n8++)
                        *reinterpret_cast<uint4*>(sX1 + ((size_t)(4 * cg + n8) * c.rx1 + row1) * 16) =
                            make_uint4(pk[4 * n8], pk[4 * n8 + 1], pk[4 * n8 + 2], pk[4 * n8 + 3]);
                } else {
                    float* orow = a.out + (row0 + tm) * C + 32 * cg;
#pragma unroll
                    for (int n0 = 0; n0 < 32; n0 += 16) {
                        float v[16];
                        tc::tmem_ld16(tlane + acc2_col + (uint32_t)n0, v);
                        if (st) {
#pragma unroll
                            for (int qd = 0; qd < 4; qd++) {
                                float4 ov;
                                ov.x = (v[4 * qd + 0] + xacc[n0 + 4 * qd + 0]) * inv_div;
                                ov.y = (v[4 * qd + 1] + xacc[n0 + 4 * qd + 1]) * inv_div;
                                ov.z = (v[4 * qd + 2] + xacc[n0 + 4 * qd + 2]) * inv_div;
                                ov.w = (v[4 * qd + 3] + xacc[n0 + 4 * qd + 3]) * inv_div;
                                *(reinterpret_cast<float4*>(orow + n0) + qd) = ov;
                            }
                        }
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(bar_acc2_free);       // the next tile's conv2 may overwrite the accumulators
            MRF2_STAMP(it, 17);
            if (post) {
                tc::fence_proxy_async();
                tc::mbar_arrive(bar_post_rdy);
                p_row0 = row0; p_len = len; p_o0 = o0; have_prev = true;
            }
        }
        if (post && have_prev) post_epilogue();
    } else if (warp == MRF2_EPI_WARPS) {
        // ===================== weight producer =====================
        if (tc::elect_one()) {
            uint32_t s = 0, ph = 1;                           // ring slot and the parity to wait for on its "empty" barrier
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                if (c.resident && tile != (int)blockIdx.x) break;
                // issue order of the MMA warp: C1(0) | C1(1) C2(0) | C1(2) C2(1) | C2(2)
                for (int step = 0; step <= a.nrb; step++)
                    for (int cv = 0; cv < 2; cv++) {
                        const int r = cv ? step - 1 : step;
                        if (r < 0 || r >= a.nrb) continue;
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
    } else if (warp == MRF2_EPI_WARPS + 1) {
        // ===================== MMA issuer =====================
        if (tc::elect_one()) {
            const uint32_t idesc = tc::make_idesc(128, C), idesc_post = tc::make_idesc(128, 16);
            const uint32_t sX_u = tc::smem_u32(sX), sX1_u = tc::smem_u32(sX1), sW_u = tc::smem_u32(sW);
            const uint64_t dhi_x = tc::make_desc(0, lbo_x, 128u), dhi_x1 = tc::make_desc(0, lbo_x1, 128u), dhi_w = tc::make_desc(0, lbo_w, 128u);
            const uint64_t dhi_wp = tc::make_desc(0, 256u, 128u);
            const uint64_t bd_step = (uint64_t)((2u * lbo_w) >> 4);
            const uint64_t ad_step_x = (uint64_t)((2u * lbo_x) >> 4), ad_step_x1 = (uint64_t)((2u * lbo_x1) >> 4);
            const uint32_t x16 = (sX_u >> 4) + (uint32_t)c.h1max, x116 = (sX1_u >> 4) + (uint32_t)c.hmax;   // 16-byte units == rows
            const uint32_t wp16 = tc::smem_u32(sWp) >> 4;
            uint32_t s = 0, ph = 0, it = 0, n_x1 = 0;           // ring slot / parity of its "full" barrier
            const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
                MRF2_STAMP((int)it, 20);
                tc::mbar_wait(bar_x, it & 1);
                MRF2_STAMP((int)it, 21);
                tc::tc_fence_after();
                for (int step = 0; step <= a.nrb; step++) {
#pragma unroll
                    for (int cv = 0; cv < 2; cv++) {
                        const int r = cv ? step - 1 : step;
                        if (r < 0 || r >= a.nrb) continue;
                        if (cv == 1) {
                            tc::mbar_wait(bar_x1, n_x1 & 1); n_x1++;                       // x1(r) staged, acc1[r&1] drained
                            if (r == 0 && it > 0) tc::mbar_wait(bar_acc2_free, (it - 1) & 1);   // previous tile's output drained
                            tc::tc_fence_after();
                        }
                        MRF2_STAMP((int)it, 22 + 2 * (2 * step + cv));
                        const int kr = a.k[r];
                        const int dil = cv ? a.d2[r] : a.d1[r];
                        const uint64_t dhi = cv ? dhi_x1 : dhi_x;
                        const uint64_t ad_step = cv ? ad_step_x1 : ad_step_x;
                        const uint32_t dcol0 = tmem_base + (cv ? acc2_col : ((uint32_t)r & 1u) * acc1_cols);
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
                        tc::umma_commit(cv ? bar_c2 : (bar_c1 + 8u * ((uint32_t)r & 1u)));
                        MRF2_STAMP((int)it, 23 + 2 * (2 * step + cv));
                    }
                }
                if (c.resident) { s = 0; }
                if (post) {
                    // conv_post: 7 taps, dilation 1, over the stage output sitting in sX1; N = 16 (column 0 is the filter)
                    tc::mbar_wait(bar_post_rdy, it & 1);
                    MRF2_STAMP((int)it, 40);
                    tc::tc_fence_after();
                    uint32_t arow16 = x116 - (uint32_t)((MRF2_POST_K - 1) >> 1);
                    for (int tap = 0; tap < MRF2_POST_K; tap++, arow16++) {
                        const uint64_t bd0 = dhi_wp | (uint64_t)((wp16 + (uint32_t)(tap * (C / 8) * 16)) & 0x3FFF);
                        for (int bb = 0; bb < c.nb; bb++) {
                            uint64_t ad = dhi_x1 | (uint64_t)((arow16 + 128u * (uint32_t)bb) & 0x3FFF);
                            uint64_t bd = bd0;
#pragma unroll
                            for (int k16 = 0; k16 < C / 16; k16++) {
                                tc::umma_bf16(tmem_base + accp_col + (uint32_t)(bb * 16), ad, bd, idesc_post, (tap > 0 || k16) ? 1u : 0u);
                                ad += ad_step_x1; bd += 32u;     // two 8-channel chunks of 16 x 16 B
                            }
                        }
                    }
                    tc::umma_commit(bar_post_done);
                    MRF2_STAMP((int)it, 41);
                }
            }
        }
    } else {
        // ===================== raw-row loader (warp 18): cp.async, 16 B per lane =====================
        uint32_t n_free = 0;
        constexpr int CH = C / 4;                    // 16-byte chunks per fp32 row
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && lane == 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
            long row0; int len, o0;
            tile_geom(tile, row0, len, o0);
            const int tbase = o0 - lead;
            const int ts = tbase > 0 ? tbase : 0;
            const int te = (tbase + c.rx < len) ? (tbase + c.rx) : len;
            MRF2_STAMP(it, 44);
            if (tile != (int)blockIdx.x) { tc::mbar_wait(bar_xf_free, n_free & 1); n_free++; }
            MRF2_STAMP(it, 45);
            const float* src0 = a.x + (row0 + ts) * C;
            const uint32_t dst0 = tc::smem_u32(sXf);
            const int total = (te - ts) * CH;
            for (int i = lane; i < total; i += 32) {
                const int r = i / CH, ch = i - r * CH;
                tc::cp_async16(dst0 + (uint32_t)(r * c.xf_pitch + 4 * ch) * 4u, src0 + (size_t)r * C + 4 * ch);
            }
            tc::cp_async_mbar_arrive(bar_xf);
            MRF2_STAMP(it, 46);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MRF2_EPI_WARPS) tc::tmem_dealloc(tmem_base, (uint32_t)c.tmem_cols);
}

// one thread per tile: utterance lookup done once, not once per thread per tile
__global__ void k_mrf2_tiles(const int* __restrict__ cu, const int* __restrict__ tile_cu, int B, int rate, int ntiles,
                             int t_step, int post_halo, int4* __restrict__ out) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= ntiles) return;
    const int b = find_segment(tile_cu, B, tile);
    const int cb0 = __ldg(cu + b), cb1 = __ldg(cu + b + 1);
    out[tile] = make_int4(cb0 * rate, (cb1 - cb0) * rate, (tile - __ldg(tile_cu + b)) * t_step - post_halo, b);
}

// ------------------------------------------------------------------------------------------ host side
static inline bool mrf2_plan(const Mrf2Args& a, Mrf2Cfg& c, int nb_pref, bool fuse_post) {
    if (a.C != 32 && a.C != 64) return false;
    if (a.nrb < 1 || a.nrb > MRF2_MAX_RB) return false;
    int hmax = 0, h1max = 0, npieces = 0;
    for (int r = 0; r < a.nrb; r++) {
        if (a.k[r] % 2 == 0 || !a.w[r][0] || !a.w[r][1]) return false;
        const int h1 = a.d1[r] * (a.k[r] - 1) / 2, h2 = a.d2[r] * (a.k[r] - 1) / 2;
        hmax = h2 > hmax ? h2 : hmax; h1max = h1 > h1max ? h1 : h1max;
        npieces += 2 * a.k[r];
    }
    c.hmax = hmax; c.h1max = h1max; c.npieces = npieces;
    c.slot_bytes = a.C * a.C * 2;
    c.post_halo = fuse_post ? (MRF2_POST_K - 1) / 2 : 0;
    if (fuse_post && hmax < c.post_halo) return false;        // the conv_post taps must stay inside the x1 tile
    c.xf_pitch = a.C + 4;
    const int limit = 225 * 1024;
    const int ng = a.C / 32;
    const int postw_bytes = fuse_post ? MRF2_POST_K * a.C * 16 * 2 : 0;
    for (int nb = nb_pref; nb >= 1; nb--) {
        if (nb == 3) continue;
        if (nb * ng > MRF2_EPI_WARPS / 4) continue;           // one (block, 32-channel group) item per epilogue warp
        if (3 * nb * a.C + (fuse_post ? nb * 16 : 0) > 512) continue;   // two conv1 buffers + conv2 (+ conv_post) accumulators
        c.nb = nb; c.span = 128 * nb; c.t_out = c.span - 2 * hmax; c.t_step = c.t_out - 2 * c.post_halo;
        if (c.t_step < 32) continue;
        c.rx = ((c.span + 2 * h1max + 7) / 8) * 8 + 1;
        c.rx1 = ((c.span + 2 * hmax + 7) / 8) * 8 + 1;
        c.xf_bytes = (c.rx * c.xf_pitch * 4 + 127) / 128 * 128;
        c.x_bytes = ((a.C / 8) * c.rx * 16 + 127) / 128 * 128;
        c.x1_bytes = ((a.C / 8) * c.rx1 * 16 + 127) / 128 * 128;
        const long fixed = (long)c.xf_bytes + c.x_bytes + c.x1_bytes;
        const long tail = (2 * MRF2_MAX_STAGES + 10) * 8 + 32 + (MRF2_MAX_RB + 1) * a.C * 4 + postw_bytes + 512;
        const long res_bytes = (long)npieces * c.slot_bytes;
        if (npieces <= MRF2_MAX_STAGES && fixed + res_bytes + tail <= limit) { c.resident = 1; c.nstages = npieces; }
        else {
            c.resident = 0;
            long room = limit - fixed - tail;
            int ns = (int)(room / c.slot_bytes);
            if (ns > MRF2_MAX_STAGES) ns = MRF2_MAX_STAGES;
            if (ns > npieces) ns = npieces;
            if (ns < 3) continue;
            c.nstages = ns;
        }
        c.bias_off = (int)((fixed + (long)c.nstages * c.slot_bytes + (2 * c.nstages + 10) * 8 + 32 + 15) / 16 * 16);
        c.postw_off = (c.bias_off + (MRF2_MAX_RB + 1) * a.C * 4 + 127) / 128 * 128;
        c.smem_bytes = c.postw_off + postw_bytes;
        if (c.smem_bytes > limit) continue;
        int cols = 32; while (cols < 3 * nb * a.C + (fuse_post ? nb * 16 : 0)) cols <<= 1;
        c.tmem_cols = cols;
        return true;
    }
    return false;
}

template <int C>
static inline cudaError_t mrf2_launch_t(const Mrf2Args& a, const Mrf2Cfg& c, int num_sms, cudaStream_t st) {
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_mrf2_tc<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_mrf2_tc<C>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    int gx = num_sms;                       // one persistent CTA per SM (TMEM: up to 512 columns each)
    if (gx > a.ntiles) gx = a.ntiles;
    if (gx < 1) return cudaSuccess;
    k_mrf2_tiles<<<(a.ntiles + 255) / 256, 256, 0, st>>>(a.cu, a.tile_cu, a.B, a.rate, a.ntiles, c.t_step, c.post_halo,
                                                         const_cast<int4*>(a.tdesc));
    k_mrf2_tc<C><<<gx, MRF2_THREADS, c.smem_bytes, st>>>(a, c);
    return cudaGetLastError();
}

static inline cudaError_t mrf2_launch(const Mrf2Args& a, const Mrf2Cfg& c, int num_sms, cudaStream_t st) {
    if (a.C == 32) return mrf2_launch_t<32>(a, c, num_sms, st);
    if (a.C == 64) return mrf2_launch_t<64>(a, c, num_sms, st);
    return cudaErrorInvalidConfiguration;
}
