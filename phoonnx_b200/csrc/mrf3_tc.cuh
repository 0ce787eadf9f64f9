// Fused multi-receptive-field stage of the HiFi-GAN generator, v3 (sm_100a, ResBlock2 family):
//
//     [optional]  x = ConvTranspose1d_u( h )                                       (models.py:320-332, h already lrelu'd)
//     out = ( sum_r  rb_r(x) ) / n_r ,   rb_r(x) = x1 + conv_{k_r, d_r2}(lrelu(x1)) ,  x1 = x + conv_{k_r, d_r1}(lrelu(x))
//     [last stage only]  audio = tanh( conv_post( lrelu_{0.01}(out) ) )             (models.py:356-366 + modules.py:355-364)
//
// in ONE kernel per stage.  The same kernel with n_r = 1 and `rb1` set is one (conv_{k,d} -> conv_{k,1}) PAIR of a ResBlock1
// (modules.py:301-314: xt = c1(lrelu x); xt = c2(lrelu xt); x = xt + x): the only differences are that the first convolution carries no
// residual and the second adds x instead of x1 -- the `high` preset's 128-, 64- and 32-channel stages, round 2.
// Tile geometry: a CTA owns a window of 128*NB rows, all convs of all resblocks run on the same M=128 blocks, only the central rows
// are stored.  What v3 changes, after the r01 phase
// timeline of v2 (profiles/r01b_mma_probe_and_mrf2_timeline.log: tensor pipe busy 11.8k of a 22.4k-cycle tile, the
// rest a serial C1(r) -> E1(r) -> C2(r) chain; and a separate polyphase ConvTranspose kernel that wrote and re-read
// 64 KB of fp32 per frame at 2.6 TB/s):
//
//   * inter-stage activations travel as bf16( lrelu_{0.1}(.) ) -- exactly the MMA operand the consumer needs.  A
//     dedicated loader warp cp.async's rows straight into the K-major operand tile (zero-filled outside the
//     utterance): no fp32 staging buffer (76 KB), no conversion pass;
//   * the residual x is recovered from that operand (lrelu is invertible: x = min(v, v/slope)); costs 1.3 dB of SNR
//     on the medium voice (60.4 -> 59.1 dB vs the fp32 oracle, gate 40 dB), frees the shared memory that makes the
//     next two points possible;
//   * in the last stage the ConvTranspose1d (u = 4) is one more tensor-core pass at the head of the tile: the input
//     tile [128*NUB + 2 rows, C_in] is the A operand, the polyphase weights [2 taps][C_in][u/2 * C] the B operands,
//     accumulator column (p, c) of input row q is x[u*q + p, c]; 16 epilogue warps (quadrant x phase) turn it into
//     the lrelu'd operand tile.  x never exists in HBM;
//   * one conv1 accumulator per resblock (TMEM: (n_r + 1) * NB * C = 512 columns; conv_post reuses the last conv1
//     buffer) and the issue order C1(0) post(previous tile) C1(1) C1(2) | C2(0) C2(1) ups(next tile) C2(2): every C1 is queued
//     before the first epilogue result is needed, so E1(r) overlaps C1(r+1..) and C2(r-1); the previous tile's conv_post goes in
//     behind this tile's first conv1 instead of idling the tensor pipe at the tile's end.
//
//   * (round 2) the input tile arrives by TMA: the K-major no-swizzle operand plane of 8 channels IS a [rows x 16 bytes] box of the
//     2-D tensor [rows, C] of bf16 operand rows, so C/8 x ceil(rows/256) cp.async.bulk.tensor (UTMALDG) issued by one lane replace
//     ~2.2k LDGSTS of the 32-lane cp.async loader; rows outside the ARRAY come back as zeros from the TMA unit, rows inside the array
//     but outside the tile's utterance (first / last tile of an utterance only) are zeroed by the loader warp afterwards.  The
//     cp.async loader stays as option `mrf_tma` = 0 (A/B, and the reference for the parity test of the two loaders).
//   * two MMA-issuing warps, each owning half of the M blocks (distinct accumulators, so the result does not depend on
//     how their instructions interleave): one thread's descriptor set-up (R2UR moves, uniform adds) does not overlap
//     the execution of its own MMAs -- measured 60-75 cycles per N=32 MMA issued against 40 executed -- but it does
//     overlap the other warp's.
//
// C = 128 (ResBlock1 pairs only): one M block per tile, 16 epilogue warps = 4 lane quadrants x 4 column groups, weights through the ring.
// Warp roles (640 threads): warps 0-15 epilogues (TMEM lane quadrant = warp % 4, work item = warp / 4), warp 16
// elected lane = weight producer (+ TMEM allocation), warps 17 and 19 elected lane = MMA issuers, warp 18 = input-row loader.
#pragma once
#include <cuda.h>          // CUtensorMap (the type only: the encoder is fetched at run time, engine.cu)
#include "mrf_tiles.cuh"

#define MRF3_THREADS 640
#define MRF3_EPI_THREADS 512
#define MRF3_EPI_WARPS 16
#define MRF3_MAX_RB 3
#define MRF3_MAX_STAGES 32
#define MRF3_POST_K 7
#define MRF3_NBAR 16
#define MRF3_TMA_BOX 256          // rows per TMA box (hardware limit of a box dimension)
#define MRF3_DBG_TILES 24

struct Mrf3Args {
    // input: mode A (up_u == 0): xb = bf16 lrelu(x) [rows, C];  mode U (up_u == 4): hb = bf16 lrelu(h) [rows / u, up_cin]
    const __nv_bfloat16* xb;
    const __nv_bfloat16* hb;  int up_u;  int up_cin;
    const __nv_bfloat16* up_w[2];  const float* up_b;      // polyphase halves A (taps -1, 0) and B (taps 0, +1): [2][up_cin/8][u/2*C][8]
    int C;
    int nrb;  int k[MRF3_MAX_RB];  int d1[MRF3_MAX_RB];  int d2[MRF3_MAX_RB];
    const __nv_bfloat16* w[MRF3_MAX_RB][2];  const float* b[MRF3_MAX_RB][2];
    const int* cu;  const int* tile_cu;  int B;  int rate;  int ntiles;
    long in_rows;                                                // host side: rows of the input array (xb, or hb in mode U), for its tensor map
    const int4* tdesc;                                           // per tile {first row of the utterance, its rows, o0, -}
    float out_div;  float slope;  int interleave;                // issue order of the convs (mrf3_step)
    int rb1;                                                     // ResBlock1 pair (modules.py:301-314): t = conv1(lrelu x), out = x + conv2(lrelu t)  [n_r must be 1]
    int accumulate;                                              // fp32 `out` only: out = (out + result) / out_div (per-resblock results summed in place)
    const __nv_bfloat16* addb[2];  int naddb;                    // up to two more row arrays [rows, C] (plain bf16 values) added to the result before out_div:
                                                                 // the other resblocks' results when the LAST resblock's last pair finishes a ResBlock1 stage
    float* out;                                                  // fp32 [rows, C]                       (or null)
    __nv_bfloat16* outb;  float outb_slope;                      // bf16 lrelu_{outb_slope}(out) [rows, C] (or null)
    const float* post_w;  float post_slope;  float* audio;      // fused conv_post (last stage) or null
    unsigned long long* dbg;
};

struct Mrf3Cfg {
    int nb, span, hmax, h1max, t_out, post_halo, t_step;
    int rx, rx1;                 // rows of the two operand tiles (odd)
    int x_bytes, x1_bytes;
    int slot_bytes, nstages, resident, npieces;
    int tmem_cols;
    int nmw;                     // MMA-issuing warps (2 when the M blocks split evenly)
    int nub, u_rows, u_bytes, upw_bytes;   // mode U: M blocks of the ups pass, rows / bytes of its input tile, bytes of both weight halves
    int bias_off, postw_off, sp_off;   // sp: per-row per-tap partial sums of conv_post, [span + 8][MRF3_SP_PITCH] floats
    int smem_bytes;
    int tma, nboxes, box_rows;   // input tile by TMA: per 8-channel plane `nboxes` boxes of [box_rows rows x 16 bytes] (rows == nboxes * box_rows)
};

namespace tc {
// 16-byte cp.async with zero fill: copies src_bytes (0 or 16) and zero-fills the rest
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
}  // namespace tc

// Issue order of the 2 * n_r convs of a tile (round 2).  Round 1 issued every conv1 first and then the conv2s back to back; with ONE
// x1 operand tile each conv2 -> stage x1(r+1) -> conv2 hand-off then idled the tensor pipe (r02a timeline: 0.4-1.0k cycles each, ~3.5k
// of a 19k-cycle tile).  Interleaved -- C1(0) C1(1) C2(0) C1(2) C2(1) ... C2(n_r - 1) -- conv1(r+1) runs while the epilogue warps turn
// conv1(r) into x1(r), and conv2(r-1) while they work on conv1(r): every conv2 finds its operand staged and the x1 tile free.
__device__ __forceinline__ void mrf3_step(int step, int nrb, int interleave, int& cv, int& r) {
    if (!interleave) { cv = step >= nrb; r = cv ? step - nrb : step; }
    else if (step == 0) { cv = 0; r = 0; }
    else if (step == 2 * nrb - 1) { cv = 1; r = nrb - 1; }
    else { const int j = step - 1; cv = j & 1; r = cv ? (j >> 1) : (j >> 1) + 1; }
}

#define MRF3_SP_PITCH 9
#define MRF3_STAMP(it_, slot_) do { if (dbg_on && (it_) < MRF3_DBG_TILES) a.dbg[(it_) * 48 + (slot_)] = (unsigned long long)clock64(); } while (0)

template <int C>
__global__ void __launch_bounds__(MRF3_THREADS, 1) k_mrf3_tc(const Mrf3Args a, const Mrf3Cfg c, const __grid_constant__ CUtensorMap tmap) {
    // PDL as in k_conv_tc: barriers, biases, conv_post weights, tensor memory and the weight producer only read constants and run
    // under the preceding grid's tail; the epilogue warps and the input loader call pdl_wait() before they touch the tile
    // descriptors (written by k_mrf_tiles just before), activations or outputs.
    pdl_trigger();
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sX = smem;                                       // lrelu(x)  bf16 K-major chunks [C/8][rx][8]
    uint8_t* sX1 = sX + c.x_bytes;                            // lrelu(x1) bf16 K-major chunks [C/8][rx1][8]; then the conv_post operand
    uint8_t* sU = sX1 + c.x1_bytes;                           // mode U: lrelu(h) bf16 K-major chunks [Cin/8][u_rows][8]
    uint8_t* sUW = sU + c.u_bytes;                            // mode U: both polyphase weight halves (resident)
    uint8_t* sW = sUW + c.upw_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)c.nstages * c.slot_bytes);
    const uint32_t bar_full0 = tc::smem_u32(bars);
    const uint32_t bar_empty0 = bar_full0 + 8u * c.nstages;
    const uint32_t bar_fix = bar_empty0 + 8u * c.nstages;
    const uint32_t bar_in = bar_fix;                          // input tile landed                  (32 cp.async arrivals / tile)
    const uint32_t bar_in_free = bar_fix + 8u;                // input tile consumed                (A: 512 / tile, U: commit / tile)
    const uint32_t bar_ups = bar_fix + 16u;                   // ups accumulators ready             (commit / tile)
    const uint32_t bar_x = bar_fix + 24u;                     // X operand staged by E0             (512 / tile, mode U)
    const uint32_t bar_x1 = bar_fix + 32u;                    // x1 operand staged, acc1[r] drained (512 / resblock)
    const uint32_t bar_c1 = bar_fix + 40u;                    // conv1 accumulators ready, one per resblock (commit)   [3]
    const uint32_t bar_c2 = bar_fix + 64u;                    // conv2 MMAs of one resblock complete (commit / resblock)
    const uint32_t bar_acc2_free = bar_fix + 72u;             // conv2 accumulators drained         (512 / tile)
    const uint32_t bar_post_rdy = bar_fix + 80u;              // conv_post operand staged           (512 / tile)
    const uint32_t bar_post_done = bar_fix + 88u;             // conv_post accumulators ready       (commit / tile)
    const uint32_t bar_post_free = bar_fix + 96u;             // conv_post accumulators drained     (512 / tile)
    const uint32_t bar_upw = bar_fix + 104u;                  // ups weights landed                 (once)
    const uint32_t bar_tma = bar_fix + 112u;                  // input tile landed in shared memory (TMA bytes / tile; the loader warp waits)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * c.nstages + MRF3_NBAR);
    float* sB = reinterpret_cast<float*>(smem + c.bias_off);  // bias1 of every resblock, then sum_r bias2_r, then the ups bias
    uint8_t* sWp = smem + c.postw_off;
    float* sP = reinterpret_cast<float*>(smem + c.sp_off);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform in a form ptxas can see (role branches stay converged)
    const bool post = a.post_w != nullptr;
    const bool modeU = a.up_u != 0;
    if (tid == 0) {
        const uint32_t nmw = (uint32_t)c.nmw;           // every commit-tracked barrier gets one arrival per MMA warp
        for (int s = 0; s < c.nstages; s++) { tc::mbar_init(bar_full0 + 8u * s, 1); tc::mbar_init(bar_empty0 + 8u * s, nmw); }
        tc::mbar_init(bar_in, 32);
        tc::mbar_init(bar_in_free, modeU ? nmw : MRF3_EPI_THREADS);
        tc::mbar_init(bar_ups, nmw);
        tc::mbar_init(bar_x, MRF3_EPI_THREADS);
        tc::mbar_init(bar_x1, MRF3_EPI_THREADS);
        for (int r = 0; r < MRF3_MAX_RB; r++) tc::mbar_init(bar_c1 + 8u * r, nmw);
        tc::mbar_init(bar_c2, nmw);
        tc::mbar_init(bar_acc2_free, MRF3_EPI_THREADS);
        tc::mbar_init(bar_post_rdy, MRF3_EPI_THREADS);
        tc::mbar_init(bar_post_done, nmw);
        tc::mbar_init(bar_post_free, MRF3_EPI_THREADS);
        tc::mbar_init(bar_upw, 1);
        tc::mbar_init(bar_tma, 1);
        tc::fence_mbar_init();
    }
    for (int i = tid; i < C; i += MRF3_THREADS) {
        float sum = 0.f;
        for (int r = 0; r < a.nrb; r++) { sB[r * C + i] = __ldg(a.b[r][0] + i); sum += __ldg(a.b[r][1] + i); }
        sB[a.nrb * C + i] = sum;
        sB[(a.nrb + 1) * C + i] = modeU ? __ldg(a.up_b + i) : 0.f;
    }
    if (post) {
        // conv_post as ONE tap-0 GEMM with the 7 filter taps as output columns: P[t][j] = sum_c w[j][c] * y[t][c]
        // (K-major bf16 B operand [C/8][16][8], columns 7-15 zero); the epilogue then adds P[t + j - 3][j] over j.
        // 8 MMAs per tile instead of 56 -- an N=16 MMA costs as much as an N=32 one (the A operand read bounds both)
        __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(sWp);
        for (int i = tid; i < C * 16; i += MRF3_THREADS) {
            const int e = i & 7, n = (i >> 3) & 15, kc = i >> 7;
            wp[i] = __float2bfloat16_rn(n < MRF3_POST_K ? __ldg(a.post_w + n * C + kc * 8 + e) : 0.f);
        }
        tc::fence_proxy_async();
    }
    if (warp == MRF3_EPI_WARPS) tc::tmem_alloc(tc::smem_u32(tmem_slot), (uint32_t)c.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    constexpr int KC = C / 8;                       // 16-byte chunks per row
    constexpr int NG = C / 32;                      // 32-channel groups per row
    const uint32_t lbo_x = (uint32_t)c.rx * 16u, lbo_x1 = (uint32_t)c.rx1 * 16u, lbo_w = (uint32_t)C * 16u;
    const uint32_t acc1_cols = (uint32_t)(c.nb * C);            // per conv1 buffer (one per resblock)
    const uint32_t acc2_col = (uint32_t)a.nrb * acc1_cols;
    const uint32_t accp_col = (uint32_t)(a.nrb - 1) * acc1_cols; // conv_post accumulators (16 columns per M block) reuse the last conv1 buffer
    const int lead = c.hmax + c.h1max;              // window row 0 of the X tile sits `lead` rows before the first stored row
    const int N2 = (a.up_u >> 1) * C;               // columns of one polyphase half
    const int KCU = a.up_cin >> 3;

    auto tile_geom = [&](int tile, long& row0, int& len, int& o0) {
        const int4 d = __ldg(a.tdesc + tile);
        row0 = (long)d.x; len = d.y; o0 = d.z;
    };
    // first input row of the ups pass: floor((o0 - lead) / u)
    auto q_base = [&](int tbase) { return tbase >= 0 ? (tbase >> 2) : -((-tbase + 3) >> 2); };

    if (warp < MRF3_EPI_WARPS) {
        // ===================== epilogues (512 threads) =====================
        pdl_wait();
        // work item of this warp: rows 128*bb + 32*q + lane, channels [32*cg, 32*cg + 32)
        const int q = warp & 3, item = warp >> 2;
        const int bb = item / NG, cg = item - bb * NG;
        const bool active = bb < c.nb;
        const int wr = 128 * bb + 32 * q + lane;                  // window row of this thread
        const float inv_div = 1.f / a.out_div, inv_slope = 1.f / a.slope;
        const float2 slope2 = make_float2(a.slope, a.slope), inv_slope2 = make_float2(inv_slope, inv_slope), inv_div2 = make_float2(inv_div, inv_div);
        uint32_t n_c1[MRF3_MAX_RB] = {0, 0, 0}, n_c2 = 0, n_post = 0, n_ups = 0;
        long p_row0 = 0; int p_len = 0, p_o0 = 0; bool have_prev = false;
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
        int it = 0;
        // conv_post epilogue of the previous tile: acc_post column 0 -> tanh -> audio (thread per row, 32-channel group 0)
        auto post_epilogue = [&]() {
            tc::mbar_wait(bar_post_done, n_post & 1); n_post++;
            tc::tc_fence_after();
            if (active && cg == 0) {
                float p8[8];
                tc::tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + accp_col + (uint32_t)(bb * 16), p8);
                float* dstp = sP + (wr + 3) * MRF3_SP_PITCH;      // row wr of the window at index wr + 3
#pragma unroll
                for (int j = 0; j < MRF3_POST_K; j++) dstp[j] = p8[j];
            }
            tc::tc_fence_before();
            tc::mbar_arrive(bar_post_free);
            asm volatile("bar.sync 1, %0;" ::"n"(MRF3_EPI_THREADS) : "memory");      // the 16 epilogue warps only
            if (active && cg == 0) {
                const int t = p_o0 - c.hmax + wr;                 // stage row == audio sample of this thread
                if (wr >= c.hmax + c.post_halo && wr < c.hmax + c.t_out - c.post_halo && t < p_len) {
                    float v = 0.f;
#pragma unroll
                    for (int j = 0; j < MRF3_POST_K; j++) v += sP[(wr + j) * MRF3_SP_PITCH + j];     // P[wr + j - 3][j]
                    a.audio[p_row0 + t] = tanhf(v);
                }
            }
        };
        // ---- E0: ups accumulators -> + bias -> lrelu -> bf16 operand tile sX (zero outside the utterance).
        // warp = (lane quadrant q, phase pp): input row qq of M block ub  ->  x row u*(qb + qq) + pp.  Runs one tile AHEAD
        // (for the first tile before the loop, then right after E1(n_r - 1) of the previous tile, under its last conv2)
        auto e0_stage = [&](int tile, int itn) {
            long row0; int len, o0;
            tile_geom(tile, row0, len, o0);
            const int tbase = o0 - lead;
            const int pp = warp >> 2;
            const int qb = q_base(tbase);
            tc::mbar_wait(bar_ups, n_ups & 1); n_ups++;
            MRF3_STAMP(itn, 1);
            tc::tc_fence_after();
            const float* ub_bias = sB + (a.nrb + 1) * C;
            for (int ub = 0; ub < c.nub; ub++) {
                const int qq = ub * 128 + 32 * q + lane;
                const int t = ((qb + qq) << 2) + pp;              // time row inside the utterance
                const int r = t - tbase;                          // row of the operand tile
                const bool inu = (t >= 0) && (t < len);
                const bool inr = (r >= 0) && (r < c.rx);
                if (__all_sync(0xffffffffu, !inr)) continue;       // e.g. the tail of the last M block: nothing of this warp lands in the tile
                const uint32_t tl = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(ub * a.up_u * C) + (uint32_t)(pp * C);
#pragma unroll
                for (int n0 = 0; n0 < C; n0 += 16) {
                    float v[16];
                    tc::tmem_ld16(tl + (uint32_t)n0, v);
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        // packed fp32 pairs (FADD2 / FMUL2): the epilogue warps are issue-bound
                        const float4 bv = *reinterpret_cast<const float4*>(ub_bias + n0 + j);
                        const float2 xa = __fadd2_rn(make_float2(v[j], v[j + 1]), make_float2(bv.x, bv.y));
                        const float2 xb2 = __fadd2_rn(make_float2(v[j + 2], v[j + 3]), make_float2(bv.z, bv.w));
                        const float2 sa = __fmul2_rn(xa, slope2), sb = __fmul2_rn(xb2, slope2);
                        pk[j >> 1] = inu ? tc::pack_bf16(fmaxf(xa.x, sa.x), fmaxf(xa.y, sa.y)) : 0u;
                        pk[(j >> 1) + 1] = inu ? tc::pack_bf16(fmaxf(xb2.x, sb.x), fmaxf(xb2.y, sb.y)) : 0u;
                    }
                    if (inr) {
                        *reinterpret_cast<uint4*>(sX + ((size_t)((n0 >> 3) + 0) * c.rx + r) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(sX + ((size_t)((n0 >> 3) + 1) * c.rx + r) * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
            }
            tc::fence_proxy_async();
            tc::tc_fence_before();
            tc::mbar_arrive(bar_x);
            MRF3_STAMP(itn, 2);
        };
        if (modeU && (int)blockIdx.x < a.ntiles) e0_stage(blockIdx.x, 0);
        uint32_t n_x1e = 0;                                       // completions of bar_x1 seen so far (n_r per tile)
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
            long row0; int len, o0;
            tile_geom(tile, row0, len, o0);
            MRF3_STAMP(it, 0);
            // the previous tile's conv_post accumulators drain here, under this tile's first conv1 MMAs; this also
            // guarantees its MMAs no longer read sX1 before E1(0) below overwrites it
            if (post && have_prev) post_epilogue();
            MRF3_STAMP(it, 3);

            const int tm = o0 - c.hmax + wr;
            const bool inr = active && (tm >= 0 && tm < len);
            // this thread's raw x row lives in the operand tile as lrelu(x): row wr + h1max, chunks 4*cg .. 4*cg + 3
            const uint8_t* xop = sX + ((size_t)(4 * cg) * c.rx + (active ? wr + c.h1max : 0)) * 16;
            float2 xacc[16];                        // sum_r (x1_r + bias2_r) of this thread's 32 channels, as packed pairs
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 bs = *reinterpret_cast<const float4*>(sB + a.nrb * C + 32 * cg + 4 * j);      // sum_r bias2_r
                xacc[2 * j] = make_float2(bs.x, bs.y); xacc[2 * j + 1] = make_float2(bs.z, bs.w);
            }
            const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(bb * C + 32 * cg);

            for (int r = 0; r < a.nrb; r++) {
                tc::mbar_wait(bar_c1 + 8u * r, n_c1[r] & 1); n_c1[r]++;
                MRF3_STAMP(it, 4 + 4 * r);
                tc::tc_fence_after();
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; j++) pk[j] = 0u;
                if (active) {                       // warp-uniform: tcgen05.ld is .sync.aligned
                    const float* b1 = sB + r * C + 32 * cg;
#pragma unroll
                    for (int n0 = 0; n0 < 32; n0 += 16) {
                        float v[16];
                        tc::tmem_ld16(tlane + (uint32_t)r * acc1_cols + (uint32_t)n0, v);
#pragma unroll
                        for (int h8 = 0; h8 < 2; h8++) {
                            // 8 channels: one 16-byte chunk of the operand tile, inverted back to x (lrelu is monotone:
                            // x = min(v, v / slope)); exact up to the operand's bf16 rounding
                            const uint4 xo = *reinterpret_cast<const uint4*>(xop + (size_t)((n0 >> 3) + h8) * c.rx * 16);
                            const uint32_t xw[4] = {xo.x, xo.y, xo.z, xo.w};
#pragma unroll
                            for (int p2 = 0; p2 < 4; p2++) {
                                const int ch = n0 + 8 * h8 + 2 * p2;
                                const float2 l = make_float2(tc::bf16_lo_f(xw[p2]), tc::bf16_hi_f(xw[p2]));
                                const float2 li = __fmul2_rn(l, inv_slope2);
                                const float2 xv = make_float2(fminf(l.x, li.x), fminf(l.y, li.y));
                                const float2 vb = __fadd2_rn(make_float2(v[8 * h8 + 2 * p2], v[8 * h8 + 2 * p2 + 1]), *reinterpret_cast<const float2*>(b1 + ch));
                                // ResBlock2: x1 = x + conv1(..) is both the next operand and the residual; ResBlock1 pair: the operand is
                                // conv1(..) alone and the residual of the pair is x itself
                                const float2 x1p = a.rb1 ? vb : __fadd2_rn(vb, xv);
                                xacc[ch >> 1] = __fadd2_rn(xacc[ch >> 1], a.rb1 ? xv : x1p);
                                const float2 xs = __fmul2_rn(x1p, slope2);
                                // conv2 zero-pads x1 beyond the utterance
                                pk[ch >> 1] = inr ? tc::pack_bf16(fmaxf(x1p.x, xs.x), fmaxf(x1p.y, xs.y)) : 0u;
                            }
                        }
                    }
                }
                MRF3_STAMP(it, 5 + 4 * r);
                if (!modeU && r == a.nrb - 1) tc::mbar_arrive(bar_in_free);      // last read of the X tile: the loader may prefetch
                if (r > 0) { tc::mbar_wait(bar_c2, n_c2 & 1); n_c2++; }         // conv2 of resblock r-1 no longer reads sX1
                MRF3_STAMP(it, 6 + 4 * r);
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
                MRF3_STAMP(it, 7 + 4 * r);
                if (a.dbg != nullptr && blockIdx.x == 0 && tid == 480 && it < MRF3_DBG_TILES) a.dbg[it * 48 + (r == 0 ? 23 : 41 + r)] = (unsigned long long)clock64();   // the same event seen by warp 15
            }
            n_x1e += (uint32_t)a.nrb;
            if (modeU && tile + (int)gridDim.x < a.ntiles) {
                // every epilogue thread has finished reading x out of sX (its arrival on bar_x1 follows its last read)
                tc::mbar_wait(bar_x1, (n_x1e - 1u) & 1u);
                e0_stage(tile + gridDim.x, it + 1);
            }
            // ---- final epilogue: out = (acc2 + sum_r(x1_r + b2_r)) / n_r on the central rows
            if (a.naddb && active && inr && (wr >= c.hmax) && (wr < c.hmax + c.t_out)) {
                // the other resblocks' results of this row (bf16, written by their last pairs): fetched while conv2 is still running
                for (int j = 0; j < a.naddb; j++) {
                    const uint4* src = reinterpret_cast<const uint4*>(a.addb[j] + (row0 + tm) * C + 32 * cg);
                    uint4 w4[4];
#pragma unroll
                    for (int n8 = 0; n8 < 4; n8++) w4[n8] = __ldg(src + n8);
#pragma unroll
                    for (int n8 = 0; n8 < 4; n8++) {
                        const uint32_t ww[4] = {w4[n8].x, w4[n8].y, w4[n8].z, w4[n8].w};
#pragma unroll
                        for (int p2 = 0; p2 < 4; p2++)
                            xacc[4 * n8 + p2] = __fadd2_rn(xacc[4 * n8 + p2], make_float2(tc::bf16_lo_f(ww[p2]), tc::bf16_hi_f(ww[p2])));
                    }
                }
            }
            tc::mbar_wait(bar_c2, n_c2 & 1); n_c2++;
            MRF3_STAMP(it, 16);
            tc::tc_fence_after();
            if (active) {
                const bool central = (wr >= c.hmax) && (wr < c.hmax + c.t_out);
                const bool st = central && inr;
                if (post || a.outb) {
                    // lrelu(out) as bf16: the conv_post operand (same placement as x1, zero outside the utterance) or the next
                    // stage's input rows in HBM
                    const float osl = post ? a.post_slope : a.outb_slope;
                    uint32_t pk[16];
#pragma unroll
                    for (int n0 = 0; n0 < 32; n0 += 16) {
                        float v[16];
                        tc::tmem_ld16(tlane + acc2_col + (uint32_t)n0, v);
#pragma unroll
                        const float2 osl2 = make_float2(osl, osl);
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const float2 o = __fmul2_rn(__fadd2_rn(make_float2(v[j], v[j + 1]), xacc[(n0 + j) >> 1]), inv_div2);
                            const float2 os = __fmul2_rn(o, osl2);
                            pk[(n0 + j) >> 1] = st ? tc::pack_bf16(fmaxf(o.x, os.x), fmaxf(o.y, os.y)) : 0u;
                        }
                    }
                    if (post) {
                        const int row1 = wr + c.hmax;
#pragma unroll
                        for (int n8 = 0; n8 < 4; n8++)
                            *reinterpret_cast<uint4*>(sX1 + ((size_t)(4 * cg + n8) * c.rx1 + row1) * 16) =
                                make_uint4(pk[4 * n8], pk[4 * n8 + 1], pk[4 * n8 + 2], pk[4 * n8 + 3]);
                    } else if (st) {
                        uint4* orow = reinterpret_cast<uint4*>(a.outb + (row0 + tm) * C + 32 * cg);
#pragma unroll
                        for (int n8 = 0; n8 < 4; n8++) orow[n8] = make_uint4(pk[4 * n8], pk[4 * n8 + 1], pk[4 * n8 + 2], pk[4 * n8 + 3]);
                    }
                } else {
                    float* orow = a.out + (row0 + tm) * C + 32 * cg;
#pragma unroll
                    for (int n0 = 0; n0 < 32; n0 += 16) {
                        float v[16];
                        tc::tmem_ld16(tlane + acc2_col + (uint32_t)n0, v);
                        if (st) {
#pragma unroll
                            for (int qd = 0; qd < 4; qd++) {
                                float4 ov, old = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (a.accumulate) old = *(reinterpret_cast<const float4*>(orow + n0) + qd);
                                ov.x = (v[4 * qd + 0] + xacc[(n0 + 4 * qd) >> 1].x + old.x) * inv_div;
                                ov.y = (v[4 * qd + 1] + xacc[(n0 + 4 * qd) >> 1].y + old.y) * inv_div;
                                ov.z = (v[4 * qd + 2] + xacc[(n0 + 4 * qd + 2) >> 1].x + old.z) * inv_div;
                                ov.w = (v[4 * qd + 3] + xacc[(n0 + 4 * qd + 2) >> 1].y + old.w) * inv_div;
                                *(reinterpret_cast<float4*>(orow + n0) + qd) = ov;
                            }
                        }
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(bar_acc2_free);       // the next tile's conv2 may overwrite the accumulators
            MRF3_STAMP(it, 17);
            if (post) {
                tc::fence_proxy_async();
                tc::mbar_arrive(bar_post_rdy);
                p_row0 = row0; p_len = len; p_o0 = o0; have_prev = true;
            }
        }
        if (post && have_prev) post_epilogue();
    } else if (warp == MRF3_EPI_WARPS) {
        // ===================== weight producer =====================
        {
            // the whole warp runs the loop converged; only the copies / arrivals are predicated on the elected lane (conv_tc.cuh)
            const uint32_t el = tc::elect_flag();
            if (modeU) {
                tc::mbar_expect_tx_e(bar_upw, (uint32_t)c.upw_bytes, el);
                const uint32_t half_bytes = (uint32_t)c.upw_bytes >> 1;
                tc::bulk_g2s_e(tc::smem_u32(sUW), a.up_w[0], half_bytes, bar_upw, el);
                tc::bulk_g2s_e(tc::smem_u32(sUW) + half_bytes, a.up_w[1], half_bytes, bar_upw, el);
            }
            uint32_t s = 0, ph = 1;                           // ring slot and the parity to wait for on its "empty" barrier
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                if (c.resident && tile != (int)blockIdx.x) break;
                // same order as the MMA warps issue the convs (mrf3_step): C1(0) C1(1) C2(0) C1(2) C2(1) C2(2)
                for (int step = 0; step < 2 * a.nrb; step++) {
                        int cv, r;
                        mrf3_step(step, a.nrb, a.interleave, cv, r);
                        const __nv_bfloat16* wsrc = a.w[r][cv];
                        for (int tap = 0; tap < a.k[r]; tap++) {
                            if (!c.resident) tc::mbar_wait(bar_empty0 + 8u * s, ph);
                            const uint32_t fb = bar_full0 + 8u * s;
                            tc::mbar_expect_tx_e(fb, (uint32_t)c.slot_bytes, el);
                            tc::bulk_g2s_e(tc::smem_u32(sW) + s * (uint32_t)c.slot_bytes, wsrc, (uint32_t)c.slot_bytes, fb, el);
                            wsrc += C * C;
                            if (++s == (uint32_t)c.nstages) { s = 0; ph ^= 1u; }
                        }
                    }
            }
        }
    } else if (warp == MRF3_EPI_WARPS + 1 || warp == MRF3_EPI_WARPS + 3) {
        // ===================== MMA issuers: warp mw owns M blocks [bb_lo, bb_hi) =====================
        const int mw = (warp == MRF3_EPI_WARPS + 3) ? 1 : 0;
        const int bb_lo = mw * (c.nb / c.nmw), bb_hi = bb_lo + c.nb / c.nmw;
        const int ub_lo = (c.nmw == 2 && c.nub == 2) ? mw : 0, ub_hi = (c.nmw == 2 && c.nub == 2) ? mw + 1 : (mw == 0 ? c.nub : 0);
        if (mw < c.nmw) {
            // warp-uniform issue loop: all 32 lanes run it converged (and wait on the barriers); tcgen05.mma / tcgen05.commit are
            // predicated on the elected lane, descriptors and barrier addresses live in uniform registers (conv_tc.cuh, round 2)
            const uint32_t el = tc::elect_flag();
            const uint32_t tmem_base = tc::uniform_u32(*tmem_slot);
            const uint32_t idesc = tc::make_idesc(128, C), idesc_post = tc::make_idesc(128, 16), idesc_up = tc::make_idesc(128, N2 > 0 ? N2 : 16);
            const uint32_t sX_u = tc::smem_u32(sX), sX1_u = tc::smem_u32(sX1), sW_u = tc::smem_u32(sW);
            const uint64_t dhi_x = tc::make_desc(0, lbo_x, 128u), dhi_x1 = tc::make_desc(0, lbo_x1, 128u), dhi_w = tc::make_desc(0, lbo_w, 128u);
            const uint64_t dhi_wp = tc::make_desc(0, 256u, 128u);
            const uint64_t dhi_u = tc::make_desc(0, (uint32_t)c.u_rows * 16u, 128u), dhi_uw = tc::make_desc(0, (uint32_t)N2 * 16u, 128u);
            const uint64_t bd_step = (uint64_t)((2u * lbo_w) >> 4);
            const uint64_t ad_step_x = (uint64_t)((2u * lbo_x) >> 4), ad_step_x1 = (uint64_t)((2u * lbo_x1) >> 4);
            const uint64_t ad_step_u = (uint64_t)((2u * (uint32_t)c.u_rows * 16u) >> 4), bd_step_u = (uint64_t)((2u * (uint32_t)N2 * 16u) >> 4);
            const uint32_t x16 = (sX_u >> 4) + (uint32_t)c.h1max, x116 = (sX1_u >> 4) + (uint32_t)c.hmax;   // 16-byte units == rows
            const uint32_t wp16 = tc::smem_u32(sWp) >> 4, u16 = tc::smem_u32(sU) >> 4, uw16 = tc::smem_u32(sUW) >> 4;
            const uint32_t uw_tap16 = (uint32_t)(KCU * N2);          // 16-byte units per tap of one polyphase half
            uint32_t s = 0, ph = 0, it = 0, n_x1 = 0;           // ring slot / parity of its "full" barrier
            const uint32_t slot16 = (uint32_t)c.slot_bytes >> 4;
            const int nbw = bb_hi - bb_lo;                     // M blocks of this warp: 1 or 2
            const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && mw == 0 && el != 0;
            // ---- ConvTranspose1d as a GEMM over input rows: accumulator (ub) columns [half * N2, +N2) = taps of that half
            auto issue_ups = [&](uint32_t itn) {
                tc::mbar_wait(bar_in, itn & 1);
                if (itn == 0) tc::mbar_wait(bar_upw, 0);
                tc::tc_fence_after();
                for (int ub = ub_lo; ub < ub_hi; ub++)
                    for (int half = 0; half < 2; half++) {
                        const uint32_t dcol = tmem_base + (uint32_t)(ub * a.up_u * C) + (uint32_t)(half * N2);
                        for (int tap = 0; tap < 2; tap++) {
                            // operand row 0 of sU is input row qb - 1: half A reads rows q-1, q; half B rows q, q+1
                            uint64_t ad = dhi_u | (uint64_t)((u16 + (uint32_t)(ub * 128 + half + tap)) & 0x3FFF);
                            uint64_t bd = dhi_uw | (uint64_t)((uw16 + (uint32_t)(half * 2 + tap) * uw_tap16) & 0x3FFF);
                            for (int k16 = 0; k16 < (a.up_cin >> 4); k16++) {
                                tc::umma_bf16_e(dcol, ad, bd, idesc_up, (tap > 0 || k16) ? 1u : 0u, el);
                                ad += ad_step_u; bd += bd_step_u;
                            }
                        }
                    }
                tc::umma_commit_e(bar_in_free, el);             // the loader may overwrite sU
                tc::umma_commit_e(bar_ups, el);
                MRF3_STAMP((int)itn, 21);
            };
            // conv_post partial sums P[t][tap] of tile `itp` over its stage output sitting in sX1 (tap offset 0, N = 16).  Issued
            // one tile LATE, between C1(0) and C1(1) of the next tile: waiting here for the final epilogue (which stages the
            // operand) used to leave the tensor pipe idle for ~3.5k of a 19k-cycle tile (r01d timeline); now conv1(0) of the next
            // tile runs under it.  The post accumulators alias conv1 buffer n_r - 1, whose MMAs wait for bar_post_free below.
            auto issue_post = [&](uint32_t itp) {
                tc::mbar_wait(bar_post_rdy, itp & 1);
                MRF3_STAMP((int)itp, 40);
                tc::tc_fence_after();
                for (int bb = bb_lo; bb < bb_hi; bb++) {
                    uint64_t ad = dhi_x1 | (uint64_t)((x116 + 128u * (uint32_t)bb) & 0x3FFF);
                    uint64_t bd = dhi_wp | (uint64_t)(wp16 & 0x3FFF);
#pragma unroll
                    for (int k16 = 0; k16 < C / 16; k16++) {
                        tc::umma_bf16_e(tmem_base + accp_col + (uint32_t)(bb * 16), ad, bd, idesc_post, k16 ? 1u : 0u, el);
                        ad += ad_step_x1; bd += 32u;     // two 8-channel chunks of 16 x 16 B
                    }
                }
                tc::umma_commit_e(bar_post_done, el);
                MRF3_STAMP((int)itp, 41);
            };
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
                MRF3_STAMP((int)it, 20);
                if (!modeU) { tc::mbar_wait(bar_in, it & 1); tc::tc_fence_after(); }
                if (modeU) {
                    if (it == 0) issue_ups(0);                // later tiles: issued one tile ahead, between C2(n_r - 2) and C2(n_r - 1)
                    tc::mbar_wait(bar_x, it & 1);
                    tc::tc_fence_after();
                }
                MRF3_STAMP((int)it, 22);
                for (int step = 0; step < 2 * a.nrb; step++) {
                    {
                        int cv, r;
                        mrf3_step(step, a.nrb, a.interleave, cv, r);
                        if (cv == 1) {
                            // the next tile's ConvTranspose goes in before the last conv2: its accumulators (conv1 buffers 0..) were
                            // drained when x1(n_r - 2) was staged, and E0 of the next tile then runs under conv2(n_r - 1)
                            if (modeU && r == a.nrb - 1 && tile + (int)gridDim.x < a.ntiles) issue_ups(it + 1);
                            MRF3_STAMP((int)it, 36 + r);
                            tc::mbar_wait(bar_x1, n_x1 & 1); n_x1++;                       // x1(r) staged, acc1[r] drained
                            if (r == 0) MRF3_STAMP((int)it, 18);
                            if (r == 0 && it > 0) tc::mbar_wait(bar_acc2_free, (it - 1) & 1);   // previous tile's output drained
                            tc::tc_fence_after();
                        } else if (post && it > 0) {
                            if (r == (a.nrb > 1 ? 1 : 0)) issue_post(it - 1);              // previous tile's conv_post, behind this tile's C1(0)
                            if (r == a.nrb - 1) {
                                tc::mbar_wait(bar_post_free, (it - 1) & 1);                // conv_post accumulators (same columns) drained
                                tc::tc_fence_after();
                            }
                        }
                        MRF3_STAMP((int)it, 24 + 2 * (cv * 3 + r));
                        const int kr = a.k[r];
                        const uint32_t dil = (uint32_t)(cv ? a.d2[r] : a.d1[r]);
                        // descriptor low words: start address (16-byte units == rows) | LBO << 16; a tap advances the A start by `dil`
                        // rows, an M block by 128 rows, a K=16 step by two chunk planes (2 * LBO)
                        const uint32_t a_lbo16 = cv ? (lbo_x1 >> 4) : (lbo_x >> 4);
                        const uint32_t a_k16 = 2u * a_lbo16, b_k16 = 2u * (lbo_w >> 4);
                        const uint32_t ahi = (uint32_t)((cv ? dhi_x1 : dhi_x) >> 32), bhi = (uint32_t)(dhi_w >> 32);
                        const uint32_t dcol0 = tmem_base + (cv ? acc2_col : (uint32_t)r * acc1_cols) + (uint32_t)(bb_lo * C);
                        uint32_t alo = ((cv ? x116 : x16) - (uint32_t)((kr - 1) >> 1) * dil + 128u * (uint32_t)bb_lo) | (a_lbo16 << 16);   // tap 0
                        const uint32_t acc_first = (cv && r > 0) ? 1u : 0u;      // conv1: fresh accumulator per resblock; conv2 accumulates across resblocks
                        for (int tap = 0; tap < kr; tap++, alo += dil) {
                            if (!c.resident || it == 0) { tc::mbar_wait(bar_full0 + 8u * s, ph); tc::tc_fence_after(); }
                            const uint32_t blo = ((sW_u >> 4) + s * slot16) | ((lbo_w >> 4) << 16);
                            const uint32_t acc0 = tap > 0 ? 1u : acc_first;
#pragma unroll
                            for (int k16 = 0; k16 < C / 16; k16++)
                                tc::umma_bf16_lh_e(dcol0, alo + (uint32_t)k16 * a_k16, ahi, blo + (uint32_t)k16 * b_k16, bhi, idesc, k16 ? 1u : acc0, el);
                            if (nbw == 2) {
#pragma unroll
                                for (int k16 = 0; k16 < C / 16; k16++)
                                    tc::umma_bf16_lh_e(dcol0 + (uint32_t)C, alo + 128u + (uint32_t)k16 * a_k16, ahi, blo + (uint32_t)k16 * b_k16, bhi, idesc, k16 ? 1u : acc0, el);
                            }
                            if (!c.resident) tc::umma_commit_e(bar_empty0 + 8u * s, el);
                            if (++s == (uint32_t)c.nstages) { s = 0; ph ^= 1u; }
                        }
                        tc::umma_commit_e(cv ? bar_c2 : (bar_c1 + 8u * (uint32_t)r), el);
                        MRF3_STAMP((int)it, 25 + 2 * (cv * 3 + r));
                    }
                }
                if (c.resident) { s = 0; }
            }
            if (post && it > 0) issue_post(it - 1);          // the last tile's conv_post
        }
    } else {
        // ===================== input-row loader (warp 18): TMA boxes, or cp.async 16 B per lane with zero fill outside the utterance =====================
        pdl_wait();
        uint32_t n_free = 0;
        const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && lane == 0;
        const uint32_t el = tc::elect_flag();
        int it = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
            long row0; int len, o0;
            tile_geom(tile, row0, len, o0);
            const int tbase = o0 - lead;
            MRF3_STAMP(it, 44);
            if (tile != (int)blockIdx.x) { tc::mbar_wait(bar_in_free, n_free & 1); n_free++; }
            MRF3_STAMP(it, 45);
            if (c.tma) {
                // tile row j of plane kc <- tensor row g0 + j, channels [8 kc, 8 kc + 8); rows [vlo, vhi) belong to this tile's utterance
                int g0, rows, planes, vlo, vhi;
                uint32_t dst0;
                if (modeU) {
                    const int qb = q_base(tbase);
                    g0 = (int)(row0 >> 2) + qb - 1; rows = c.u_rows; planes = KCU; dst0 = tc::smem_u32(sU);
                    vlo = 1 - qb; vhi = (len >> 2) - qb + 1;
                } else {
                    g0 = (int)row0 + tbase; rows = c.rx; planes = KC; dst0 = tc::smem_u32(sX);
                    vlo = -tbase; vhi = len - tbase;
                }
                vlo = min(max(vlo, 0), rows); vhi = min(max(vhi, vlo), rows);
                tc::fence_proxy_async();                          // the previous tile's generic-proxy reads of this buffer are behind us
                tc::mbar_expect_tx_e(bar_tma, (uint32_t)(planes * rows) * 16u, el);
                for (int kc = 0; kc < planes; kc++)
                    for (int bx = 0; bx < c.nboxes; bx++)
                        tc::tma_load_2d_e(dst0 + (uint32_t)(kc * rows + bx * c.box_rows) * 16u, &tmap, kc * 8, g0 + bx * c.box_rows, bar_tma, el);
                tc::mbar_wait(bar_tma, (uint32_t)it & 1u);
                // rows of OTHER utterances (inside the array, so the TMA unit delivered them): zero, as the convolutions' padding
                const int nz = vlo + (rows - vhi);
                for (int i = lane; i < nz * planes; i += 32) {
                    const int kc = i / nz, j = i - kc * nz;
                    const int r = j < vlo ? j : vhi + (j - vlo);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst0 + (uint32_t)(kc * rows + r) * 16u), "r"(0u) : "memory");
                }
                tc::fence_proxy_async();
                tc::mbar_arrive(bar_in);
                MRF3_STAMP(it, 46);
                continue;
            }
            if (modeU) {
                const int qb = q_base(tbase);
                const long rin0 = row0 >> 2;                      // first input row of the utterance (rate / u rows per frame)
                const int len_in = len >> 2;
                const uint32_t dst0 = tc::smem_u32(sU);
                const int total = c.u_rows * KCU;
                for (int i = lane; i < total; i += 32) {
                    const int j = i / KCU, kc = i - j * KCU;      // kc fastest: contiguous global reads
                    const int qi = qb - 1 + j;
                    const bool ok = (qi >= 0) && (qi < len_in);
                    const __nv_bfloat16* src = a.hb + (rin0 + (ok ? qi : 0)) * a.up_cin + kc * 8;
                    tc::cp_async16_zfill(dst0 + (uint32_t)(kc * c.u_rows + j) * 16u, src, ok ? 16u : 0u);
                }
            } else {
                const uint32_t dst0 = tc::smem_u32(sX);
                const int total = c.rx * KC;
                for (int i = lane; i < total; i += 32) {
                    const int r = i / KC, kc = i - r * KC;
                    const int t = tbase + r;
                    const bool ok = (t >= 0) && (t < len);
                    const __nv_bfloat16* src = a.xb + (row0 + (ok ? t : 0)) * C + kc * 8;
                    tc::cp_async16_zfill(dst0 + (uint32_t)(kc * c.rx + r) * 16u, src, ok ? 16u : 0u);
                }
            }
            // the tile is read by the tensor core (async proxy): complete this lane's copies, make them visible, then arrive
            asm volatile("cp.async.wait_all;" ::: "memory");
            tc::fence_proxy_async();
            tc::mbar_arrive(bar_in);
            MRF3_STAMP(it, 46);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MRF3_EPI_WARPS) tc::tmem_dealloc(tmem_base, (uint32_t)c.tmem_cols);
}

// ------------------------------------------------------------------------------------------ host side
// min_hmax: plan with at least this second-conv halo, so that kernels of one stage with different kernel sizes share one tile table
static inline bool mrf3_plan(const Mrf3Args& a, Mrf3Cfg& c, int nb_pref, bool fuse_post, bool use_tma = true, int min_hmax = 0) {
    memset(&c, 0, sizeof c);
    if (a.C != 32 && a.C != 64 && !(a.C == 128 && a.rb1)) return false;      // 128 channels: ResBlock1 pairs only (one M block per tile)
    if (a.nrb < 1 || a.nrb > MRF3_MAX_RB) return false;
    const bool modeU = a.up_u != 0;
    if (modeU && (a.up_u != 4 || a.up_cin % 16 || a.up_cin < 16 || a.up_cin > 128 || (a.up_u / 2) * a.C > 256)) return false;
    int hmax = 0, h1max = 0, npieces = 0;
    for (int r = 0; r < a.nrb; r++) {
        if (a.k[r] % 2 == 0) return false;
        const int h1 = a.d1[r] * (a.k[r] - 1) / 2, h2 = a.d2[r] * (a.k[r] - 1) / 2;
        hmax = h2 > hmax ? h2 : hmax; h1max = h1 > h1max ? h1 : h1max;
        npieces += 2 * a.k[r];
    }
    if (hmax < min_hmax) hmax = min_hmax;
    if (a.rb1 && a.nrb != 1) return false;
    c.hmax = hmax; c.h1max = h1max; c.npieces = npieces;
    c.slot_bytes = a.C * a.C * 2;
    c.post_halo = fuse_post ? (MRF3_POST_K - 1) / 2 : 0;
    if (fuse_post && hmax < c.post_halo) return false;        // the conv_post taps must stay inside the x1 tile
    const int limit = 225 * 1024;
    const int ng = a.C / 32;
    const int postw_bytes = fuse_post ? a.C * 16 * 2 : 0;
    for (int nb = nb_pref; nb >= 1; nb--) {
        if (nb == 3) continue;
        if (nb * ng > MRF3_EPI_WARPS / 4) continue;           // one (block, 32-channel group) item per epilogue warp
        if ((a.nrb + 1) * nb * a.C > 512) continue;           // one conv1 buffer per resblock + conv2 accumulators (conv_post aliases)
        if (fuse_post && nb * 16 > nb * a.C) continue;
        c.nb = nb; c.nmw = (nb % 2 == 0) ? 2 : 1; c.span = 128 * nb; c.t_out = c.span - 2 * hmax; c.t_step = c.t_out - 2 * c.post_halo;
        if (c.t_step < 32) continue;
        // rows of a TMA-fed tile: a whole number of equal boxes of a multiple of 8 rows (planes stay 128-byte aligned); otherwise odd
        // (the cp.async loader's 16-byte writes of one row's chunks then land in different banks)
        auto tma_rows = [&](int need) {
            const int r8 = (need + 7) / 8 * 8, nbx = (r8 + MRF3_TMA_BOX - 1) / MRF3_TMA_BOX;
            c.nboxes = nbx; c.box_rows = ((r8 + nbx - 1) / nbx + 7) / 8 * 8;
            return c.nboxes * c.box_rows;
        };
        c.tma = use_tma ? 1 : 0;
        c.rx = (use_tma && !modeU) ? tma_rows(c.span + 2 * h1max) : ((c.span + 2 * h1max + 7) / 8) * 8 + 1;
        c.rx1 = ((c.span + 2 * hmax + 7) / 8) * 8 + 1;
        c.x_bytes = ((a.C / 8) * c.rx * 16 + 127) / 128 * 128;
        c.x1_bytes = ((a.C / 8) * c.rx1 * 16 + 127) / 128 * 128;
        c.nub = 0; c.u_rows = 0; c.u_bytes = 0; c.upw_bytes = 0;
        if (modeU) {
            // input rows floor(tbase/u) - 1 ... : the ups pass must cover rx + (u - 1) output rows
            c.nub = (c.rx + a.up_u - 1 + 128 * a.up_u - 1) / (128 * a.up_u);
            // ups accumulators live in conv1 buffers 0 .. n_r - 2: those are drained when the next tile's ups pass is issued (before
            // the last conv2), and the last buffer doubles as the conv_post accumulator
            if (a.nrb < 2 || c.nub * a.up_u * a.C > (a.nrb - 1) * nb * a.C) continue;
            c.u_rows = use_tma ? tma_rows(c.nub * 128 + 2) : ((c.nub * 128 + 2 + 7) / 8) * 8 + 1;
            c.u_bytes = ((a.up_cin / 8) * c.u_rows * 16 + 127) / 128 * 128;
            c.upw_bytes = 2 * 2 * a.up_cin * (a.up_u / 2) * a.C * 2;
        }
        const long fixed = (long)c.x_bytes + c.x1_bytes + c.u_bytes + c.upw_bytes;
        const long tail = (2 * MRF3_MAX_STAGES + MRF3_NBAR) * 8 + 32 + (MRF3_MAX_RB + 2) * a.C * 4 + postw_bytes + 512 +
                          (fuse_post ? (128 * nb + 8) * MRF3_SP_PITCH * 4 + 128 : 0);
        const long res_bytes = (long)npieces * c.slot_bytes;
        if (npieces <= MRF3_MAX_STAGES && fixed + res_bytes + tail <= limit) { c.resident = 1; c.nstages = npieces; }
        else {
            c.resident = 0;
            long room = limit - fixed - tail;
            int ns = (int)(room / c.slot_bytes);
            if (ns > MRF3_MAX_STAGES) ns = MRF3_MAX_STAGES;
            if (ns > npieces) ns = npieces;
            if (ns < 3) continue;
            c.nstages = ns;
        }
        c.bias_off = (int)((fixed + (long)c.nstages * c.slot_bytes + (2 * c.nstages + MRF3_NBAR) * 8 + 32 + 15) / 16 * 16);
        c.postw_off = (c.bias_off + (MRF3_MAX_RB + 2) * a.C * 4 + 127) / 128 * 128;
        c.sp_off = (c.postw_off + postw_bytes + 127) / 128 * 128;
        c.smem_bytes = c.sp_off + (fuse_post ? (c.span + 8) * MRF3_SP_PITCH * 4 : 0);
        if (c.smem_bytes > limit) continue;
        int cols = 32; while (cols < (a.nrb + 1) * nb * a.C) cols <<= 1;
        c.tmem_cols = cols;
        return true;
    }
    return false;
}

template <int C>
static inline cudaError_t mrf3_launch_t(const Mrf3Args& a, const Mrf3Cfg& c, const CUtensorMap& tmap, int num_sms, cudaStream_t st) {
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_mrf3_tc<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_mrf3_tc<C>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    int gx = num_sms;                       // one persistent CTA per SM (TMEM: 512 columns each)
    if (gx > a.ntiles) gx = a.ntiles;
    if (gx < 1) return cudaSuccess;
    launch_k(k_mrf3_tc<C>, gx, MRF3_THREADS, c.smem_bytes, st, a, c, tmap);
    return cudaGetLastError();
}

// the per-tile descriptors the kernel reads (launched separately so that the kernel proper can be timed alone)
static inline cudaError_t mrf3_tiles_launch(const Mrf3Args& a, const Mrf3Cfg& c, cudaStream_t st) {
    if (a.ntiles < 1) return cudaSuccess;
    launch_k(k_mrf_tiles, (a.ntiles + 255) / 256, 256, 0, st, a.cu, a.tile_cu, a.B, a.rate, a.ntiles, c.t_step, c.post_halo,
                                                         const_cast<int4*>(a.tdesc));
    return cudaGetLastError();
}

// tmap: the tensor map of the kernel's input rows (mode A: xb [rows, C]; mode U: hb [rows / u, up_cin]) with box {8, c.box_rows};
// ignored (may be zero-filled) when c.tma == 0
static inline cudaError_t mrf3_kernel_launch(const Mrf3Args& a, const Mrf3Cfg& c, const CUtensorMap& tmap, int num_sms, cudaStream_t st) {
    if (a.C == 32) return mrf3_launch_t<32>(a, c, tmap, num_sms, st);
    if (a.C == 64) return mrf3_launch_t<64>(a, c, tmap, num_sms, st);
    if (a.C == 128) return mrf3_launch_t<128>(a, c, tmap, num_sms, st);
    return cudaErrorInvalidConfiguration;
}
