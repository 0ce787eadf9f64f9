// tcgen05 / TMEM implicit-GEMM conv1d for sm_100a (bf16 operands, fp32 accumulate in TMEM).
//
// Mapping (DESIGN.md "conv_tc"):  D[M = 128 time rows, N = C_out tile] += A[M, K = C_in] * B[N, K]^T
// once per tap; a dilated / polyphase tap is the SAME shared-memory activation tile read at a
// shifted row, which with the K-major no-swizzle canonical layout
//     addr(row, k) = (k / 8) * LBO + row * 16 B + (k % 8) * 2 B        (SBO = 128 B)
// is just a 16-byte-aligned start-address offset in the UMMA smem descriptor -- no im2col
// copy and no re-fetch per tap.
//   * activations: fp32 channel-last in HBM -> (leaky-relu) -> bf16 -> smem by the 4 epilogue
//     warps (the fused "input activation"), zero-filled outside the utterance;
//   * weights: pre-packed bf16 [tap][C_in/8][N][8] streamed by one producer thread with
//     cp.async.bulk (TMA bulk, SASS UBLKCP) through an mbarrier ring, or kept resident in smem
//     for the whole persistent CTA when they fit;
//   * MMAs: one elected thread, tcgen05.mma.cta_group::1.kind::f16 (SASS UTCHMMA), M=128,
//     N<=256, K=16; completion via tcgen05.commit -> mbarrier;
//   * epilogue: tcgen05.ld 32x32b (SASS LDTM) -> bias / speaker bias / residual / gate /
//     split-accumulate / tanh -> fp32 stores.
#pragma once
#include "common.cuh"
#include "kernels_f32.cuh"

#ifndef TC_DESC_SWAP_LBO_SBO
#define TC_DESC_SWAP_LBO_SBO 0
#endif

#define TC_M 128
#define TC_THREADS 192
#define TC_PIECE_CH 64
#define TC_MAX_STAGES 24
#define TC_SPIN_LIMIT (1u << 28)

struct TcCfg {
    int ntile;        // N columns per CTA (multiple of 16, <= 256)
    int tmem_cols;    // power of two >= 32
    int rows_a;       // activation rows held in smem (odd: conflict-free 16 B chunk scatter)
    int min_off;      // smallest tap offset
    int piece_ch;     // channels per weight piece (<= 64)
    int cpt;          // pieces per tap
    int npieces;
    int nstages;
    int resident;     // all pieces stay in smem for the CTA's lifetime
    int slot_bytes;
    int a_bytes;
    int smem_bytes;
    int vec;          // epilogue may use 128-bit accesses
};

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try(bar, parity)) {
        if (++spins > TC_SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, SWIZZLE_NONE canonical layout descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start>>4 | [16,30) leading byte offset>>4 | [32,46) stride byte offset>>4 | [46,48) version=1 | [61,64) layout=0
// LBO = byte distance between the two 8-element K chunks of one K=16 MMA; SBO = distance between
// 8-row core-matrix groups (128 B: rows are linear at a 16 B pitch).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
#if TC_DESC_SWAP_LBO_SBO
    const uint32_t t = lbo_bytes; lbo_bytes = sbo_bytes; sbo_bytes = t;
#endif
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// kind::f16 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&p);
}

}  // namespace tc

__global__ void __launch_bounds__(TC_THREADS) k_conv_tc(const ConvArgs a, const TcCfg c) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sW = smem + c.a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)c.nstages * c.slot_bytes);
    // bars: full[nstages], empty[nstages], a_full, acc_full ; then tmem ptr
    const uint32_t bar_full0 = tc::smem_u32(bars);
    const uint32_t bar_empty0 = bar_full0 + 8u * c.nstages;
    const uint32_t bar_afull = bar_empty0 + 8u * c.nstages;
    const uint32_t bar_acc = bar_afull + 8u;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * c.nstages + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ny = blockIdx.y;

    if (tid == 0) {
        for (int s = 0; s < c.nstages; s++) { tc::mbar_init(bar_full0 + 8u * s, 1); tc::mbar_init(bar_empty0 + 8u * s, 1); }
        tc::mbar_init(bar_afull, 128);
        tc::mbar_init(bar_acc, 1);
        tc::fence_mbar_init();
    }
    if (warp == 4) tc::tmem_alloc(tc::smem_u32(tmem_slot), (uint32_t)c.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int kc_total = a.cin >> 3;                    // 16-byte chunks per activation row
    const uint32_t lbo_a = (uint32_t)c.rows_a * 16u;
    const uint32_t lbo_b = (uint32_t)c.ntile * 16u;

    if (warp < 4) {
        // ================= activation loader + epilogue (128 threads) =================
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
            const int b = find_segment(a.tile_cu, a.B, tile);
            const int t0 = (tile - __ldg(a.tile_cu + b)) * TC_M;
            const int cb0 = __ldg(a.cu + b), cb1 = __ldg(a.cu + b + 1);
            const long row0 = (long)cb0 * a.rate;
            const int len = (cb1 - cb0) * a.rate;
            // previous tile's MMAs are complete (we waited on acc_full) -> sA may be overwritten
            const int items = c.rows_a * kc_total;
            for (int i = tid; i < items; i += 128) {
                const int r = i / kc_total, kc = i - r * kc_total;
                const int t = t0 + c.min_off + r;
                uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                if (t >= 0 && t < len) {
                    const float4* src = reinterpret_cast<const float4*>(a.x + (row0 + t) * a.ldx + a.xcol + kc * 8);
                    float4 v0 = __ldg(src), v1 = __ldg(src + 1);
                    if (a.in_act) {
                        v0.x = leaky(v0.x, a.in_slope); v0.y = leaky(v0.y, a.in_slope); v0.z = leaky(v0.z, a.in_slope); v0.w = leaky(v0.w, a.in_slope);
                        v1.x = leaky(v1.x, a.in_slope); v1.y = leaky(v1.y, a.in_slope); v1.z = leaky(v1.z, a.in_slope); v1.w = leaky(v1.w, a.in_slope);
                    }
                    pk.x = tc::pack_bf16(v0.x, v0.y); pk.y = tc::pack_bf16(v0.z, v0.w);
                    pk.z = tc::pack_bf16(v1.x, v1.y); pk.w = tc::pack_bf16(v1.z, v1.w);
                }
                *reinterpret_cast<uint4*>(sA + ((size_t)kc * c.rows_a + r) * 16) = pk;
            }
            tc::fence_proxy_async();       // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc::mbar_arrive(bar_afull);
            // ---- epilogue
            tc::mbar_wait(bar_acc, it & 1);
            tc::tc_fence_after();
            const int m = warp * 32 + lane;
            const int t = t0 + m;
            const bool rowok = (t < len);
            const long row = row0 + t;
            const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
            for (int n0 = 0; n0 < c.ntile; n0 += 16) {
                float v[16];
                tc::tmem_ld16(trow + (uint32_t)n0, v);
                const int ng = ny * c.ntile + n0;           // global output column of v[0]
                if (!rowok || ng >= a.n) continue;
                if (a.epi == EPI_GATE) {
                    // interleaved (tanh-arg, sigmoid-arg) column pairs; commons.py:99-106
                    const float* ur = a.utab ? (a.utab + (long)__ldg(a.uidx + b) * a.utab_ld) : nullptr;
                    float g[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float va = v[2 * j], vb = v[2 * j + 1];
                        const int n = ng + 2 * j;
                        if (n + 1 < a.n) {
                            if (a.bias) { va += __ldg(a.bias + n); vb += __ldg(a.bias + n + 1); }
                            if (ur) { va += __ldg(ur + n); vb += __ldg(ur + n + 1); }
                        }
                        g[j] = tanhf(va) * (1.f / (1.f + expf(-vb)));
                    }
                    float* dst = a.out + row * a.ldo + a.ocol + (ng >> 1);
                    if (c.vec && ng + 16 <= a.n) {
                        *reinterpret_cast<float4*>(dst) = make_float4(g[0], g[1], g[2], g[3]);
                        *reinterpret_cast<float4*>(dst + 4) = make_float4(g[4], g[5], g[6], g[7]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; j++) if (ng + 2 * j + 1 < a.n) dst[j] = g[j];
                    }
                } else if (c.vec && a.epi == EPI_STORE && ng + 16 <= a.n) {
                    // fast path: 128-bit bias / residual / accumulate / store
                    const float* ur = a.utab ? (a.utab + (long)__ldg(a.uidx + b) * a.utab_ld + ng) : nullptr;
                    float* dst = a.out + row * a.ldo + a.ocol + ng;
                    const float* rs = a.res ? (a.res + row * a.ldres + a.rescol + ng) : nullptr;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        if (a.bias) { const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + ng) + q); o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w; }
                        if (ur) { const float4 uu = __ldg(reinterpret_cast<const float4*>(ur) + q); o.x += uu.x; o.y += uu.y; o.z += uu.z; o.w += uu.w; }
                        if (rs) { const float4 rr = *(reinterpret_cast<const float4*>(rs) + q); o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w; }
                        if (a.out_act == ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        if (a.accumulate) { const float4 pp = *(reinterpret_cast<const float4*>(dst) + q); o.x += pp.x; o.y += pp.y; o.z += pp.z; o.w += pp.w; }
                        if (a.out_div != 1.f) { o.x = o.x / a.out_div; o.y = o.y / a.out_div; o.z = o.z / a.out_div; o.w = o.w / a.out_div; }
                        if (a.out_act == ACT_TANH) { o.x = tanhf(o.x); o.y = tanhf(o.y); o.z = tanhf(o.z); o.w = tanhf(o.w); }
                        *(reinterpret_cast<float4*>(dst) + q) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (ng + j < a.n) conv_epilogue_store(a, b, row, ng + j, v[j]);
                }
            }
            tc::tc_fence_before();   // TMEM reads done before the next tile's a_full arrive releases the MMA warp
        }
    } else if (warp == 4) {
        // ================= weight producer (one thread, cp.async.bulk ring) =================
        if (lane == 0) {
            uint32_t gp = 0;
            const uint32_t kc_bytes = (uint32_t)c.ntile * 16u;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                if (c.resident && tile != (int)blockIdx.x) break;
                for (int p = 0; p < c.npieces; p++, gp++) {
                    const int tap = p / c.cpt, cc = p - tap * c.cpt;
                    const int ch0 = cc * c.piece_ch;
                    const int nkc = min(c.piece_ch, a.cin - ch0) >> 3;
                    const uint32_t s = gp % (uint32_t)c.nstages;
                    if (!c.resident) tc::mbar_wait(bar_empty0 + 8u * s, ((gp / (uint32_t)c.nstages) & 1u) ^ 1u);
                    const uint32_t fb = bar_full0 + 8u * s;
                    tc::mbar_expect_tx(fb, kc_bytes * (uint32_t)nkc);
                    const uint32_t dst = tc::smem_u32(sW + (size_t)s * c.slot_bytes);
                    const __nv_bfloat16* src = a.wtc + (((long)tap * kc_total + (ch0 >> 3)) * a.npad16 + (long)ny * c.ntile) * 8;
                    for (int kc = 0; kc < nkc; kc++)
                        tc::bulk_g2s(dst + (uint32_t)kc * kc_bytes, src + (long)kc * a.npad16 * 8, kc_bytes, fb);
                }
            }
        }
    } else {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc(TC_M, c.ntile);
            const uint32_t sA_u = tc::smem_u32(sA), sW_u = tc::smem_u32(sW);
            uint32_t gp = 0, it = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
                tc::mbar_wait(bar_afull, it & 1);
                tc::tc_fence_after();
                uint32_t accum = 0;
                for (int p = 0; p < c.npieces; p++, gp++) {
                    const int tap = p / c.cpt, cc = p - tap * c.cpt;
                    const int ch0 = cc * c.piece_ch;
                    const int nk16 = min(c.piece_ch, a.cin - ch0) >> 4;
                    const uint32_t s = c.resident ? (uint32_t)p : (gp % (uint32_t)c.nstages);
                    if (!c.resident || it == 0) {
                        tc::mbar_wait(bar_full0 + 8u * s, c.resident ? 0u : ((gp / (uint32_t)c.nstages) & 1u));
                        tc::tc_fence_after();
                    }
                    const uint32_t arow = (uint32_t)(a.toff[tap] - c.min_off);
                    const uint32_t a0 = sA_u + (uint32_t)(ch0 >> 3) * lbo_a + arow * 16u;
                    const uint32_t b0 = sW_u + s * (uint32_t)c.slot_bytes;
                    for (int k = 0; k < nk16; k++) {
                        const uint64_t ad = tc::make_desc(a0 + (uint32_t)(2 * k) * lbo_a, lbo_a, 128u);
                        const uint64_t bd = tc::make_desc(b0 + (uint32_t)(2 * k) * lbo_b, lbo_b, 128u);
                        tc::umma_bf16(tmem_base, ad, bd, idesc, accum);
                        accum = 1;
                    }
                    if (!c.resident) tc::umma_commit(bar_empty0 + 8u * s);   // frees the weight slot when these MMAs retire
                }
                tc::umma_commit(bar_acc);                                     // accumulator ready for the epilogue
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem_base, (uint32_t)c.tmem_cols);
}

// ------------------------------------------------------------------------------------------ host side
static inline bool conv_tc_supported(const ConvArgs& a) {
    if (!a.wtc) return false;
    if (a.cin % 16 || a.n % 16 || a.cin < 16 || a.n < 16) return false;
    if (a.ldx % 4 || a.xcol % 4) return false;
    if (a.epi == EPI_SPLIT && (a.split % 16)) return false;
    return true;
}

static inline bool conv_tc_plan(const ConvArgs& a, TcCfg& c) {
    int mn = a.toff[0], mx = a.toff[0];
    for (int i = 1; i < a.ntaps; i++) { mn = a.toff[i] < mn ? a.toff[i] : mn; mx = a.toff[i] > mx ? a.toff[i] : mx; }
    c.min_off = mn;
    int rows = TC_M + (mx - mn);
    rows = ((rows + 7) / 8) * 8 + 1;           // odd row count: (kc + r) mod 8 spreads 16 B chunks over all banks
    c.rows_a = rows;
    c.a_bytes = ((a.cin / 8) * rows * 16 + 127) / 128 * 128;
    c.piece_ch = a.cin < TC_PIECE_CH ? a.cin : TC_PIECE_CH;
    c.cpt = (a.cin + c.piece_ch - 1) / c.piece_ch;
    c.npieces = a.ntaps * c.cpt;
    const int limit = 200 * 1024;
    int nt = a.npad16 <= 256 ? a.npad16 : 0;
    if (!nt) for (int cand = 256; cand >= 16; cand -= 16) if (a.npad16 % cand == 0) { nt = cand; break; }
    for (;;) {
        c.ntile = nt;
        c.slot_bytes = c.piece_ch * nt * 2;
        const int bar_bytes = (2 * TC_MAX_STAGES + 2) * 8 + 16;
        const long res_bytes = (long)c.npieces * c.slot_bytes;
        if (c.npieces <= TC_MAX_STAGES && res_bytes <= 72 * 1024 && c.a_bytes + res_bytes + bar_bytes <= limit) {
            c.resident = 1; c.nstages = c.npieces;
        } else {
            c.resident = 0; c.nstages = c.npieces < 4 ? c.npieces : 4;
        }
        c.smem_bytes = c.a_bytes + c.nstages * c.slot_bytes + (2 * c.nstages + 2) * 8 + 16;
        if (c.smem_bytes <= limit) break;
        // shrink the N tile (keeps divisibility of npad16)
        int next = 0;
        for (int cand = nt - 16; cand >= 16; cand -= 16) if (a.npad16 % cand == 0) { next = cand; break; }
        if (!next) return false;
        nt = next;
    }
    int tc_cols = 32; while (tc_cols < c.ntile) tc_cols <<= 1;
    c.tmem_cols = tc_cols;
    c.vec = (a.ldo % 4 == 0) && (a.ocol % 4 == 0) && (!a.res || (a.ldres % 4 == 0 && a.rescol % 4 == 0)) &&
            (!a.utab || a.utab_ld % 4 == 0);
    if (a.epi == EPI_GATE) c.vec = c.vec && (a.ocol % 4 == 0) && (a.ldo % 4 == 0);
    return true;
}

static inline cudaError_t conv_tc_launch(const ConvArgs& a, int num_sms, cudaStream_t st) {
    TcCfg c;
    if (!conv_tc_plan(a, c)) return cudaErrorInvalidConfiguration;
    static int max_set = 0;
    if (c.smem_bytes > max_set) {
        cudaError_t e = cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        max_set = 227 * 1024;
    }
    int occ = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_conv_tc, TC_THREADS, c.smem_bytes);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    const int tmem_occ = 512 / c.tmem_cols;
    if (occ > tmem_occ) occ = tmem_occ;
    const int ny = a.npad16 / c.ntile;
    int gx = num_sms * occ / (ny < 1 ? 1 : 1);
    if (gx > a.ntiles) gx = a.ntiles;
    if (gx < 1) gx = 1;
    dim3 grid(gx, ny);
    k_conv_tc<<<grid, TC_THREADS, c.smem_bytes, st>>>(a, c);
    return cudaGetLastError();
}
