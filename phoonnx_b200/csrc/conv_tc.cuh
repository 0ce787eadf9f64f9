// tcgen05 / TMEM implicit-GEMM conv1d for sm_100a (bf16 operands, fp32 accumulate in TMEM).
//
// Mapping (DESIGN.md "conv_tc"):  D[M = 128 time rows, N = C_out tile] += A[M, K = C_in] * B[N, K]^T
// once per tap; a dilated / polyphase tap is the SAME shared-memory activation tile read at a
// shifted row, which with the K-major no-swizzle canonical layout
//     addr(row, k) = (k / 8) * LBO + row * 16 B + (k % 8) * 2 B        (SBO = 128 B)
// is just a 16-byte-aligned start-address offset in the UMMA smem descriptor -- no im2col
// copy and no re-fetch per tap.
//   * activations: fp32 channel-last rows -> (leaky-relu) -> bf16 (or bf16 hi / lo planes for the fp32-faithful bf16x3 product)
//     -> smem by the 6 loader warps, or rows a producer already wrote as bf16 MMA operands via cp.async; zero-filled outside
//     the utterance;
//   * weights: pre-packed bf16 [tap][C_in/8][N][8] (N-tiled copies for N > 256) streamed by one producer thread, one
//     cp.async.bulk (SASS UBLKCP) per <= 64-channel piece through an mbarrier ring, or kept resident in smem for the whole
//     persistent CTA when they fit; optional 2-CTA cluster mode multicasts every piece to a CTA pair;
//   * MMAs: one elected thread, tcgen05.mma.cta_group::1.kind::f16 (SASS UTCHMMA), M=128, N<=256, K=16; K slices (with their
//     own tap lists) accumulate in TMEM; completion via tcgen05.commit -> mbarrier;
//   * epilogue (8 warps, one kernel instantiation per variant): tcgen05.ld 32x32b (SASS LDTM) -> bias / speaker bias /
//     residual(s) / gate / split-accumulate -> smem transpose -> coalesced fp32 or bf16-operand-row stores.
// What bounds it: the SM's 128 B/clk shared-memory port, shared by the MMA operand reads, the loader and the
// epilogue transpose.
#pragma once
#include <cuda.h>          // CUtensorMap (type + enums; the encoder is fetched from the driver at run time)
#include <type_traits>
#include "common.cuh"
#include "kernels_f32.cuh"

#ifndef TC_DESC_SWAP_LBO_SBO
#define TC_DESC_SWAP_LBO_SBO 0
#endif

#ifndef TC_EXPERIMENT_SKIP_RESB
#define TC_EXPERIMENT_SKIP_RESB 0      // timing experiment only (wrong results): skip the bf16 residual prefetch
#endif
#define TC_M 128
#define TC_LOAD_BATCH 9          // fp32 loader: (row, chunk) items whose 2 x LDG.128 are in flight per thread before any conversion.  9 x 192
                                 // threads cover the 130 x 12 items of a k = 3, 96-channel K slice in ONE memory round trip (8 left 24 items
                                 // to a second one; the 128 x 12 items of a 1x1 slice need exactly 8 once the unused pad row is skipped)
#define TC_EPI_WARPS 8
#define TC_NFIXBAR 10               // fixed mbarriers behind the weight ring's (a_full, a_empty, acc_full, acc_empty, a_tma: two each)
#define TC_LOAD_WARPS 6
#define TC_THREADS ((TC_EPI_WARPS + TC_LOAD_WARPS + 2) * 32)
#define TC_LOAD_THREADS (TC_LOAD_WARPS * 32)
#define TC_EPI_PITCH 36            // floats per staged row: 16 B aligned, conflict-free for 8-lane phases
#define TC_PIECE_CH 64
#define TC_MAX_STAGES 24
#define TC_SPIN_LIMIT (1u << 24)

struct TcCfg {
    int ntile;        // N columns per CTA (multiple of 16, <= 256)
    int tmem_cols;    // power of two >= 32
    int rows_a;       // activation rows held in smem (odd: conflict-free 16 B chunk scatter)
    int rows_need;    // rows the MMAs actually read: 128 + (max tap offset - min tap offset); the rest of rows_a is padding, never loaded
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
    int nabuf;        // activation tile buffers (1 or 2)
    int naccbuf;      // TMEM accumulator buffers (1 or 2)
    int epi_off;      // byte offset of the epilogue transpose buffers (4 warps x 32 x TC_EPI_PITCH floats)
    int nepi;         // epilogue warps: 8, or 12 when the loader is the cp.async one (bf16 operand rows: 2 loader warps suffice)
    int nload;        // loader warps = 14 - nepi
    int gw;           // accumulator columns per epilogue work item (32, or 16 when that spreads the items better over the warps)
    int tma, nboxes, box_rows;   // activation tile by TMA (bf16 operand rows): per 8-channel plane `nboxes` boxes of [box_rows x 16 bytes]
    int cluster;      // 2: CTA pairs (cluster 2x1x1) share every weight piece -- each CTA fetches every other piece and multicasts it to both
                      //    (the ring-mode launches are bound by L2 -> SM weight traffic: a 128-row tile re-streams all taps); 1: off
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
    // try_wait with a suspend-time hint: the warp sleeps in hardware until the phase flips (or the hint expires)
    // instead of spinning -- a hot spin loop in 8 waiting warps starves the single MMA-issuing thread of issue slots
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU.
// TC_WAIT_POLL (round 2, default): poll with the non-blocking test_wait and a short back-off.  The try_wait form with a
// suspend-time hint compiles to SYNCS...TRYWAIT + NANOSLEEP.SYNCS; the r02c timeline of the fused MRF kernel showed the sleeping
// warp resuming ~1000 cycles AFTER the last arrival on the barrier (x1 staged at 31 868, the MMA warp past its wait at 32 938) --
// at eight such hand-offs per 19k-cycle tile that is the tensor pipe's idle time.  A poll costs one SYNCS.PHASECHK per ~100 cycles
// per waiting warp: noise next to the epilogue warps' own instruction stream.
#ifndef TC_WAIT_POLL
#define TC_WAIT_POLL 0
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
#if TC_WAIT_POLL
    while (!mbar_test(bar, parity)) {
#if TC_WAIT_POLL == 1
        if (++spins > (1u << 20)) __nanosleep(256);       // ~50 ms of pure polling first: __nanosleep's real granularity is of the order of a microsecond
#else
        if (++spins > 4u) __nanosleep(20);
#endif
        if (spins > (TC_SPIN_LIMIT << 2)) __trap();
    }
#else
    while (!mbar_try(bar, parity)) {
        if (++spins > 64u) __nanosleep(spins > 4096u ? 256 : 32);
        if (spins > TC_SPIN_LIMIT) __trap();
    }
#endif
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// ---- 2-CTA cluster helpers (weight multicast)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one bulk copy delivered to the same shared-memory offset (and signalling the mbarrier at the same offset) in every CTA of `mask`
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
// tcgen05.commit arriving on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
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
// tcgen05.mma with the two shared-memory descriptors given as (lo, hi) 32-bit halves: only `lo` (start address >> 4 | LBO << 16)
// changes between the MMAs of a conv, by plain 32-bit adds; `hi` (SBO, version) is loop-invariant
__device__ __forceinline__ void umma_bf16_lh(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accum) : "memory");
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
// 32 consecutive accumulator columns of this thread's TMEM lane: one instruction, one wait
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
// one lane of a converged warp; ptxas knows the predicate is single-lane, so UTCHMMA / UBLKCP / UTCBAR inside the
// elected region take their uniform-register operands directly (a plain `lane == 0` test makes it wrap EVERY such
// instruction in an ELECT / BRA.U.ANY waterfall loop plus R2UR moves: ~100 cycles per MMA issued, measured)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xFFFFFFFFu));
    return pred != 0;
}
// ---- warp-uniform issue path (round 2).  The whole issuing warp runs the loop, converged; only the asynchronous instruction itself
// is predicated on the elect.sync flag.  With the loop inside a single-lane branch (round 1) every descriptor lived in a vector
// register and each UTCHMMA cost 4 R2UR moves plus vector adds (~95 cycles per MMA issued against 40 executed); in converged code
// ptxas keeps descriptors, ring slots and barrier addresses in uniform registers and emits UIADD3 / UTCHMMA / UTCBAR / UBLKCP
// back to back (the elect predicate folds away: a uniform-datapath instruction executes once per warp).
__device__ __forceinline__ uint32_t elect_flag() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xFFFFFFFFu));
    return pred;
}
__device__ __forceinline__ void umma_bf16_e(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum, uint32_t el) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(el) : "memory");
}
__device__ __forceinline__ void umma_bf16_lh_e(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accum, uint32_t el) {
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %7, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accum), "r"(el) : "memory");
}
__device__ __forceinline__ void umma_commit_e(uint32_t bar, uint32_t el) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(el) : "memory");
}
__device__ __forceinline__ void umma_commit_mc_e(uint32_t bar, uint16_t mask, uint32_t el) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
                 ::"r"(bar), "h"(mask), "r"(el) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_e(uint32_t bar, uint32_t bytes, uint32_t el) {
    asm volatile("{\n\t.reg .b64 st;\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes), "r"(el) : "memory");
}
__device__ __forceinline__ void bulk_g2s_e(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint32_t el) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "r"(el) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc_e(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask, uint32_t el) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n\t}"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask), "r"(el) : "memory");
}
// TMA tile load (cp.async.bulk.tensor, SASS UTMALDG): one box of a 2-D tensor map -> shared memory, completion counted in bytes on `bar`.
// Coordinates are signed; rows outside the tensor arrive as zeros.
__device__ __forceinline__ void tma_load_2d_e(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar, uint32_t el) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "r"(el) : "memory");
}
// a value every lane of the (converged) warp holds identically, in a form ptxas can keep in a uniform register
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&p);
}
// the two bf16 halves of a packed word, widened back to fp32 (exact)
__device__ __forceinline__ float bf16_lo_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

}  // namespace tc

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Warp-specialised persistent kernel (TC_THREADS = 512 = 16 warps: the register file is handed out to groups of 4 warps, so 18
// warps would be charged as 20 and cap the kernel at 96 registers per thread; 16 warps leave 128 for the epilogue's prefetch):
//   warps 0-7   epilogue  (TMEM -> registers -> smem transpose -> coalesced global), TMEM lane quadrant = warp % 4,
//               the two warps of a quadrant alternate over the 32-column groups of the accumulator
//   warps 8-13  activation loaders (global fp32 -> lrelu -> bf16 -> smem operand tile, or cp.async of bf16 operand rows)
//   warp  14    weight producer (elected lane) + TMEM allocator
//   warp  15    MMA issuer (elected lane)
// Activation tiles and TMEM accumulators are double-buffered when they fit, so the load of tile i+1, the MMAs of
// tile i and the epilogue of tile i-1 overlap inside one CTA.
// r01 timeline (profiles/r01c_conv_timeline.log): with 4 epilogue warps and one generic, branchy epilogue the
// TMEM -> HBM drain of one 128 x 192 tile took 40-80k cycles (2100 warp-instructions per 32 x 32 sub-tile at one warp
// per scheduler) against 6-12k for the loaders and 1-9k for the MMAs -- hence 8 + 8 warps, a lean epilogue
// specialised per mode at compile time, and batched (independent) residual / accumulate loads.
// one thread per 128-row tile: {first row of the utterance, its rows, first row of the tile inside it, utterance}.  The
// kernel's roles read ONE 16-byte descriptor per tile, a tile ahead, instead of a binary search of dependent global loads
// per role per tile (r01d ncu: 18 % of all stall samples sat on that search and the geometry loads behind it)
__global__ void k_tile_desc(const int* __restrict__ cu, const int* __restrict__ tile_cu, int B, int rate, int ntiles, int tm,
                            int4* __restrict__ out) {
    pdl_enter();
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= ntiles) return;
    const int b = find_segment(tile_cu, B, tile);
    const int cb0 = __ldg(cu + b), cb1 = __ldg(cu + b + 1);
    out[tile] = make_int4(cb0 * rate, (cb1 - cb0) * rate, (tile - __ldg(tile_cu + b)) * tm, b);
}

// [rows][n16][8] -> [n16 / wt][rows][wt][8]  (rows = tap x 8-channel chunk): the N-tiled weight copy, built once at load time
__global__ void k_retile_weights(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long rows, int n16, int wt) {
    pdl_enter();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte (8-element) unit per thread
    if (i >= rows * n16) return;
    const long r = i / n16; const int n = (int)(i - r * n16);
    const int ty = n / wt, nn = n - ty * wt;
    reinterpret_cast<uint4*>(out)[((long)ty * rows + r) * wt + nn] = reinterpret_cast<const uint4*>(in)[i];
}

#define TC_DBG_TILES 16
#define TC_STAMP(it_, slot_) do { if (dbg_on && (it_) < TC_DBG_TILES) a.dbg[(it_) * 16 + (slot_)] = (unsigned long long)clock64(); } while (0)

// RESK: residual operand kind (0 none, 1 fp32 rows, 2 bf16 lrelu rows); ACC: the destination may be accumulated into.  Compile-time so
// that the prefetch registers exist only for operands the launch really has.
template <int EPI, int RESK, int ACC>
__device__ __forceinline__ void tc_epilogue(const ConvArgs& a, const TcCfg& c, uint8_t* smem, uint32_t tmem_base,
                                            uint32_t bar_accfull0, uint32_t bar_accempty0, int warp, int lane, bool dbg_on) {
    const int q = warp & 3, hh = warp >> 2, ny = blockIdx.y;
    const int nhh = c.nepi >> 2;                       // warps per TMEM lane quadrant: they take the column groups round-robin
    float* sE = reinterpret_cast<float*>(smem + c.epi_off) + warp * (32 * TC_EPI_PITCH);
    float* srow = sE + lane * TC_EPI_PITCH;
    const int ngroups = (c.ntile + c.gw - 1) / c.gw;
    const float inv_div = 1.f / a.out_div;
    constexpr bool has_res = (RESK == 1), has_resb = (RESK == 2 || RESK == 3);
    constexpr int NRB = (RESK == 3) ? 3 : 1;          // bf16 residual row groups summed (RESK 3: up to three, a.nresb)
    const float resb_inv = has_resb ? 1.f / a.resb_slope : 1.f;
    const float2 resb_inv2 = make_float2(resb_inv, resb_inv), inv_div2 = make_float2(inv_div, inv_div), outb_sl2 = make_float2(a.outb_slope, a.outb_slope);
    const bool scaled = a.out_div != 1.f;
    // gate outputs straight to bf16 operand rows (lrelu slope 1 = identity; 16-byte aligned rows)
    const bool gate_direct = (EPI == EPI_GATE) && a.outb != nullptr && a.outb_slope == 1.f && (a.ldo % 8 == 0) && (a.ocol % 8 == 0);
    const float* sBias = reinterpret_cast<const float*>(smem + c.epi_off) + c.nepi * (32 * TC_EPI_PITCH);   // this N tile's bias
    // (round 2, measured and removed: storing bf16 operand rows straight from registers -- thread = row, 16 contiguous bytes per 8
    // channels, residual rows read the same way, no shared-memory transpose -- is bit-identical and SLOWER: medium decoder 182.5 vs
    // 179.2 ms per step, `high` 161.1 vs 159.5 (profiles/r02z_ab_conv_direct_epilogue*.log).  A warp-wide 16-byte access that touches
    // 32 different 128-byte lines costs the LSU 32 wavefronts; the transpose's coalesced 8-byte stores cost 2-4.  The gate epilogue
    // keeps its direct stores: there the alternative was a second phase, not a cheaper one.)
    uint32_t it = 0;
    int4 dnext = ((int)blockIdx.x < a.ntiles) ? __ldg(a.tdesc + blockIdx.x) : make_int4(0, 0, 0, 0);
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
        TC_STAMP(it, 0);
        // tile geometry: one 16-byte descriptor, fetched a tile ahead (k_tile_desc)
        const int4 dsc = dnext;
        if (tile + (int)gridDim.x < a.ntiles) dnext = __ldg(a.tdesc + tile + gridDim.x);
        const int b = dsc.w, t0 = dsc.z, len = dsc.y;
        const long row0 = (long)dsc.x;
        const float* ur = a.utab ? (a.utab + (long)__ldg(a.uidx + b) * a.utab_ld) : nullptr;
        TC_STAMP(it, 1);
        const uint32_t cbuf = it % (uint32_t)c.naccbuf, cuse = it / (uint32_t)c.naccbuf;
        tc::mbar_wait(bar_accfull0 + 8u * cbuf, cuse & 1u);
        TC_STAMP(it, 2);
        tc::tc_fence_after();
        const int trow0 = t0 + q * 32;                    // first time row of this warp
        const int nrows = len - trow0;                    // rows of this warp inside the utterance (<= 0: nothing to store)
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + cbuf * (uint32_t)c.ntile;
        long long ph1 = 0, ph2 = 0, tq = 0, ph1a = 0, ph1b = 0;
        if (nrows > 0) {                                  // warp-uniform
            for (int g = hh; g < ngroups; g += nhh) {
                if (dbg_on) tq = clock64();
                const int n0 = g * c.gw;
                const int ng = ny * c.ntile + n0;         // global accumulator column of this group
                const int ncols = min(c.gw, c.ntile - n0);  // 16 or 32
                // phase-2 mapping: `lpr` lanes cover one staged row (128-bit each), 32 / lpr rows per warp instruction
                const int ocols = (EPI == EPI_GATE) ? (ncols >> 1) : ncols;      // 8, 16 or 32 staged output columns
                const int og = (EPI == EPI_GATE) ? (ng >> 1) : ng;               // first output column
                // the body is instantiated per lanes-per-row shift so that every loop bound, predicate and row stride below is a
                // compile-time constant (r01d: ~18k warp-instructions per 128 x 128 tile with the shift as a run-time value)
                auto group_body = [&](auto lsh_c) {
                constexpr int lsh = decltype(lsh_c)::value;
                constexpr int lpr = 1 << lsh, rpi = 32 >> lsh;                   // lanes per row, rows per instruction; iterations = lpr
                const int cq = (lane & (lpr - 1)) * 4, rsub = lane >> lsh;
                const int n = og + cq;
                float* dst0; long ldd; int acc;
                if (EPI == EPI_SPLIT && n >= a.split) { dst0 = a.out2 + a.ocol2 + (n - a.split); ldd = a.ldo2; acc = a.accumulate2; }
                else { dst0 = a.out + a.ocol + n; ldd = a.ldo; acc = a.accumulate; }
                dst0 += (row0 + trow0 + rsub) * ldd;
                const float* res0 = has_res ? (a.res + (row0 + trow0 + rsub) * a.ldres + a.rescol + n) : nullptr;
                const __nv_bfloat16* resb0 = has_resb ? (a.resb + (row0 + trow0 + rsub) * (long)a.ldresb + a.rescol + n) : nullptr;
                __nv_bfloat16* outb0 = a.outb ? (a.outb + (row0 + trow0 + rsub) * (long)a.ldo + a.ocol + n) : nullptr;
                (void)res0; (void)resb0; (void)outb0;
                const float* st0 = sE + rsub * TC_EPI_PITCH + cq;
                // ---- prefetch: every global operand of this group's phase-2 iterations is independent of the accumulator; issuing the
                // loads HERE puts their latency under phase 1 (r01 timeline: with the loads inside phase 2 a 128 x 128 tile spent
                // 12-15k cycles in four serial load -> use rounds, against 2k for a store-only epilogue)
                float4 rr[has_res ? 8 : 1], pp[ACC ? 8 : 1];
                uint2 rb[NRB][has_resb ? 8 : 1];
#pragma unroll
                for (int itr = 0; itr < 8; itr++) {
                    if (itr >= lpr) continue;                  // compile-time
                    const int rl = itr * rpi + rsub;
                    const bool okp = rl < nrows;
                    if (has_res) { rr[itr] = make_float4(0.f, 0.f, 0.f, 0.f); if (okp) rr[itr] = *reinterpret_cast<const float4*>(res0 + (long)(itr * rpi) * a.ldres); }
                    if (has_resb) {
#pragma unroll
                        for (int j = 0; j < NRB; j++) {
                            rb[j][itr] = make_uint2(0u, 0u);
                            if (okp && (NRB == 1 || j < a.nresb) && !TC_EXPERIMENT_SKIP_RESB) rb[j][itr] = *reinterpret_cast<const uint2*>(resb0 + (long)(itr * rpi) * a.ldresb + j * a.resb_stride);
                        }
                    }
                    if (ACC) { pp[itr] = make_float4(0.f, 0.f, 0.f, 0.f); if (okp && acc) pp[itr] = *reinterpret_cast<const float4*>(dst0 + (long)(itr * rpi) * ldd); }
                }
                // ---- phase 1 (thread = TMEM lane = time row): TMEM -> registers, bias / speaker bias / gate -> transpose buffer
                // one 32-column TMEM load per group where the registers allow (fp32 residual + accumulate operands already
                // hold 64 prefetch registers: those variants load 16 columns at a time)
                if (dbg_on) { const long long t2 = clock64(); ph1a += t2 - tq; }
                constexpr bool WIDE = !ACC && RESK != 3;
                float vv[WIDE ? 32 : 16];
                if (WIDE) {
                    if (ncols == 32) tc::tmem_ld32(trow + (uint32_t)n0, vv);      // warp-uniform
                    else tc::tmem_ld16(trow + (uint32_t)n0, vv);
                }
                if (dbg_on) { const long long t2 = clock64(); ph1b += t2 - tq; }
#pragma unroll
                for (int hcol = 0; hcol < 32; hcol += 16) {
                    if (hcol < ncols) {
                        float* v = WIDE ? vv + hcol : vv;
                        if (!WIDE) tc::tmem_ld16(trow + (uint32_t)(n0 + hcol), vv);
                        const int nc = ng + hcol;
                        {
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const float4 bb = *reinterpret_cast<const float4*>(sBias + n0 + hcol + 4 * k);
                                const float2 s0 = __fadd2_rn(make_float2(v[4 * k], v[4 * k + 1]), make_float2(bb.x, bb.y));
                                const float2 s1 = __fadd2_rn(make_float2(v[4 * k + 2], v[4 * k + 3]), make_float2(bb.z, bb.w));
                                v[4 * k] = s0.x; v[4 * k + 1] = s0.y; v[4 * k + 2] = s1.x; v[4 * k + 3] = s1.y;
                            }
                        }
                        if (ur) {
#pragma unroll
                            for (int k = 0; k < 4; k++) { const float4 uu = __ldg(reinterpret_cast<const float4*>(ur + nc) + k); v[4 * k] += uu.x; v[4 * k + 1] += uu.y; v[4 * k + 2] += uu.z; v[4 * k + 3] += uu.w; }
                        }
                        if (EPI == EPI_GATE) {
                            // interleaved (tanh-arg, sigmoid-arg) pairs (commons.py:99-106); MUFU tanh: the gate output is
                            // rounded to bf16 by the next conv's loader, far coarser than tanh.approx's 2^-11
                            float gt[8];
#pragma unroll
                            for (int j = 0; j < 8; j++) gt[j] = tanh_fast(v[2 * j]) * fmaf(0.5f, tanh_fast(0.5f * v[2 * j + 1]), 0.5f);
                            if (gate_direct) {
                                // bf16 operand rows for the next GEMM: this thread's row gets 8 gate outputs = 16 contiguous bytes; 32 lanes fill
                                // 32 full 32-byte sectors per pair of stores.  No transpose, no second phase (the staged path's
                                // STS -> syncwarp -> LDS -> STG chain was ~40 % of this epilogue's latency).
                                if (lane < nrows) {
                                    __nv_bfloat16* orow = a.outb + (row0 + trow0 + lane) * (long)a.ldo + a.ocol + og + (hcol >> 1);
                                    *reinterpret_cast<uint4*>(orow) = make_uint4(tc::pack_bf16(gt[0], gt[1]), tc::pack_bf16(gt[2], gt[3]),
                                                                                 tc::pack_bf16(gt[4], gt[5]), tc::pack_bf16(gt[6], gt[7]));
                                }
                            } else {
                            *reinterpret_cast<float4*>(srow + (hcol >> 1)) = make_float4(gt[0], gt[1], gt[2], gt[3]);
                            *reinterpret_cast<float4*>(srow + (hcol >> 1) + 4) = make_float4(gt[4], gt[5], gt[6], gt[7]);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                *reinterpret_cast<float4*>(srow + hcol + 4 * k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                        }
                    }
                }
                if (gate_direct) { if (dbg_on) { const long long t2 = clock64(); ph1 += t2 - tq; tq = t2; } return; }
                __syncwarp();
                if (dbg_on) { const long long t2 = clock64(); ph1 += t2 - tq; tq = t2; }
                // ---- phase 2: staged rows + prefetched residual / accumulate operands -> global
                // branch-free per iteration (predicated stores only): the 8 iterations are independent and must overlap -- with an
                // early `continue` per iteration each LDS -> math -> STG chain ran alone (270 cycles per iteration, measured)
#pragma unroll
                for (int h2 = 0; h2 < 2; h2++) {
                    if (h2 * 4 >= lpr) continue;               // compile-time
                    float4 o[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int itr = h2 * 4 + u;
                        if (itr < lpr) o[u] = *reinterpret_cast<const float4*>(st0 + itr * rpi * TC_EPI_PITCH);
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int itr = h2 * 4 + u;
                        if (itr >= lpr) continue;              // compile-time
                        const bool okp = itr * rpi + rsub < nrows;
                        // packed fp32 pairs (FADD2 / FMUL2 on sm_100): the epilogue is issue-bound, every pair op saves a slot
                        float2 va = make_float2(o[u].x, o[u].y), vb = make_float2(o[u].z, o[u].w);
                        float2 ra = make_float2(0.f, 0.f), rb2 = ra;
                        if (has_res) { ra = make_float2(rr[itr].x, rr[itr].y); rb2 = make_float2(rr[itr].z, rr[itr].w); }
                        if (has_resb) {
                            // the residual is the conv's own input, stored as the bf16 lrelu operand: invert the (monotone) lrelu
#pragma unroll
                            for (int j = 0; j < NRB; j++) {
                                const uint32_t bx = rb[j][itr].x, by = rb[j][itr].y;
                                const float2 la = make_float2(tc::bf16_lo_f(bx), tc::bf16_hi_f(bx)), lb = make_float2(tc::bf16_lo_f(by), tc::bf16_hi_f(by));
                                const float2 ia = __fmul2_rn(la, resb_inv2), ib = __fmul2_rn(lb, resb_inv2);
                                ra = __fadd2_rn(ra, make_float2(fminf(la.x, ia.x), fminf(la.y, ia.y)));
                                rb2 = __fadd2_rn(rb2, make_float2(fminf(lb.x, ib.x), fminf(lb.y, ib.y)));
                            }
                        }
                        if (EPI == EPI_SUBFROM) {
                            va = make_float2(ra.x - va.x, ra.y - va.y); vb = make_float2(rb2.x - vb.x, rb2.y - vb.y);
                        } else if (EPI != EPI_GATE) {
                            va = __fadd2_rn(va, ra); vb = __fadd2_rn(vb, rb2);
                            if (a.out_act == ACT_RELU) { va.x = fmaxf(va.x, 0.f); va.y = fmaxf(va.y, 0.f); vb.x = fmaxf(vb.x, 0.f); vb.y = fmaxf(vb.y, 0.f); }
                            if (ACC) { va = __fadd2_rn(va, make_float2(pp[itr].x, pp[itr].y)); vb = __fadd2_rn(vb, make_float2(pp[itr].z, pp[itr].w)); }
                            if (scaled) { va = __fmul2_rn(va, inv_div2); vb = __fmul2_rn(vb, inv_div2); }      // warp-uniform
                        }
                        if ((EPI == EPI_STORE || EPI == EPI_GATE) && a.outb) {
                            // the consumer's MMA operand: bf16(lrelu(v)), 8 bytes per lane
                            const float2 sa = __fmul2_rn(va, outb_sl2), sb = __fmul2_rn(vb, outb_sl2);
                            const uint2 pk2 = make_uint2(tc::pack_bf16(fmaxf(va.x, sa.x), fmaxf(va.y, sa.y)),
                                                         tc::pack_bf16(fmaxf(vb.x, sb.x), fmaxf(vb.y, sb.y)));
                            if (okp) *reinterpret_cast<uint2*>(outb0 + (long)(itr * rpi) * a.ldo) = pk2;
                        } else if (okp)
                            *reinterpret_cast<float4*>(dst0 + (long)(itr * rpi) * ldd) = make_float4(va.x, va.y, vb.x, vb.y);
                    }
                }
                __syncwarp();
                };
                if (ocols == 32) group_body(std::integral_constant<int, 3>{});
                else if (ocols == 16) group_body(std::integral_constant<int, 2>{});
                else group_body(std::integral_constant<int, 1>{});
                if (dbg_on) ph2 += clock64() - tq;
            }
        }
        if (dbg_on && it < TC_DBG_TILES) { a.dbg[it * 16 + 13] = (unsigned long long)ph1 | ((unsigned long long)ph1a << 20) | ((unsigned long long)ph1b << 40); a.dbg[it * 16 + 14] = (unsigned long long)ph2; }
        tc::tc_fence_before();                       // TMEM reads retired before the accumulator is handed back
        tc::mbar_arrive(bar_accempty0 + 8u * cbuf);
        TC_STAMP(it, 3);
    }
}

// scalar epilogue for layouts that cannot use 128-bit accesses (never on the production shapes; kept for generality)
__device__ __forceinline__ void tc_epilogue_scalar(const ConvArgs& a, const TcCfg& c, uint32_t tmem_base, uint32_t bar_accfull0,
                                                   uint32_t bar_accempty0, int warp, int lane) {
    const int q = warp & 3, hh = warp >> 2, ny = blockIdx.y;
    const int nhh = c.nepi >> 2;
    const int ngroups = (c.ntile + 15) >> 4;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
        const int4 dsc = __ldg(a.tdesc + tile);
        const int b = dsc.w, t0 = dsc.z, len = dsc.y;
        const long row0 = (long)dsc.x;
        const uint32_t cbuf = it % (uint32_t)c.naccbuf, cuse = it / (uint32_t)c.naccbuf;
        tc::mbar_wait(bar_accfull0 + 8u * cbuf, cuse & 1u);
        tc::tc_fence_after();
        const int t = t0 + q * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + cbuf * (uint32_t)c.ntile;
        for (int g = hh; g < ngroups; g += nhh) {
            float v[16];
            tc::tmem_ld16(trow + (uint32_t)(g * 16), v);
            if (t < len) {
                const int nc = ny * c.ntile + g * 16;
                if (a.epi == EPI_GATE) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int n2 = nc + 2 * j;
                        float va = v[2 * j], vb = v[2 * j + 1];
                        if (a.bias) { va += __ldg(a.bias + n2); vb += __ldg(a.bias + n2 + 1); }
                        if (a.utab) { const float* ur = a.utab + (long)__ldg(a.uidx + b) * a.utab_ld; va += __ldg(ur + n2); vb += __ldg(ur + n2 + 1); }
                        if (n2 < a.n) a.out[(row0 + t) * a.ldo + a.ocol + (n2 >> 1)] = tanhf(va) * (1.f / (1.f + __expf(-vb)));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j++) if (nc + j < a.n) conv_epilogue_store(a, b, row0 + t, nc + j, v[j]);
                }
            }
        }
        tc::tc_fence_before();
        tc::mbar_arrive(bar_accempty0 + 8u * cbuf);
    }
}

// One instantiation per epilogue variant (EPI < 0: the scalar epilogue): each gets its own register allocation and a hot loop
// that fits the instruction cache (as ONE kernel with a run-time dispatch this was 18k SASS instructions, spilling for the
// heaviest variant's sake and stalling on instruction fetch -- r01d ncu `no_inst`)
template <int EPI, int RESK, int ACC>
__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_tc(const ConvArgs a, const TcCfg c, const __grid_constant__ CUtensorMap tmap) {
    // PDL: the prologue below (barriers, bias, tensor-memory allocation) and the whole weight ring read constants only, so they run
    // while the preceding grid is still finishing; the epilogue and loader warps -- the ones that touch activations, tile descriptors
    // or outputs -- call pdl_wait() first.  The MMA warp only follows their barriers.
    pdl_trigger();
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sW = smem + (size_t)c.nabuf * c.a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)c.nstages * c.slot_bytes);
    // bars: w_full[nstages], w_empty[nstages], a_full[2], a_empty[2], acc_full[2], acc_empty[2], a_tma[2]; then tmem ptr
    const uint32_t bar_full0 = tc::smem_u32(bars);
    const uint32_t bar_empty0 = bar_full0 + 8u * c.nstages;
    const uint32_t bar_afull0 = bar_empty0 + 8u * c.nstages;
    const uint32_t bar_aempty0 = bar_afull0 + 16u;
    const uint32_t bar_accfull0 = bar_aempty0 + 16u;
    const uint32_t bar_accempty0 = bar_accfull0 + 16u;
    const uint32_t bar_tma0 = bar_accempty0 + 16u;               // a_tma[2]: activation tile landed (TMA bytes; the loader warp waits)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * c.nstages + TC_NFIXBAR);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform in a form ptxas can see (role branches stay converged)
    const int ny = blockIdx.y;
    const bool dbg_on_cta = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
    const bool dbg_on = dbg_on_cta && lane == 0 &&
                        (warp == 0 || warp == c.nepi || warp >= TC_EPI_WARPS + TC_LOAD_WARPS);
    if (dbg_on && warp == 0) a.dbg[15] = (unsigned long long)clock64();

    if (tid == 0) {
        for (int s = 0; s < c.nstages; s++) { tc::mbar_init(bar_full0 + 8u * s, 1); tc::mbar_init(bar_empty0 + 8u * s, (uint32_t)c.cluster); }
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(bar_afull0 + 8u * i, c.tma ? 32u : (uint32_t)c.nload * 32u);
            tc::mbar_init(bar_tma0 + 8u * i, 1);
            tc::mbar_init(bar_aempty0 + 8u * i, 1);
            tc::mbar_init(bar_accfull0 + 8u * i, 1);
            tc::mbar_init(bar_accempty0 + 8u * i, (uint32_t)c.nepi * 32u);
        }
        tc::fence_mbar_init();
    }
    {
        float* sBias = reinterpret_cast<float*>(smem + c.epi_off) + c.nepi * (32 * TC_EPI_PITCH);
        for (int i = tid; i < c.ntile; i += TC_THREADS) sBias[i] = a.bias ? __ldg(a.bias + ny * c.ntile + i) : 0.f;
    }
    if (warp == TC_EPI_WARPS + TC_LOAD_WARPS) tc::tmem_alloc(tc::smem_u32(tmem_slot), (uint32_t)c.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    if (c.cluster == 2) tc::cluster_sync_all();         // the peer's barriers exist before anything is multicast to them
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // cluster mode: both CTAs of a pair run the ring for the same number of tile iterations (the even CTA's count); a CTA without
    // a tile in the last iteration still fetches its share of the pieces and releases the slots
    const int cl_rank = (c.cluster == 2) ? (int)tc::cluster_ctarank() : 0;
    const int ring_iters = (c.cluster == 2) ? (a.ntiles - ((int)blockIdx.x & ~1) + (int)gridDim.x - 1) / (int)gridDim.x
                                            : (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    const int kc_total = a.cin >> 3;                    // 16-byte chunks per activation row (per plane)
    // split3: weights hold [wh | wl | wh] per tap; the ring streams only (wh, wl) -- the wh piece serves both xh*wh and xl*wh (the
    // third copy is never fetched: the text-side GEMMs are bound by L2 -> SM weight traffic, r01g timeline: 2.65 MB per 128-row tile)
    const int nseg = a.split3 ? 2 : 1;
    const uint32_t lbo_a = (uint32_t)c.rows_a * 16u;
    const uint32_t lbo_b = (uint32_t)c.ntile * 16u;

    if (warp < c.nepi) {
        // ================= epilogue (8 warps; 12 when the loader is the cp.async one) =================
        pdl_wait();
        if constexpr (EPI < 0) tc_epilogue_scalar(a, c, tmem_base, bar_accfull0, bar_accempty0, warp, lane);
        else tc_epilogue<EPI, RESK, ACC>(a, c, smem, tmem_base, bar_accfull0, bar_accempty0, warp, lane, dbg_on);
    } else if (warp < TC_EPI_WARPS + TC_LOAD_WARPS) {
        // ================= activation loaders (the warps between the epilogue warps and warp 14) =================
        pdl_wait();
        if (c.tma) {
            // ---- TMA loader (bf16 operand rows): the first loader warp alone.  An 8-channel plane of the K-major no-swizzle tile is a
            // [rows_a x 16 bytes] box of the 2-D tensor [rows, ldxb]: cin/8 x nboxes cp.async.bulk.tensor per (tile, K slice), issued by
            // one lane, replace rows_a * cin/8 LDGSTS spread over the loader warps -- whose stream the r01g timeline showed pacing the
            // epilogue's own loads through the LSU queue.  Rows outside the array arrive as zeros; rows of neighbouring utterances
            // (first / last tile of an utterance) are zeroed here afterwards.
            if (warp == c.nepi) {
                const uint32_t el = tc::elect_flag();
                uint32_t it = 0;
                for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
                    TC_STAMP(it, 4);
                    const int4 dsc = __ldg(a.tdesc + tile);
                    const int t0 = dsc.z, len = dsc.y;
                    const int g0 = dsc.x + t0 + c.min_off;                              // tensor row of tile row 0
                    const int tlo = min(max(-(t0 + c.min_off), 0), c.rows_a), thi = min(max(len - (t0 + c.min_off), tlo), c.rows_a);
                    for (int ks = 0; ks < a.nks; ks++) {
                        const uint32_t lit = it * (uint32_t)a.nks + (uint32_t)ks;
                        const uint32_t abuf = lit % (uint32_t)c.nabuf, ause = lit / (uint32_t)c.nabuf;
                        tc::mbar_wait(bar_aempty0 + 8u * abuf, (ause & 1u) ^ 1u);     // MMAs that read this buffer have retired
                        TC_STAMP(it, 5);
                        const uint32_t dA = tc::smem_u32(sA + (size_t)abuf * c.a_bytes), bt = bar_tma0 + 8u * abuf;
                        const int col0 = a.xcol + ks * a.cin;
                        tc::mbar_expect_tx_e(bt, (uint32_t)(kc_total * c.rows_a) * 16u, el);
                        for (int kc = 0; kc < kc_total; kc++)
                            for (int bx = 0; bx < c.nboxes; bx++)
                                tc::tma_load_2d_e(dA + (uint32_t)(kc * c.rows_a + bx * c.box_rows) * 16u, &tmap, col0 + kc * 8, g0 + bx * c.box_rows, bt, el);
                        tc::mbar_wait(bt, ause & 1u);
                        const int nz = tlo + (c.rows_a - thi);
                        for (int i = lane; i < nz * kc_total; i += 32) {
                            const int kc = i / nz, j = i - kc * nz;
                            const int r = j < tlo ? j : thi + (j - tlo);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dA + (uint32_t)(kc * c.rows_a + r) * 16u), "r"(0u) : "memory");
                        }
                        tc::fence_proxy_async();
                        tc::mbar_arrive(bar_afull0 + 8u * abuf);
                        TC_STAMP(it, 6);
                    }
                }
            }
        } else {
        const int lt = tid - c.nepi * 32;
        const int nlt = c.nload * 32;                   // loader threads
        uint32_t it = 0;
        const int items = c.rows_need * kc_total;      // pad rows are never read by an MMA: not loaded (for a 1x1 conv the pad row alone
                                                       // cost every K slice a second memory round trip: 129 x 12 items for 8 x 192 slots)
        const int dr = nlt / kc_total, dk = nlt - dr * kc_total;   // advance of (r, kc) per nlt items
        int4 dnext = ((int)blockIdx.x < a.ntiles) ? __ldg(a.tdesc + blockIdx.x) : make_int4(0, 0, 0, 0);
        const int lr0 = lt / kc_total, lk0 = lt - lr0 * kc_total;          // this thread's first (row, chunk) item
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
            TC_STAMP(it, 4);
            const int4 dsc = dnext;
            if (tile + (int)gridDim.x < a.ntiles) dnext = __ldg(a.tdesc + tile + gridDim.x);
            const int t0 = dsc.z, len = dsc.y;
            const long row0 = (long)dsc.x;
            const int tlo = -(t0 + c.min_off), thi = len - (t0 + c.min_off);      // valid tile rows: tlo <= r < thi
            for (int ks = 0; ks < a.nks; ks++) {
            const uint32_t lit = it * (uint32_t)a.nks + (uint32_t)ks;         // activation tiles are counted per (tile, K slice)
            const uint32_t abuf = lit % (uint32_t)c.nabuf, ause = lit / (uint32_t)c.nabuf;
            tc::mbar_wait(bar_aempty0 + 8u * abuf, (ause & 1u) ^ 1u);     // MMAs that read this buffer have retired
            TC_STAMP(it, 5);
            uint8_t* dstA = sA + (size_t)abuf * c.a_bytes;
            const float* xbase = a.x + (row0 + t0 + c.min_off) * a.ldx + a.xcol + ks * a.cin;
            if (a.xb) {
                // the producer already wrote MMA-operand rows (bf16, activated): 16-byte cp.async straight into the K-major
                // tile, zero-filled outside the utterance; no registers, no conversion
                const __nv_bfloat16* xbb = a.xb + (row0 + t0 + c.min_off) * (long)a.ldxb + a.xcol + ks * a.cin;
                const uint32_t dA = tc::smem_u32(dstA);
                int r2 = lr0, kc2 = lk0;
                for (int i = lt; i < items; i += nlt) {
                    const bool okr = (r2 >= tlo) && (r2 < thi);
                    const __nv_bfloat16* src = xbb + (okr ? (long)r2 * a.ldxb : 0) + kc2 * 8;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dA + (uint32_t)(kc2 * c.rows_a + r2) * 16u), "l"(okr ? src : a.xb), "r"(okr ? 16u : 0u) : "memory");
                    r2 += dr; kc2 += dk;
                    if (kc2 >= kc_total) { kc2 -= kc_total; r2++; }
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                tc::fence_proxy_async();
                tc::mbar_arrive(bar_afull0 + 8u * abuf);
                TC_STAMP(it, 6);
                continue;
            }
            // 16-byte (8-channel) chunks, kc fastest so global reads are contiguous; TC_LOAD_BATCH items (2 x LDG.128 each) in
            // flight per thread before any conversion so the load latency is paid once per batch
            int r = lr0, kc = lk0;
            for (int base = lt; base < items; base += TC_LOAD_BATCH * nlt) {
                float4 v0[TC_LOAD_BATCH], v1[TC_LOAD_BATCH];
                const int r_s = r, kc_s = kc;
#pragma unroll
                for (int u = 0; u < TC_LOAD_BATCH; u++) {
                    v0[u] = make_float4(0.f, 0.f, 0.f, 0.f); v1[u] = v0[u];
                    if (base + u * nlt < items && r >= tlo && r < thi) {
                        const float4* src = reinterpret_cast<const float4*>(xbase + (long)r * a.ldx + kc * 8);
                        v0[u] = __ldg(src); v1[u] = __ldg(src + 1);
                    }
                    r += dr; kc += dk;
                    if (kc >= kc_total) { kc -= kc_total; r++; }
                }
                int r2 = r_s, kc2 = kc_s;              // the same walk again (cheaper than keeping every item's row / chunk in registers)
#pragma unroll
                for (int u = 0; u < TC_LOAD_BATCH; u++) {
                    const int ru = r2, ku = kc2;
                    r2 += dr; kc2 += dk;
                    if (kc2 >= kc_total) { kc2 -= kc_total; r2++; }
                    if (base + u * nlt >= items) continue;
                    float4 a0 = v0[u], a1 = v1[u];
                    if (a.in_act) {
                        a0.x = leaky(a0.x, a.in_slope); a0.y = leaky(a0.y, a.in_slope); a0.z = leaky(a0.z, a.in_slope); a0.w = leaky(a0.w, a.in_slope);
                        a1.x = leaky(a1.x, a.in_slope); a1.y = leaky(a1.y, a.in_slope); a1.z = leaky(a1.z, a.in_slope); a1.w = leaky(a1.w, a.in_slope);
                    }
                    uint4 pk;
                    pk.x = tc::pack_bf16(a0.x, a0.y); pk.y = tc::pack_bf16(a0.z, a0.w);
                    pk.z = tc::pack_bf16(a1.x, a1.y); pk.w = tc::pack_bf16(a1.z, a1.w);
                    *reinterpret_cast<uint4*>(dstA + ((size_t)ku * c.rows_a + ru) * 16) = pk;
                    if (a.split3) {
                        // lo plane: bf16(x - bf16(x)), exact subtraction in fp32
                        uint4 lo;
                        lo.x = tc::pack_bf16(a0.x - tc::bf16_lo_f(pk.x), a0.y - tc::bf16_hi_f(pk.x));
                        lo.y = tc::pack_bf16(a0.z - tc::bf16_lo_f(pk.y), a0.w - tc::bf16_hi_f(pk.y));
                        lo.z = tc::pack_bf16(a1.x - tc::bf16_lo_f(pk.z), a1.y - tc::bf16_hi_f(pk.z));
                        lo.w = tc::pack_bf16(a1.z - tc::bf16_lo_f(pk.w), a1.w - tc::bf16_hi_f(pk.w));
                        *reinterpret_cast<uint4*>(dstA + ((size_t)(ku + kc_total) * c.rows_a + ru) * 16) = lo;
                    }
                }
            }
            tc::fence_proxy_async();       // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc::mbar_arrive(bar_afull0 + 8u * abuf);
            TC_STAMP(it, 6);
            }
        }
        }
    } else if (warp == TC_EPI_WARPS + TC_LOAD_WARPS) {
        // ================= weight producer (cp.async.bulk ring; warp-uniform loop, copies predicated on the elected lane) =================
        {
            const uint32_t el = tc::elect_flag();
            const bool dbg_on = dbg_on_cta && el != 0;
            uint32_t s = 0, ph = 1;                       // ring slot and the parity to wait for on its "empty" barrier
            const uint32_t kc_bytes = (uint32_t)c.ntile * 16u;
            const uint32_t sW_u = tc::smem_u32(sW);
            const bool tiled = a.wtile != 0 && a.wtile == c.ntile;      // N-tiled weight copy: this N tile's chunks are contiguous
            const long kc_stride = tiled ? (long)c.ntile * 8 : (long)a.npad16 * 8;          // elements between consecutive 8-channel chunks
            const long tile_block = (long)a.ntaps * (a.split3 ? 3 : 1) * kc_total * c.ntile * 8;   // elements of one N tile's block (tiled copy)
            uint32_t piece = 0;
            for (int itp = 0; itp < ring_iters; itp++) {
                if (c.resident && itp > 0) break;
                TC_STAMP(itp, 7);
                for (int ks = 0; ks < a.nks; ks++) {
                const __nv_bfloat16* src = tiled ? a.wtc_t_ks[ks] + (long)ny * tile_block
                                                 : a.wtc_ks[ks] + (long)ny * c.ntile * 8;     // (tap 0, chunk 0) of this N tile
                const int nt_ks = a.ntaps_ks[0] ? a.ntaps_ks[ks] : a.ntaps;
                for (int tapseg = 0; tapseg < nt_ks * nseg; tapseg++) {
                    for (int ch0 = 0; ch0 < a.cin; ch0 += c.piece_ch) {
                        const int nkc = min(c.piece_ch, a.cin - ch0) >> 3;
                        if (!c.resident) tc::mbar_wait(bar_empty0 + 8u * s, ph);
                        const uint32_t fb = bar_full0 + 8u * s;
                        tc::mbar_expect_tx_e(fb, kc_bytes * (uint32_t)nkc, el);
                        uint32_t dst = sW_u + s * (uint32_t)c.slot_bytes;
                        // one N tile == all output columns: the chunks of a piece are contiguous in global memory -> ONE bulk copy per piece.
                        // (r01g experiment: with one 2-4 KB copy per 8-channel chunk the single producer thread's issue rate -- ~90 cycles per
                        // copy, 240 copies per tile of the summed second convs -- bounded every ring-mode launch, not L2 and not the ring depth)
                        const bool contig = tiled || (c.ntile == a.npad16);
                        if (c.cluster == 2) {
                            // pieces alternate between the two CTAs; the one whose turn it is fetches for both
                            if ((int)(piece & 1u) == cl_rank) {
                                if (contig) tc::bulk_g2s_mc_e(dst, src, kc_bytes * (uint32_t)nkc, fb, (uint16_t)3, el);
                                else for (int kc = 0; kc < nkc; kc++) tc::bulk_g2s_mc_e(dst + (uint32_t)kc * kc_bytes, src + (long)kc * kc_stride, kc_bytes, fb, (uint16_t)3, el);
                            }
                            src += (long)nkc * kc_stride;
                        } else if (contig) {
                            tc::bulk_g2s_e(dst, src, kc_bytes * (uint32_t)nkc, fb, el);
                            src += (long)nkc * kc_stride;
                        } else
                        for (int kc = 0; kc < nkc; kc++, dst += kc_bytes, src += kc_stride) tc::bulk_g2s_e(dst, src, kc_bytes, fb, el);
                        piece++;
                        if (++s == (uint32_t)c.nstages) { s = 0; ph ^= 1u; }
                    }
                    if (a.split3 && (tapseg & 1)) src += (long)kc_total * kc_stride;      // skip the repeated wh segment of this tap
                }
                }
                TC_STAMP(itp, 8);
            }
        }
    } else {
        // ================= MMA issuer (warp-uniform loop, tcgen05.mma / commit predicated on the elected lane) =================
        {
            const uint32_t el = tc::elect_flag();
            const uint32_t tmem_base = tc::uniform_u32(*tmem_slot);
            const bool dbg_on = dbg_on_cta && el != 0;
            const uint32_t idesc = tc::make_idesc(TC_M, c.ntile);
            const uint32_t sA_u = tc::smem_u32(sA), sW_u = tc::smem_u32(sW);
            const uint64_t dhi_a = tc::make_desc(0, lbo_a, 128u), dhi_b = tc::make_desc(0, lbo_b, 128u);   // start-address field = 0
            const uint32_t a_k16 = (2u * lbo_a) >> 4, b_k16 = (2u * lbo_b) >> 4;               // start-field advance per K = 16 step
            const uint32_t ahi = (uint32_t)(dhi_a >> 32), bhi = (uint32_t)(dhi_b >> 32);
            const uint32_t a_lbo_hi = (uint32_t)dhi_a & 0xFFFF0000u, b_lbo_hi = (uint32_t)dhi_b & 0xFFFF0000u;   // LBO fields of the low words
            const uint32_t piece_a16 = ((uint32_t)(c.piece_ch >> 3) * lbo_a) >> 4;     // A advance per 64-channel piece (16 B units)
            uint32_t s = 0, ph = 0, it = 0;                   // ring slot / parity of its "full" barrier
            uint32_t abuf = 0, aph = 0, cbuf = 0, cph = 0;
            for (int tile = blockIdx.x; (int)it < ring_iters; tile += gridDim.x, it++) {
                // cluster mode: `real` is false for the pair's last iteration when only the peer has a tile -- the ring handshakes still run
                const bool real = tile < a.ntiles;
                TC_STAMP((int)it, 9);
                if (real) tc::mbar_wait(bar_accempty0 + 8u * cbuf, cph ^ 1u);
                TC_STAMP((int)it, 11);
                const uint32_t dcol = tmem_base + cbuf * (uint32_t)c.ntile;
                uint32_t accum = 0;
                for (int ks = 0; ks < a.nks; ks++) {
                if (real) tc::mbar_wait(bar_afull0 + 8u * abuf, aph);
                TC_STAMP((int)it, 10);
                tc::tc_fence_after();
                const uint32_t a16 = ((sA_u + abuf * (uint32_t)c.a_bytes) >> 4) - (uint32_t)c.min_off;   // row 0 <-> tap offset 0
                const int nt_ks = a.ntaps_ks[0] ? a.ntaps_ks[ks] : a.ntaps;
                const int tap0 = a.ntaps_ks[0] ? a.tap0_ks[ks] : 0;
                const uint32_t lo_plane16 = ((uint32_t)kc_total * lbo_a) >> 4;     // the lo plane sits kc_total chunks behind the hi plane
                for (int tapseg = 0; tapseg < nt_ks * nseg; tapseg++) {
                    const int tap = a.split3 ? (tapseg >> 1) : tapseg;
                    const int seg = tapseg - tap * nseg;
                    // split3: piece wh (seg 0) multiplies the hi AND the lo activation plane, piece wl (seg 1) the hi plane
                    const int npass = (a.split3 && seg == 0) ? 2 : 1;
                    uint32_t arow16 = a16 + (uint32_t)a.toff[tap0 + tap];
                    for (int ch0 = 0; ch0 < a.cin; ch0 += c.piece_ch, arow16 += piece_a16) {
                        const int nk16 = min(c.piece_ch, a.cin - ch0) >> 4;
                        if (!c.resident || it == 0) { tc::mbar_wait(bar_full0 + 8u * s, ph); tc::tc_fence_after(); }
                        for (int pass = 0; pass < (real ? npass : 0); pass++) {
                        // descriptor low words: start address (16-byte units) | LBO << 16; only the start field moves (by plain 32-bit
                        // adds), the high words (SBO, version) are loop-invariant.  With 64-bit descriptors rebuilt per MMA the single
                        // issuing thread needed ~145 cycles per MMA (r01h timeline: 24 MMAs of a k3 conv in 3.3k cycles) against 64
                        // executed at N = 128 -- the issue loop, not the tensor pipe, paced every launch with N <= 256
                        uint32_t alo = ((arow16 + (pass ? lo_plane16 : 0u)) & 0x3FFFu) | a_lbo_hi;
                        uint32_t blo = (((sW_u + s * (uint32_t)c.slot_bytes) >> 4) & 0x3FFFu) | b_lbo_hi;
#pragma unroll 4
                        for (int k = 0; k < nk16; k++) {
                            tc::umma_bf16_lh_e(dcol, alo, ahi, blo, bhi, idesc, accum, el);
                            accum = 1;
                            alo += a_k16; blo += b_k16;
                        }
                        }
                        if (!c.resident) {                                          // frees the weight slot when these MMAs retire
                            if (c.cluster == 2) tc::umma_commit_mc_e(bar_empty0 + 8u * s, (uint16_t)3, el);   // ... in both CTAs of the pair
                            else tc::umma_commit_e(bar_empty0 + 8u * s, el);
                        }
                        if (++s == (uint32_t)c.nstages) { s = 0; ph ^= 1u; }
                    }
                }
                if (!real) continue;
                tc::umma_commit_e(bar_aempty0 + 8u * abuf, el);                     // activation buffer may be refilled
                if (++abuf == (uint32_t)c.nabuf) { abuf = 0; aph ^= 1u; }
                }
                if (c.resident) s = 0;
                if (!real) continue;
                tc::umma_commit_e(bar_accfull0 + 8u * cbuf, el);                    // accumulator ready for the epilogue
                TC_STAMP((int)it, 12);
                if (++cbuf == (uint32_t)c.naccbuf) { cbuf = 0; cph ^= 1u; }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (c.cluster == 2) tc::cluster_sync_all();         // no CTA leaves while its peer may still multicast into it / signal its barriers
    if (warp == TC_EPI_WARPS + TC_LOAD_WARPS) tc::tmem_dealloc(tmem_base, (uint32_t)c.tmem_cols);
}

// ------------------------------------------------------------------------------------------ host side
static inline bool conv_tc_supported(const ConvArgs& a) {
    if (!a.wtc) return false;
    if (a.cin % 16 || a.n % 16 || a.cin < 16 || a.n < 16) return false;
    if (a.xb ? (a.ldxb % 8 || a.xcol % 8 || a.split3) : (a.ldx % 4 || a.xcol % 4)) return false;
    if (a.epi == EPI_SPLIT && (a.split % 16)) return false;
    if (a.out_act == ACT_TANH) return false;          // no tanh epilogue on this path (conv_post has its own kernels)
    return true;
}

// cuTensorMapEncodeTiled, fetched from the driver at run time (the library links cudart only)
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline tmap_encode_fn tmap_encoder() {
    static tmap_encode_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<tmap_encode_fn>(p);
    }();
    return fn;
}
// 2-D tensor map over bf16 operand rows [rows, ld] (row-major) with boxes of [box_rows rows x 8 channels]: one box == one 8-channel
// plane segment of the K-major no-swizzle operand tile.  Rows outside [0, rows) read as zeros.
static inline bool make_rows_tmap(CUtensorMap* tm, const void* base, long rows, int ld, int box_rows) {
    tmap_encode_fn enc = tmap_encoder();
    if (!enc || rows < 1 || box_rows < 1 || box_rows > 256 || (reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {8u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static int g_tc_tma = 1;             // engine option "conv_tma" (process-wide): bf16 operand rows reach the activation tile by TMA
static int g_tc_tma_max_cin = 128;   // engine option "conv_tma_max_cin"

// Epilogue warps when the loader is the cp.async one.  12 (two loader warps) was measured NEUTRAL to slightly worse against 8
// (profiles/r02m_ab_conv_nepi.log: medium decoder 181-185 vs 184-186 ms, flow 53.5 vs 56 ms, `high` 203.5 vs 206.3 ms): the
// epilogue is not short of warps -- the small-channel launches move ~1.3 GB per conv and sit at 50-67 % of HBM bandwidth.
static int g_tc_nepi_xb = 8;         // engine option "conv_nepi" (8 or 12), process-wide
static inline bool conv_tc_plan(const ConvArgs& a, TcCfg& c, int num_sms = 0) {
    int mn = a.toff[0], mx = a.toff[0];
    for (int i = 1; i < a.ntaps; i++) { mn = a.toff[i] < mn ? a.toff[i] : mn; mx = a.toff[i] > mx ? a.toff[i] : mx; }
    c.min_off = mn;
    int rows = TC_M + (mx - mn);
    c.rows_need = rows;
    // TMA for activation tiles of up to g_tc_tma_max_cin channels per K slice: a box is 16 bytes wide (one 8-channel plane of the
    // no-swizzle layout), so the TMA unit fetches a 32-byte sector per row and plane and uses half of it; measured (profiles/r02y):
    // the 32-128-channel decoder launches gain 1.5 % (their loader warps' LDGSTS stream no longer paces the epilogue's loads through
    // the LSU queue), the 192-channel coupling-flow convs lose 8 % (52 KB tiles: the doubled sector traffic shows).
    // Not in latency mode (a launch of a few tiles, e.g. TTSVoice's one utterance per call): every launch carries its own tensor map, and
    // the TMA unit's first fetch of that descriptor is a ~1 us round trip per launch that the cp.async loader does not pay (C1: 1.82 ->
    // 1.87 ms per call with TMA everywhere).
    c.tma = (a.xb && !a.split3 && a.xb_rows > 0 && g_tc_tma && a.cin <= g_tc_tma_max_cin && tmap_encoder() != nullptr &&
             (num_sms <= 0 || (long)a.ntiles * 2 > (long)num_sms)) ? 1 : 0;
    c.nboxes = 1; c.box_rows = 0;
    if (c.tma) {
        // a whole number of equal boxes of a multiple of 8 rows each (<= 256, the TMA box limit): planes stay 128-byte aligned
        const int r8 = (rows + 7) / 8 * 8;
        c.nboxes = (r8 + 255) / 256; c.box_rows = ((r8 + c.nboxes - 1) / c.nboxes + 7) / 8 * 8;
        rows = c.nboxes * c.box_rows;
    } else
        rows = ((rows + 7) / 8) * 8 + 1;       // odd row count: (kc + r) mod 8 spreads the loaders' 16 B chunks over all banks
    c.rows_a = rows;
    const int planes = a.split3 ? 2 : 1, nseg = a.split3 ? 2 : 1;     // weight segments streamed per tap (split3: wh, wl)
    c.a_bytes = (planes * (a.cin / 8) * rows * 16 + 127) / 128 * 128;
    c.piece_ch = a.cin < TC_PIECE_CH ? a.cin : TC_PIECE_CH;
    c.cpt = (a.cin + c.piece_ch - 1) / c.piece_ch;
    c.npieces = (a.ntaps_ks[0] ? a.ntaps : a.ntaps * (a.nks > 0 ? a.nks : 1)) * nseg * c.cpt;
    const int limit = 222 * 1024;
    // epilogue-bound launches (every conv of <= 256 channels, r01/r02 timelines) get 12 epilogue warps when the activations arrive as
    // bf16 operand rows: the cp.async loader needs two warps, not six
    c.nepi = (a.xb && !a.split3) ? g_tc_nepi_xb : TC_EPI_WARPS;
    c.nload = TC_EPI_WARPS + TC_LOAD_WARPS - c.nepi;
    const int epi_bytes = c.nepi * 32 * TC_EPI_PITCH * 4 + 256 * 4;      // transpose buffers + this N tile's bias
    int nt = a.npad16 <= 256 ? a.npad16 : 0;
    if (!nt) for (int cand = 256; cand >= 16; cand -= 16) if (a.npad16 % cand == 0) { nt = cand; break; }
    // latency mode (round 2): what TTSVoice sends is ONE utterance per call (voice.py:350-351) -- a single 128-row tile, so a
    // launch was 1-3 CTAs streaming the whole weight matrix through one SM each (C1: 1.27 ms of a 2.03 ms call on the text side,
    // profiles/r02e_bench_C1.json).  With fewer CTAs than half the SMs the output columns are spread over more, narrower N tiles:
    // same K order per output (bit-identical results), 1/8 of the weights per CTA.
    if (num_sms > 0 && a.ntiles > 0)
        while (nt > 32 && (long)a.ntiles * (a.npad16 / nt) * 2 <= (long)num_sms) {
            int next = 0;
            for (int cand = nt - 16; cand >= 32; cand -= 16) if (a.npad16 % cand == 0) { next = cand; break; }
            if (!next) break;
            nt = next;
        }
    for (;;) {
        c.ntile = nt;
        c.slot_bytes = c.piece_ch * nt * 2;
        const long res_bytes = (long)c.npieces * c.slot_bytes;
        auto total = [&](int nabuf, int nstages) {
            return (nabuf * c.a_bytes + nstages * c.slot_bytes + (2 * nstages + TC_NFIXBAR) * 8 + 16 + 127) / 128 * 128 + epi_bytes;
        };
        bool ok = false;
        if (c.npieces <= TC_MAX_STAGES && total(2, c.npieces) <= limit) {     // (res_bytes up to ~120 KB: weights read once per CTA, not per tile)
            c.resident = 1; c.nstages = c.npieces; c.nabuf = 2; ok = true;
        } else {
            // widest N tile first (the activation tile is re-read once per N tile), then double-buffered
            // activations, then ring depth
            c.resident = 0;
            const int smax = c.npieces < 8 ? c.npieces : 8;          // ring depth: bytes in flight hide the L2 round trip
            for (int nabuf = 2; nabuf >= 1 && !ok; nabuf--)
                for (int ns = smax; ns >= (smax < 2 ? smax : 2) && !ok; ns--)
                    if (total(nabuf, ns) <= limit) { c.nabuf = nabuf; c.nstages = ns; ok = true; }
        }
        if (ok) {
            c.epi_off = (c.nabuf * c.a_bytes + c.nstages * c.slot_bytes + (2 * c.nstages + TC_NFIXBAR) * 8 + 16 + 127) / 128 * 128;
            c.smem_bytes = c.epi_off + epi_bytes;
            break;
        }
        // shrink the N tile (keeps divisibility of npad16)
        int next = 0;
        for (int cand = nt - 16; cand >= 16; cand -= 16) if (a.npad16 % cand == 0) { next = cand; break; }
        if (!next) return false;
        nt = next;
    }
    {   // column-group width of the epilogue: the one that leaves the busiest warp of a quadrant the fewest columns (32 on ties)
        const int nhh = c.nepi / 4;
        auto busiest = [&](int gw) { const int groups = (c.ntile + gw - 1) / gw; return ((groups + nhh - 1) / nhh) * gw; };
        c.gw = busiest(16) < busiest(32) ? 16 : 32;
    }
    c.naccbuf = (2 * c.ntile <= 512) ? 2 : 1;
    int tc_cols = 32; while (tc_cols < c.naccbuf * c.ntile) tc_cols <<= 1;
    c.tmem_cols = tc_cols;
    c.vec = (a.ldo % 4 == 0) && (a.ocol % 4 == 0) && (!a.res || (a.ldres % 4 == 0 && a.rescol % 4 == 0)) &&
            (!a.utab || a.utab_ld % 4 == 0);
    if (a.epi == EPI_GATE) c.vec = c.vec && (a.ocol % 4 == 0) && (a.ldo % 4 == 0);
    return true;
}

static inline cudaError_t conv_tc_launch(const ConvArgs& a, int num_sms, cudaStream_t st, bool cluster_ok = false) {
    TcCfg c;
    if (!conv_tc_plan(a, c, num_sms)) return cudaErrorInvalidConfiguration;
    // epilogue variant (compile-time in the kernel)
    typedef void (*KFn)(const ConvArgs, const TcCfg, const CUtensorMap);
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    if (c.tma && !make_rows_tmap(&tm, a.xb, a.xb_rows, a.ldxb, c.box_rows)) return cudaErrorInvalidValue;
    const bool anyacc = a.accumulate || (a.epi == EPI_SPLIT && a.accumulate2);
    KFn fn; int vid;
    if (!c.vec) { fn = k_conv_tc<-1, 0, 0>; vid = 0; }
    else if (a.epi == EPI_GATE) { fn = k_conv_tc<EPI_GATE, 0, 0>; vid = 1; }
    else if (a.epi == EPI_SUBFROM) { fn = k_conv_tc<EPI_SUBFROM, 1, 0>; vid = 2; }
    else if (a.epi == EPI_SPLIT) { if (a.res) { fn = k_conv_tc<EPI_SPLIT, 1, 1>; vid = 3; } else { fn = k_conv_tc<EPI_SPLIT, 0, 1>; vid = 4; } }
    else if (a.resb && a.nresb > 1) { if (anyacc || a.nresb > 3) return cudaErrorInvalidConfiguration; fn = k_conv_tc<EPI_STORE, 3, 0>; vid = 11; }
    else if (a.resb) { if (anyacc) { fn = k_conv_tc<EPI_STORE, 2, 1>; vid = 5; } else { fn = k_conv_tc<EPI_STORE, 2, 0>; vid = 6; } }
    else if (a.res) { if (anyacc) { fn = k_conv_tc<EPI_STORE, 1, 1>; vid = 7; } else { fn = k_conv_tc<EPI_STORE, 1, 0>; vid = 8; } }
    else { if (anyacc) { fn = k_conv_tc<EPI_STORE, 0, 1>; vid = 9; } else { fn = k_conv_tc<EPI_STORE, 0, 0>; vid = 10; } }
    static bool attr_set[64][12] = {{false}};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev][vid]) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        attr_set[dev][vid] = true;
    }
    // resident CTAs per SM: shared memory (228 KB/SM, 1 KB reserved per CTA), registers (64K / (regs * threads)),
    // TMEM columns (512 / tmem_cols)
    const int regs_per_thread = 128;            // __launch_bounds__(TC_THREADS, 1): 16 warps x 128 registers is the whole file
    int occ = (228 * 1024) / (c.smem_bytes + 1024);
    const int reg_occ = 65536 / (((regs_per_thread + 7) / 8 * 8) * TC_THREADS);
    if (occ > reg_occ) occ = reg_occ;
    if (occ > 8) occ = 8;
    if (occ < 1) occ = 1;
    const int tmem_occ = 512 / c.tmem_cols;
    if (occ > tmem_occ) occ = tmem_occ;
    const int ny = a.npad16 / c.ntile;
    int gx = num_sms * occ / (ny < 1 ? 1 : 1);
    if (gx > a.ntiles) gx = a.ntiles;
    if (gx < 1) gx = 1;
    // ring-mode launches: CTA pairs share the weight stream (multicast); needs an even grid
    c.cluster = (!c.resident && cluster_ok && gx >= 2) ? 2 : 1;
    if (c.cluster == 2) gx &= ~1;
    dim3 grid(gx, ny);
    if (c.cluster == 2) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = (size_t)c.smem_bytes; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, fn, a, c, tm);
    }
    launch_k(fn, grid, TC_THREADS, c.smem_bytes, st, a, c, tm);
    return cudaGetLastError();
}
