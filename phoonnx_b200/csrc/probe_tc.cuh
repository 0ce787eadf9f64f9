// Test-only micro-probe: rate of tcgen05.mma (kind::f16, M=128, K=16) as a function of N and of WHERE / HOW the operands sit:
//   mode 0  A and B in shared memory, K-major no-swizzle core matrices (the layout conv_tc.cuh / mrf3_tc.cuh use)
//   mode 1  A and B in shared memory, K-major SWIZZLE_128B (rows of 64 channels = 128 B; what a tensor-map TMA load writes)
//   mode 2  A in TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc), B in shared memory no-swizzle
//   mode 3  A in tensor memory, B SWIZZLE_128B
//   mode 4  like 2, but every MMA is preceded by a tcgen05.cp.128x256b that copies ITS A operand (same smem descriptor the smem-A
//           MMA would use, so the tap row shift still works) from shared memory into a rotating tensor-memory staging slot
//   mode 5  the tcgen05.cp of mode 4 alone (no MMA): what the copy costs by itself
//   mode 6  cta_group::2 (k_mma_probe2, a cluster of two CTAs): M = 256 over the pair, each CTA feeds its own 128 rows of A and HALF of
//           the B operand (N/2 rows) from its shared memory, no swizzle; cycles per MMA as seen by the issuing CTA
//   mode 7  mode 6 with A in tensor memory
// One thread issues `iters` MMAs, rotating over `nd` accumulators and `na` activation row offsets, commits, waits; cycles = clock64
// delta.  Modes 1-3 answer round 1's open question: is the 32 + N/4 cycle cost of small-N MMAs a property of
// the no-swizzle layout, or of reading a 4 KB A operand through the 128 B/clk shared-memory port at all?
#pragma once
#include "conv_tc.cuh"

template <int mode>
__global__ void __launch_bounds__(128, 1) k_mma_probe(int N, int iters, int nd, int na, int rows, unsigned long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x;
    const int a_bytes = (8 * rows * 16 + 1023) & ~1023;    // 8 chunks of 8 channels (== rows x 128 B swizzled); B starts on a 1024 B atom
    uint8_t* base = smem + ((1024u - (tc::smem_u32(smem) & 1023u)) & 1023u);
    uint4* p = reinterpret_cast<uint4*>(base);
    for (int i = tid; i < (a_bytes + 8 * 256 * 16) / 16; i += 128) p[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0u, 0x3c003c00u);
    if (tid == 0) { tc::mbar_init(tc::smem_u32(&bar), 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    if (tid < 32) tc::tmem_alloc(tc::smem_u32(&tslot), 512u);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tslot;
    if (tid < 32) {
      if (tc::elect_one()) {
        const uint32_t idesc = tc::make_idesc(128, N);
        const uint32_t sA = tc::smem_u32(base), sB = sA + a_bytes;
        constexpr bool a_tmem = mode >= 2, sw = (mode & 1) != 0 && mode < 4;
        // SWIZZLE_128B K-major: 8-row x 128-byte atoms (SBO = 1024 B), layout type 2 at bits 61-63, LBO unused; a K=16 step is +32 B
        const uint64_t sw_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        const uint64_t dhi_a = sw ? sw_hi : tc::make_desc(0, (uint32_t)rows * 16u, 128u), dhi_b = sw ? sw_hi : tc::make_desc(0, (uint32_t)N * 16u, 128u);
        const uint64_t a_step = sw ? 2u : (uint64_t)(2 * rows), b_step = sw ? 2u : (uint64_t)(2 * N);
        const uint32_t a_row_units = sw ? 8u : 1u;            // 16-byte units per activation row
        const long long t0 = clock64();
        int d = 0, ar = 0;
        for (int i = 0; i < iters; i += 4) {
            const uint32_t dcol = tmem + (uint32_t)(d * N);
            // activation row offset per rotation: 3 rows (a tap shift) unswizzled; a whole 8-row atom swizzled (no base-offset field needed)
            uint64_t ad = dhi_a | (uint64_t)(((sA >> 4) + (uint32_t)(ar * (sw ? 8 : 3)) * a_row_units) & 0x3FFF);
            uint64_t bd = dhi_b | (uint64_t)((sB >> 4) & 0x3FFF);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (mode >= 4) {
                    const uint32_t at = tmem + 384u + (uint32_t)(((i + k) & 15) * 8);
                    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(at), "l"(ad) : "memory");
                    if (mode == 4)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                     ::"r"(dcol), "r"(at), "l"(bd), "r"(idesc), "r"((i >= 4 * nd || k) ? 1u : 0u) : "memory");
                } else if (a_tmem) {
                    // A operand from tensor memory: 128 lanes x 8 columns (16 bf16) per K step, behind the accumulators
                    const uint32_t at = tmem + 384u + (uint32_t)(((ar & 3) * 4 + k) * 8);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(dcol), "r"(at), "l"(bd), "r"(idesc), "r"((i >= 4 * nd || k) ? 1u : 0u) : "memory");
                } else
                tc::umma_bf16(dcol, ad, bd, idesc, (i >= 4 * nd || k) ? 1u : 0u);
                ad += a_step; bd += b_step;
            }
            if (++d == nd) d = 0;
            if (++ar == na) ar = 0;
        }
        const long long t1 = clock64();
        tc::umma_commit(tc::smem_u32(&bar));
        tc::mbar_wait(tc::smem_u32(&bar), 0);
        const long long t2 = clock64();
        out[2 * blockIdx.x] = (unsigned long long)(t1 - t0);
        out[2 * blockIdx.x + 1] = (unsigned long long)(t2 - t0);
      }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (tid < 32) tc::tmem_dealloc(tmem, 512u);
}

// cta_group::2 variant: launched as clusters of two CTAs; rank 0 of each pair issues for both.  Every wait is bounded.
template <int A_TMEM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_mma_probe2(int N, int iters, int nd, int na, int rows, unsigned long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x;
    const int a_bytes = (8 * rows * 16 + 1023) & ~1023;
    uint8_t* base = smem + ((1024u - (tc::smem_u32(smem) & 1023u)) & 1023u);
    uint4* p = reinterpret_cast<uint4*>(base);
    for (int i = tid; i < (a_bytes + 8 * 256 * 16) / 16; i += 128) p[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0u, 0x3c003c00u);
    if (tid == 0) { tc::mbar_init(tc::smem_u32(&bar), 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    if (tid < 32) {
        // both CTAs of the pair allocate (same warp, same shared-memory slot offset), cute::TMEM::Allocator2Sm
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync_all();
    tc::tc_fence_after();
    const uint32_t tmem = tslot;
    const uint32_t rank = tc::cluster_ctarank();
    if (rank == 0 && tid < 32) {
      if (tc::elect_one()) {
        const uint32_t idesc = tc::make_idesc(256, N);
        const uint32_t sA = tc::smem_u32(base), sB = sA + a_bytes;
        const int nb = N / 2;                                   // B rows held by each CTA
        const uint64_t dhi_a = tc::make_desc(0, (uint32_t)rows * 16u, 128u), dhi_b = tc::make_desc(0, (uint32_t)nb * 16u, 128u);
        const uint64_t a_step = (uint64_t)(2 * rows), b_step = (uint64_t)(2 * nb);
        const long long t0 = clock64();
        int d = 0, ar = 0;
        for (int i = 0; i < iters; i += 4) {
            const uint32_t dcol = tmem + (uint32_t)(d * N);
            uint64_t ad = dhi_a | (uint64_t)(((sA >> 4) + (uint32_t)(ar * 3)) & 0x3FFF);
            uint64_t bd = dhi_b | (uint64_t)((sB >> 4) & 0x3FFF);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t acc = (i >= 4 * nd || k) ? 1u : 0u;
                if (A_TMEM) {
                    const uint32_t at = tmem + 384u + (uint32_t)(((ar & 3) * 4 + k) * 8);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(dcol), "r"(at), "l"(bd), "r"(idesc), "r"(acc) : "memory");
                } else {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(dcol), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
                }
                ad += a_step; bd += b_step;
            }
            if (++d == nd) d = 0;
            if (++ar == na) ar = 0;
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(&bar)) : "memory");
        uint32_t spins = 0;
        bool ok = true;
        while (!tc::mbar_test(tc::smem_u32(&bar), 0)) { if (++spins > (1u << 26)) { ok = false; break; } }
        const long long t2 = clock64();
        out[2 * (blockIdx.x >> 1)] = ok ? (unsigned long long)(t1 - t0) : 0ull;
        out[2 * (blockIdx.x >> 1) + 1] = ok ? (unsigned long long)(t2 - t0) : 0ull;
      }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync_all();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

