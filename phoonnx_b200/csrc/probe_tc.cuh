// Test-only micro-probe: issue rate of tcgen05.mma (kind::f16, M=128, K=16, operands in shared memory in the
// K-major no-swizzle layout conv_tc.cuh / mrf2_tc.cuh use) as a function of N.  One thread issues `iters` MMAs,
// rotating over `nd` accumulators and `na` activation row offsets, commits, waits; cycles = clock64 delta.
#pragma once
#include "conv_tc.cuh"

__global__ void __launch_bounds__(128, 1) k_mma_probe(int N, int iters, int nd, int na, int rows, unsigned long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x;
    const int a_bytes = 8 * rows * 16;                 // 8 chunks of 8 channels
    uint4* p = reinterpret_cast<uint4*>(smem);
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
        const uint32_t sA = tc::smem_u32(smem), sB = sA + a_bytes;
        const uint64_t dhi_a = tc::make_desc(0, (uint32_t)rows * 16u, 128u), dhi_b = tc::make_desc(0, (uint32_t)N * 16u, 128u);
        const uint64_t a_step = (uint64_t)(2 * rows), b_step = (uint64_t)(2 * N);
        const long long t0 = clock64();
        int d = 0, ar = 0;
        for (int i = 0; i < iters; i += 4) {
            const uint32_t dcol = tmem + (uint32_t)(d * N);
            uint64_t ad = dhi_a | (uint64_t)(((sA >> 4) + (uint32_t)(ar * 3)) & 0x3FFF);
            uint64_t bd = dhi_b | (uint64_t)((sB >> 4) & 0x3FFF);
#pragma unroll
            for (int k = 0; k < 4; k++) {
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
