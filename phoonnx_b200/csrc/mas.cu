// Monotonic alignment search on the device (include/mas_b200.h).
//
// Reference: maximum_path_each / maximum_path_c, /root/reference/phoonnx_train/vits/monotonic_align/core.pyx:7-42 -- a dynamic
// programme over value[t_y][t_x] (row y depends on row y-1 only) followed by a greedy backtrack from (t_y-1, t_x-1), run on the CPU
// under OpenMP `prange` over the batch, between a D2H copy of neg_cent and an H2D copy of the path (monotonic_align/__init__.py:14-21).
//
// Here, two launches:
//
//  1. k_mas_wave (t_x_max <= 1024; k_mas_rows above): one CTA per batch item, one thread per text position x.  The recurrence only
//     looks left and up, so a warp needs nothing from the other warps except ONE number per row -- value[y-1][32w - 1], the last
//     column of the warp to its left.  The warps therefore run as a WAVEFRONT, each a block of 8 rows behind its left neighbour,
//     and hand that column over through a small shared-memory ring with a progress counter per warp (no CTA-wide barrier in the
//     loop; the running row lives in registers, the left neighbour comes from a shuffle).  A row's dependent chain is
//     shuffle -> select -> max -> add (~100 cycles per row measured) instead of load / store / barrier through shared memory
//     (~290 cycles per row measured with the barrier-per-row kernel below, which remains the general path).  neg_cent is staged by per-thread 4-byte cp.async into a shared-memory FIFO
//     32-64 rows deep: a row of one item is only t_x * 4 bytes, so the depth of the prefetch, not the width of the load, is what
//     buys bandwidth.
//     The forward pass does not store value[][] at all: the backtrack only ever asks "value[y-1][x] < value[y-1][x-1] ?", so each
//     row leaves ONE BIT per cell (a warp ballot per row; shared memory if the item's bit matrix fits, else a global scratch),
//     decided on exactly the float32 numbers the reference would have stored -- including cells outside the band, which keep their
//     raw neg_cent value in the reference's in-place loop and therefore here.  Warp 0 then walks the bits backwards, 32 rows per
//     round (the column moves left by at most one per row, so the 32 columns a round can visit are known up front: each lane
//     fetches its row's window of them in parallel, the serial part is a shift and a compare per row), leaving the path's column per row.
//  2. k_mas_write: the whole [b][t_y_max][t_x_max] output -- zeros and ones -- in one coalesced, grid-wide pass of 16-byte stores
//     (no separate memset; one CTA per item could not saturate HBM on the write side).
//
// HBM traffic is the algorithmic minimum: values read once, paths written once (+ 4 bytes per row of scratch).
#include "../../include/mas_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <type_traits>

namespace {

thread_local char g_err[512] = "";
thread_local float g_ms = 0.f, g_ms_fwd = 0.f;

int fail(int code, const char* fmt, const char* what) {
    snprintf(g_err, sizeof g_err, fmt, what);
    return code;
}

constexpr float kNeg = -1e9f;                 // core.pyx:7 max_neg_val
constexpr int kMaxThreads = 1024;
constexpr int kMaxTx = 24576;                 // k_mas_rows: two float rows of kMaxTx + 1 columns in shared memory
constexpr int kSmemBudget = 200 * 1024;
constexpr int kR = 8;                         // wavefront: rows per hand-over block
constexpr int kD = 64;                        // wavefront: depth of the boundary ring (rows)

// one cell of the forward pass (core.pyx:17-29).  pv = value[y-1][x], pl = value[y-1][x-1] as stored; returns the new value[y][x]
// and the backtrack's decision for row y at column x (core.pyx:33)
__device__ __forceinline__ float mas_cell(float nc, float pv, float pl, int x, int y, int lo, int hi, bool& bit) {
    bit = (y > 0) && (x > 0) && (pv < pl);
    if (x >= lo && x < hi) {
        const float v_cur = (x == y) ? kNeg : pv;
        const float v_prev = (x == 0) ? (y == 0 ? 0.f : kNeg) : pl;
        nc += (v_cur > v_prev) ? v_cur : v_prev;                  // Cython's max(v_prev, v_cur): `v_cur > v_prev ? v_cur : v_prev`
    }
    return nc;                                                    // outside the band the reference leaves the raw neg_cent in place
}

// Backtrack of one item by one warp (core.pyx:31-34): idx[y] = column of the path in row y.  32 rows per round: the column moves left
// by at most one per row, so a round can only visit columns [index - 31, index]; every lane extracts that 32-bit window of ITS row's
// decision bits up front, and the serial walk keeps the column RELATIVE to the window (rel = column - base), so one step is
// shift -> test -> predicated decrement.  t_x, t_y and hence index are warp-uniform.
__device__ __forceinline__ void mas_backtrack(const uint32_t* bits, int wpr, int t_y, int t_x, int* idx, int lane) {
    if (t_x <= 0) return;
    __syncwarp();
    int index = t_x - 1;
    for (int yb = t_y - 1; yb >= 0; yb -= 32) {
        const int y = yb - lane;
        const int base = index - 31;
        const int wa = max(base, 0) >> 5, wb = index >> 5;
        uint32_t a = 0, b = 0;
        if (y >= 0) { a = bits[(size_t)y * wpr + wa]; b = bits[(size_t)y * wpr + wb]; }
        // V: bit r = decision bit of column base + r in this lane's row
        const uint32_t V = base < 0 ? (a << (-base)) : (wa == wb ? a : __funnelshift_r(a, b, base & 31));
        // fold the two other reasons into the same 32-bit mask, per lane and before the serial walk: `index == y` (core.pyx:33) sets
        // the bit of column y, `index != 0` clears the bit of column 0; then one step is  rel -= (M_k >> rel) & 1
        const uint32_t diag = (uint32_t)(y - base) < 32u ? 1u << (y - base) : 0u;
        const uint32_t zero = (uint32_t)(-base) < 32u ? 1u << (-base) : 0u;
        const uint32_t M = y >= 0 ? ((V | diag) & ~zero) : 0u;    // rows above the item: no move (unused)
        int rel = 31, mine = 0;
#pragma unroll
        for (int k = 0; k < 32; k++) {
            const uint32_t Mk = __shfl_sync(0xffffffffu, M, k);
            mine = lane == k ? rel : mine;
            rel -= (int)((Mk >> rel) & 1u);
        }
        if (y >= 0) idx[y] = base + mine;
        index = base + rel;
    }
}

__device__ __forceinline__ int ld_volatile_s32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
// Predicated forms (one instruction, no branch): lane-dependent `if`s inside the wavefront loop made ptxas treat the warp as diverged
// and route every shuffle through its WARPSYNC.COLLECTIVE slow path (measured: 2-5x slower rows).
__device__ __forceinline__ void cp_async4_if(bool p, void* smem_dst, const void* gsrc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 4;\n\t}"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"((int)p) : "memory");
}
__device__ __forceinline__ void st_u32_if(bool p, uint32_t* dst, uint32_t v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.b32 [%0], %1;\n\t}" ::"l"(dst), "r"(v), "r"((int)p) : "memory");
}
__device__ __forceinline__ void st_ring_if(bool p, unsigned long long* slot, float v, int tag) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.volatile.shared.v2.b32 [%0], {%1, %2};\n\t}"
                 ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "r"(__float_as_uint(v)), "r"(tag), "r"((int)p) : "memory");
}

struct WaveSmem {
    int bits_words;                           // 0 = the bit matrix lives in the global scratch
    int bytes;
};

// ------------------------------------------------------------------------------------------ wavefront kernel (t_x_max <= 1024)
// F: rows of the cp.async FIFO (a multiple of kR)
template <int F>
__global__ void __launch_bounds__(kMaxThreads, 1)
k_mas_wave(const float* __restrict__ values, const int* __restrict__ t_ys, const int* __restrict__ t_xs, int t_y_max, int t_x_max,
           int wpr, uint32_t* __restrict__ gbits, int* __restrict__ gidx, WaveSmem sm) {
    constexpr int G = F / kR;                 // cp.async groups in flight
    extern __shared__ __align__(16) uint8_t smem[];
    const int item = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, nwarps = nthr >> 5;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform in a form ptxas can see: the shuffles below stay plain SHFLs
    float* fifo = reinterpret_cast<float*>(smem);                              // [F][nthr]
    unsigned long long* ring = reinterpret_cast<unsigned long long*>(fifo + (size_t)F * nthr);    // [nwarps][kD] {row tag, value}: last column of warp w
    int* prog = reinterpret_cast<int*>(ring + nwarps * kD);                    // [nwarps]: rows completed
    uint32_t* sbits = reinterpret_cast<uint32_t*>(prog + ((nwarps + 3) & ~3));
    // the reference indexes memoryviews of the full arrays: extents beyond them are a caller error, clamped here instead of trusted
    const int t_y = __shfl_sync(0xffffffffu, min(max(t_ys[item], 0), t_y_max), 0), t_x = __shfl_sync(0xffffffffu, min(max(t_xs[item], 0), t_x_max), 0);
    const float* val = values + (size_t)item * t_y_max * t_x_max;
    uint32_t* bits = sm.bits_words ? sbits : gbits + (size_t)item * t_y_max * wpr;
    if (tid < nwarps) prog[tid] = 0;
    for (int i = tid; i < nwarps * kD; i += nthr) ring[i] = ~0ull;            // tag -1: no row
    __syncthreads();

    const int x = tid;
    const bool warp_on = 32 * w < t_x;                             // this warp owns at least one column of the item
    const bool has_right = 32 * (w + 1) < t_x;                     // somebody consumes this warp's last column
    if (warp_on) {
        // Lanes past t_x run along on garbage: it can only flow to the right, i.e. into other such lanes; their decision bits are
        // masked out of the ballots and their loads are suppressed.
        const uint32_t colmask = (t_x - 32 * w >= 32) ? 0xffffffffu : ((1u << (t_x - 32 * w)) - 1u);
        const int ty_eff = x < t_x ? t_y : 0;                      // `row exists for this thread`  <=>  y < ty_eff
        // band of core.pyx:17: max(0, t_x + y - t_y) <= x < min(t_x, y + 1)  <=>  0 <= y - x <= t_y - t_x   (empty if t_x > t_y)
        const int dband = t_y - t_x;
        const int ylo = dband >= 0 ? x : 0x3fffffff;
        const uint32_t dbandu = (uint32_t)max(dband, 0);
        const size_t pitch = (size_t)t_x_max;
        const int rowstride = nthr;                                // floats between two rows of the FIFO
        float* fslot0 = fifo + tid;
        // prologue: rows 0 .. F-1
        {
            const float* gp = val + x;
#pragma unroll 1
            for (int g = 0; g < G; g++) {
#pragma unroll
                for (int r = 0; r < kR; r++) {
                    const int y = g * kR + r;
                    cp_async4_if(y < ty_eff, fslot0 + (size_t)y * rowstride, gp);
                    gp += pitch;
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        }
        const float* gnext = val + (size_t)F * pitch + x;          // next row to prefetch: y0 + F + r
        float cur = 0.f;                                           // value[y-1][x]
        const unsigned long long* ring_left = ring + (size_t)(w > 0 ? w - 1 : 0) * kD;
        unsigned long long* ring_mine = ring + (size_t)w * kD;
        uint32_t* brow = bits + w;                                 // decision word of (row y, this warp): brow[y * wpr]

        // one block of kR rows; TAIL: the last, partial block (rows past t_y are computed and dropped -- no shuffle sits under a
        // condition, so they all stay plain SHFL / VOTE instructions)
        auto block = [&](const int y0, auto tail_tag) {
            constexpr bool TAIL = decltype(tail_tag)::value;
            asm volatile("cp.async.wait_group %0;" ::"n"(G - 1) : "memory");       // this thread's rows of the block have landed
            float* fs = fslot0 + (size_t)(y0 % F) * rowstride;
            float nc[kR];
#pragma unroll
            for (int r = 0; r < kR; r++) nc[r] = fs[(size_t)r * rowstride];
            // lane 0's left neighbour per row: value[y-1][32w - 1] from the ring, or for column 0 the reference's v_prev
            // (0 in row 0, else max_neg_val: core.pyx:22-26)
            // left neighbour of lane 0 per row: value[y-1][32w - 1] from the ring, or for column 0 the reference's v_prev (0 in row 0,
            // else max_neg_val: core.pyx:22-26).  A ring slot is one 64-bit word {value, row tag}: the tag arriving IS the hand-over,
            // so no fence is needed (a fence here would also wait for this thread's 32-64 outstanding cp.async rows: measured 3.2k
            // cycles per block).  Lane r polls the slot of row y0 + r - 1; the loop is warp-uniform (vote).
            float bl = (w == 0 && y0 + lane > 0) ? kNeg : 0.f;
            if (w > 0) {
                const int yl = y0 + (lane & (kR - 1)) - 1;
                const bool want = yl >= 0 && yl + 1 < t_y;
                const volatile unsigned long long* slot = ring_left + ((yl + kD) % kD);
                unsigned long long sv;
                do { sv = *slot; } while (!__all_sync(0xffffffffu, !want || (int)(sv >> 32) == yl));
                bl = want ? __uint_as_float((uint32_t)sv) : 0.f;
            }
            if (has_right && y0 + kR - kD + 1 > 0) {
                // the ring slots about to be overwritten hold rows y0 - kD ...: the right neighbour must be past them
                while (ld_volatile_s32(prog + w + 1) < y0 + kR - kD + 1) { }
            }
            unsigned long long* rslot = ring_mine + (y0 % kD);
            uint32_t myword = 0;
#pragma unroll
            for (int r = 0; r < kR; r++) {
                const int y = y0 + r;
                float left = __shfl_up_sync(0xffffffffu, cur, 1);
                const float bnd = __shfl_sync(0xffffffffu, bl, r);
                left = lane == 0 ? bnd : left;
                // the backtrack's question at row y (core.pyx:33), asked of the values as stored; in row 0 both sides are the initial 0
                const uint32_t word = __ballot_sync(0xffffffffu, cur < left) & colmask;
                const float v_cur = (y == ylo) ? kNeg : cur;                              // core.pyx:18-21 (x == y)
                const float m = (v_cur > left) ? v_cur : left;                            // Cython's max(v_prev, v_cur)
                const float sum = nc[r] + m;                                              // core.pyx:29
                const float v = ((uint32_t)(y - ylo) <= dbandu) ? sum : nc[r];            // outside the band the raw neg_cent stays in place
                const bool row_on = !TAIL || y < t_y;
                cur = row_on ? v : cur;
                myword = lane == r ? word : myword;                                       // lane r keeps row r's decision word
                st_ring_if(row_on && lane == 31 && has_right, rslot + r, v, y);
            }
            st_u32_if(lane < kR && y0 + lane < t_y, brow + (size_t)(y0 + lane) * wpr, myword);
            // progress: the left neighbour may reuse the ring slots this warp has read
            st_u32_if(w > 0 && lane == 0, reinterpret_cast<uint32_t*>(prog + w), (uint32_t)(y0 + kR));
            // refill the FIFO rows just consumed (their loads have been used)
#pragma unroll
            for (int r = 0; r < kR; r++) {
                cp_async4_if(y0 + F + r < ty_eff, fs + (size_t)r * rowstride, gnext);
                gnext += pitch;
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        int y0 = 0;
        for (; y0 + kR <= t_y; y0 += kR) block(y0, std::false_type{});
        if (y0 < t_y) block(y0, std::true_type{});
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    if (!sm.bits_words) __threadfence_block();                     // bits written to global by this CTA are read back by warp 0
    __syncthreads();
    if (w == 0) mas_backtrack(bits, wpr, t_y, t_x, gidx + (size_t)item * t_y_max, lane);
}

// ------------------------------------------------------------------------------------------ row-barrier kernel (any t_x_max <= kMaxTx)
struct RowsSmem {
    int row_floats;                           // one row buffer: columns + 1 (slot 0 = column -1)
    int bits_words;                           // 0 = the bit matrix lives in the global scratch
    int bytes;
};

// NX: columns per thread, P: rows prefetched (register ring; NX * P = 16 keeps 1024-thread CTAs inside 64 registers).  The running
// row lives in shared memory (two buffers), one CTA barrier per row.
template <int NX, int P>
__global__ void __launch_bounds__(kMaxThreads, 1)
k_mas_rows(const float* __restrict__ values, const int* __restrict__ t_ys, const int* __restrict__ t_xs, int t_y_max, int t_x_max,
           int wpr, uint32_t* __restrict__ gbits, int* __restrict__ gidx, RowsSmem sm) {
    extern __shared__ __align__(16) uint8_t smem[];
    float* rowbuf = reinterpret_cast<float*>(smem);
    uint32_t* sbits = reinterpret_cast<uint32_t*>(rowbuf + 2 * sm.row_floats);
    const int item = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    const int t_y = __shfl_sync(0xffffffffu, min(max(t_ys[item], 0), t_y_max), 0), t_x = __shfl_sync(0xffffffffu, min(max(t_xs[item], 0), t_x_max), 0);
    const float* val = values + (size_t)item * t_y_max * t_x_max;
    uint32_t* bits = sm.bits_words ? sbits : gbits + (size_t)item * t_y_max * wpr;

    for (int i = tid; i < 2 * sm.row_floats; i += nthr) rowbuf[i] = 0.f;
    __syncthreads();

    float nc[P][NX];
#pragma unroll
    for (int p = 0; p < P; p++)
#pragma unroll
        for (int j = 0; j < NX; j++) {
            const int x = j * nthr + tid;
            nc[p][j] = (p < t_y && x < t_x) ? __ldg(val + (size_t)p * t_x_max + x) : 0.f;
        }
    int buf = 0;
    for (int y0 = 0; y0 < t_y; y0 += P) {
        float nx[P][NX];
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
            for (int j = 0; j < NX; j++) {
                const int x = j * nthr + tid, y = y0 + P + p;
                nx[p][j] = (y < t_y && x < t_x) ? __ldg(val + (size_t)y * t_x_max + x) : 0.f;
            }
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int y = y0 + p;
            if (y < t_y) {                                         // uniform over the CTA
                const float* prev = rowbuf + buf * sm.row_floats;
                float* cur = rowbuf + (buf ^ 1) * sm.row_floats;
                const int lo = max(0, t_x + y - t_y), hi = min(t_x, y + 1);
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    const int x = j * nthr + tid;
                    bool bit = false;
                    if (x < t_x) cur[x + 1] = mas_cell(nc[p][j], prev[x + 1], prev[x], x, y, lo, hi, bit);
                    const uint32_t word = __ballot_sync(0xffffffffu, bit);
                    if (lane == 0 && (x >> 5) < wpr) bits[(size_t)y * wpr + (x >> 5)] = word;
                }
                __syncthreads();
                buf ^= 1;
            }
        }
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
            for (int j = 0; j < NX; j++) nc[p][j] = nx[p][j];
    }
    if (!sm.bits_words) __threadfence_block();
    __syncthreads();
    if (__shfl_sync(0xffffffffu, tid >> 5, 0) == 0) mas_backtrack(bits, wpr, t_y, t_x, gidx + (size_t)item * t_y_max, lane);
}

// ------------------------------------------------------------------------------------------ output: zeros and ones, grid-wide
// threads (32, 8): x strides over a row's 16-byte chunks (or columns), y over rows; grid (row groups, items) -- no divisions
template <typename OutT>
__global__ void __launch_bounds__(256) k_mas_write(OutT* __restrict__ paths, const int* __restrict__ t_ys, const int* __restrict__ t_xs,
                                                   const int* __restrict__ gidx, int t_y_max, int t_x_max, int vec) {
    const int item = blockIdx.y;
    const int t_y = min(max(t_ys[item], 0), t_y_max), t_x = min(max(t_xs[item], 0), t_x_max);
    const int* idx = gidx + (size_t)item * t_y_max;
    OutT* out = paths + (size_t)item * t_y_max * t_x_max;
    for (int y = blockIdx.x * blockDim.y + threadIdx.y; y < t_y_max; y += gridDim.x * blockDim.y) {
        const int ix = (y < t_y && t_x > 0) ? __ldg(idx + y) : -1;
        OutT* row = out + (size_t)y * t_x_max;
        if (vec) {
            for (int x4 = threadIdx.x * 4; x4 < t_x_max; x4 += 128) {      // 16-byte stores: four columns per thread
                OutT o[4];
#pragma unroll
                for (int e = 0; e < 4; e++) o[e] = (OutT)(x4 + e == ix ? 1 : 0);
                *reinterpret_cast<int4*>(row + x4) = *reinterpret_cast<const int4*>(o);
            }
        } else {
            for (int x = threadIdx.x; x < t_x_max; x += 32) row[x] = (OutT)(x == ix ? 1 : 0);
        }
    }
}

template <int F>
cudaError_t launch_wave(const float* values, const int* t_ys, const int* t_xs, int b, int t_y_max, int t_x_max, int wpr,
                        uint32_t* gbits, int* gidx, const WaveSmem& sm, int threads, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_mas_wave<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return e;
    k_mas_wave<F><<<b, threads, sm.bytes, st>>>(values, t_ys, t_xs, t_y_max, t_x_max, wpr, gbits, gidx, sm);
    return cudaGetLastError();
}

template <int NX, int P>
cudaError_t launch_rows(const float* values, const int* t_ys, const int* t_xs, int b, int t_y_max, int t_x_max, int wpr,
                        uint32_t* gbits, int* gidx, const RowsSmem& sm, int threads, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_mas_rows<NX, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return e;
    k_mas_rows<NX, P><<<b, threads, sm.bytes, st>>>(values, t_ys, t_xs, t_y_max, t_x_max, wpr, gbits, gidx, sm);
    return cudaGetLastError();
}

// forward pass + backtrack of the whole batch: leaves the path's column per row in gidx [b][t_y_max]
cudaError_t mas_forward(const float* values, const int* t_ys, const int* t_xs, int b, int t_y_max, int t_x_max, int* gidx,
                        uint32_t** gbits_out, int force_rows, cudaStream_t st) {
    const int wpr = (t_x_max + 31) / 32;
    const long bits_bytes = (long)t_y_max * wpr * 4;
    *gbits_out = nullptr;
    cudaError_t e;
    if (t_x_max <= kMaxThreads && !force_rows) {
        int threads = (t_x_max + 31) / 32 * 32;
        if (threads < 32) threads = 32;
        const int nwarps = threads / 32;
        const int F = threads <= 512 ? 64 : 32;
        long used = (long)F * threads * 4 + (long)nwarps * kD * 8 + (long)((nwarps + 3) & ~3) * 4;
        WaveSmem sm;
        sm.bits_words = (used + bits_bytes <= kSmemBudget) ? t_y_max * wpr : 0;
        sm.bytes = (int)(used + (long)sm.bits_words * 4);
        if (!sm.bits_words && (e = cudaMallocAsync(gbits_out, (size_t)b * bits_bytes, st)) != cudaSuccess) return e;
        return F == 64 ? launch_wave<64>(values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, *gbits_out, gidx, sm, threads, st)
                       : launch_wave<32>(values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, *gbits_out, gidx, sm, threads, st);
    }
    // geometry: NX columns per thread
    int nx = 1;
    while (nx < 24 && (long)nx * kMaxThreads < t_x_max) nx = nx < 16 ? nx * 2 : 24;
    int threads = ((t_x_max + nx - 1) / nx + 31) / 32 * 32;
    if (threads < 64) threads = 64;
    if (threads > kMaxThreads) threads = kMaxThreads;
    RowsSmem sm;
    sm.row_floats = ((nx * threads + 1) + 3) / 4 * 4;
    const long used = 2L * sm.row_floats * 4;
    sm.bits_words = (used + bits_bytes <= kSmemBudget) ? t_y_max * wpr : 0;
    sm.bytes = (int)(used + (long)sm.bits_words * 4);
    if (!sm.bits_words && (e = cudaMallocAsync(gbits_out, (size_t)b * bits_bytes, st)) != cudaSuccess) return e;
    uint32_t* gbits = *gbits_out;
    switch (nx) {
        case 1: return launch_rows<1, 16>(values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st);
        case 2: return launch_rows<2, 8>(values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st);
        case 4: return launch_rows<4, 4>(values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st);
        case 8: return launch_rows<8, 2>(values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st);
        case 16: return launch_rows<16, 1>(values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st);
        default: return launch_rows<24, 1>(values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st);
    }
}

}  // namespace

extern "C" {

const char* mas_last_error(void) { return g_err; }
float mas_last_ms(void) { return g_ms; }
float mas_last_forward_ms(void) { return g_ms_fwd; }

#define MAS_CK(call, what)                                                      \
    do {                                                                        \
        cudaError_t e_ = (call);                                                \
        if (e_ != cudaSuccess) {                                                \
            snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e_)); \
            rc = MAS_E_CUDA;                                                    \
            goto done;                                                          \
        }                                                                       \
    } while (0)

int mas_maximum_path(void* paths, const float* values, const int32_t* t_ys, const int32_t* t_xs, int b, int t_y_max, int t_x_max,
                     int flags, void* stream) {
    g_err[0] = 0;
    g_ms = g_ms_fwd = 0.f;
    if (b < 0 || t_y_max < 0 || t_x_max < 0) return fail(MAS_E_INVALID, "mas_maximum_path: %s", "negative extent");
    if (b == 0 || t_y_max == 0 || t_x_max == 0) return MAS_OK;                  // nothing to write
    if (!paths || !values || !t_ys || !t_xs) return fail(MAS_E_INVALID, "mas_maximum_path: %s", "null pointer");
    if (t_x_max > kMaxTx) return fail(MAS_E_INVALID, "mas_maximum_path: %s", "t_x_max exceeds 24576 columns");
    if (b > 65535) return fail(MAS_E_INVALID, "mas_maximum_path: %s", "more than 65535 items in one call");
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
        return fail(MAS_E_CUDA, "mas_maximum_path: %s", "no CUDA device (this library has no CPU fallback)");
    if (prop.major != 10) return fail(MAS_E_CUDA, "mas_maximum_path: %s", "device is not compute capability 10.x (built for sm_100a only)");

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool on_dev = (flags & MAS_DEVICE_PTRS) != 0, f32 = (flags & MAS_PATH_F32) != 0, timed = (flags & MAS_TIMED) != 0;
    const size_t cells = (size_t)b * t_y_max * t_x_max;
    int rc = MAS_OK;
    float* d_val = nullptr; void* d_path = nullptr; int* d_ty = nullptr; int* d_tx = nullptr;
    uint32_t* gbits = nullptr; int* gidx = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm = nullptr;
    {
        const float* v_in = values; const int* ty_in = t_ys; const int* tx_in = t_xs; void* p_out = paths;
        if (!on_dev) {
            MAS_CK(cudaMallocAsync(&d_val, cells * 4, st), "cudaMallocAsync(values)");
            MAS_CK(cudaMallocAsync(&d_path, cells * 4, st), "cudaMallocAsync(paths)");
            MAS_CK(cudaMallocAsync(&d_ty, (size_t)b * 4, st), "cudaMallocAsync(t_ys)");
            MAS_CK(cudaMallocAsync(&d_tx, (size_t)b * 4, st), "cudaMallocAsync(t_xs)");
            MAS_CK(cudaMemcpyAsync(d_val, values, cells * 4, cudaMemcpyHostToDevice, st), "H2D values");
            MAS_CK(cudaMemcpyAsync(d_ty, t_ys, (size_t)b * 4, cudaMemcpyHostToDevice, st), "H2D t_ys");
            MAS_CK(cudaMemcpyAsync(d_tx, t_xs, (size_t)b * 4, cudaMemcpyHostToDevice, st), "H2D t_xs");
            v_in = d_val; ty_in = d_ty; tx_in = d_tx; p_out = d_path;
        }
        MAS_CK(cudaMallocAsync(&gidx, (size_t)b * t_y_max * 4, st), "cudaMallocAsync(idx)");
        if (timed) {
            MAS_CK(cudaEventCreate(&ev0), "cudaEventCreate");
            MAS_CK(cudaEventCreate(&ev1), "cudaEventCreate");
            MAS_CK(cudaEventCreate(&evm), "cudaEventCreate");
            MAS_CK(cudaEventRecord(ev0, st), "cudaEventRecord");
        }
        MAS_CK(mas_forward(v_in, ty_in, tx_in, b, t_y_max, t_x_max, gidx, &gbits, (flags & MAS_ROW_KERNEL) != 0, st), "forward kernel launch");
        if (timed) MAS_CK(cudaEventRecord(evm, st), "cudaEventRecord");
        {
            const int vec = (t_x_max & 3) == 0 && (reinterpret_cast<uintptr_t>(p_out) & 15) == 0;
            long gx = (t_y_max + 8 * 4 - 1) / (8 * 4);                              // ~4 rows per thread row
            const long cap = (148L * 8 + b - 1) / b;                               // enough CTAs to fill the machine, no more
            if (gx > cap) gx = cap;
            if (gx < 1) gx = 1;
            const dim3 grid((unsigned)gx, (unsigned)b), blk(32, 8);
            if (f32) k_mas_write<float><<<grid, blk, 0, st>>>(static_cast<float*>(p_out), ty_in, tx_in, gidx, t_y_max, t_x_max, vec);
            else k_mas_write<int32_t><<<grid, blk, 0, st>>>(static_cast<int32_t*>(p_out), ty_in, tx_in, gidx, t_y_max, t_x_max, vec);
            MAS_CK(cudaGetLastError(), "k_mas_write launch");
        }
        if (timed) MAS_CK(cudaEventRecord(ev1, st), "cudaEventRecord");
        if (!on_dev) MAS_CK(cudaMemcpyAsync(paths, d_path, cells * 4, cudaMemcpyDeviceToHost, st), "D2H paths");
        if (!on_dev || timed) MAS_CK(cudaStreamSynchronize(st), "cudaStreamSynchronize");
        if (timed) MAS_CK(cudaEventElapsedTime(&g_ms, ev0, ev1), "cudaEventElapsedTime");
        if (timed) MAS_CK(cudaEventElapsedTime(&g_ms_fwd, ev0, evm), "cudaEventElapsedTime");
    }
done:
    if (gbits) cudaFreeAsync(gbits, st);
    if (gidx) cudaFreeAsync(gidx, st);
    if (d_val) cudaFreeAsync(d_val, st);
    if (d_path) cudaFreeAsync(d_path, st);
    if (d_ty) cudaFreeAsync(d_ty, st);
    if (d_tx) cudaFreeAsync(d_tx, st);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (evm) cudaEventDestroy(evm);
    return rc;
}

}  // extern "C"
