// Monotonic alignment search on the device (include/mas_b200.h).
//
// Reference: maximum_path_each / maximum_path_c, /root/reference/phoonnx_train/vits/monotonic_align/core.pyx:7-42 -- a dynamic
// programme over value[t_y][t_x] (row y depends on row y-1 only) followed by a greedy backtrack from (t_y-1, t_x-1), run on the CPU
// under OpenMP `prange` over the batch, between a D2H copy of neg_cent and an H2D copy of the path (monotonic_align/__init__.py:14-21).
//
// Here: one CTA per batch item, one thread per text position x (NX positions per thread above 1024 columns).  The running row lives
// in shared memory (two buffers), so a row costs one barrier; the item's neg_cent rows are prefetched P rows ahead into registers
// (a row is only t_x * 4 bytes: without P rows in flight a CTA would be bound by one HBM round trip per row).  The forward pass does
// not store value[][] at all: the backtrack only ever asks "value[y-1][x] < value[y-1][x-1] ?", so each row leaves ONE BIT per cell
// (a warp ballot; shared memory if the item's bit matrix fits, else a global scratch), decided on exactly the float32 numbers the
// reference would have stored -- including cells outside the band, which keep their raw neg_cent value in the reference's in-place
// loop and therefore here.  Warp 0 then walks the bits backwards, 32 rows per round (the column moves by at most one per row, so the
// needed words of the next 32 rows are known up front and are fetched by the 32 lanes in parallel), and finally all threads write
// the item's whole [t_y_max][t_x_max] output slab -- zeros and ones -- in one coalesced pass (no separate memset).
//
// The kernel is memory-latency / barrier bound by construction (t_y dependent steps); its HBM traffic is the algorithmic minimum:
// values read once, paths written once.
#include "../../include/mas_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace {

thread_local char g_err[512] = "";
thread_local float g_ms = 0.f;

int fail(int code, const char* fmt, const char* what) {
    snprintf(g_err, sizeof g_err, fmt, what);
    return code;
}

constexpr float kNeg = -1e9f;                 // core.pyx:7 max_neg_val
constexpr int kMaxThreads = 1024;
constexpr int kMaxTx = 24576;                 // two float rows of kMaxTx + 1 columns in shared memory
constexpr int kSmemBudget = 200 * 1024;

struct MasSmem {
    int row_floats;                           // one row buffer: columns + 1 (slot 0 = column -1)
    int bits_words, idx_ints;                 // 0 = that table lives in the global scratch
    int bytes;
};

// NX: columns per thread, P: rows prefetched (register ring; NX * P = 16 keeps 1024-thread CTAs inside 64 registers); OutT: int32_t or float
template <int NX, int P, typename OutT>
__global__ void __launch_bounds__(kMaxThreads, 1)
k_mas(OutT* __restrict__ paths, const float* __restrict__ values, const int* __restrict__ t_ys, const int* __restrict__ t_xs,
      int t_y_max, int t_x_max, int wpr, uint32_t* __restrict__ gbits, int* __restrict__ gidx, MasSmem sm) {
    extern __shared__ __align__(16) uint8_t smem[];
    float* rowbuf = reinterpret_cast<float*>(smem);
    uint32_t* sbits = reinterpret_cast<uint32_t*>(rowbuf + 2 * sm.row_floats);
    int* sidx = reinterpret_cast<int*>(sbits + sm.bits_words);
    const int item = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    // the reference indexes memoryviews of the full arrays: extents beyond them are a caller error, clamped here instead of trusted
    const int t_y = min(max(t_ys[item], 0), t_y_max), t_x = min(max(t_xs[item], 0), t_x_max);
    const float* val = values + (size_t)item * t_y_max * t_x_max;
    uint32_t* bits = sm.bits_words ? sbits : gbits + (size_t)item * t_y_max * wpr;
    int* idx = sm.idx_ints ? sidx : gidx + (size_t)item * t_y_max;

    for (int i = tid; i < 2 * sm.row_floats; i += nthr) rowbuf[i] = 0.f;
    __syncthreads();

    // ---- forward pass: value[y][x] += max(value[y-1][x-1], value[y-1][x]) inside the band (core.pyx:16-29)
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
            if (y < t_y) {                                     // uniform over the CTA
                const float* prev = rowbuf + buf * sm.row_floats;
                float* cur = rowbuf + (buf ^ 1) * sm.row_floats;
                const int lo = max(0, t_x + y - t_y), hi = min(t_x, y + 1);
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    const int x = j * nthr + tid;
                    bool bit = false;
                    if (x < t_x) {
                        const float pv = prev[x + 1], pl = prev[x];        // value[y-1][x], value[y-1][x-1]
                        bit = (y > 0) && (x > 0) && (pv < pl);             // the backtrack's question at row y (core.pyx:33)
                        float v = nc[p][j];
                        if (x >= lo && x < hi) {
                            const float v_cur = (x == y) ? kNeg : pv;
                            const float v_prev = (x == 0) ? (y == 0 ? 0.f : kNeg) : pl;
                            v += (v_cur > v_prev) ? v_cur : v_prev;        // Cython's max(v_prev, v_cur)
                        }
                        cur[x + 1] = v;
                    }
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
    if (!sm.bits_words) __threadfence_block();                   // bits written to global by this CTA, read back by warp 0 below
    __syncthreads();

    // ---- backtrack (core.pyx:31-34): warp 0, 32 rows per round
    if (tid < 32) {
        int index = t_x - 1;
        for (int yb = t_y - 1; yb >= 0 && t_x > 0; yb -= 32) {
            const int y = yb - lane;
            const int wa = max(index - 31, 0) >> 5, wb = index >> 5;   // the column can only move left, by at most one per row
            uint32_t a = 0, b = 0;
            if (y >= 0) { a = bits[(size_t)y * wpr + wa]; b = bits[(size_t)y * wpr + wb]; }
            for (int k = 0; k < 32 && yb - k >= 0; k++) {
                const uint32_t ak = __shfl_sync(0xffffffffu, a, k), bk = __shfl_sync(0xffffffffu, b, k);
                const int yk = yb - k;
                if (lane == k) idx[yk] = index;
                const uint32_t bit = (((index >> 5) == wb ? bk : ak) >> (index & 31)) & 1u;
                if (index != 0 && (index == yk || bit)) index--;
            }
        }
    }
    if (!sm.idx_ints) __threadfence_block();
    __syncthreads();

    // ---- the item's output slab, zeros included
    OutT* out = paths + (size_t)item * t_y_max * t_x_max;
    if ((t_x_max & 3) == 0 && (reinterpret_cast<uintptr_t>(paths) & 15) == 0) {
        // 16-byte stores: four columns per thread
        const int q = t_x_max >> 2;
        const size_t total = (size_t)t_y_max * q;
        for (size_t i = tid; i < total; i += nthr) {
            const int y = (int)(i / q), x4 = (int)(i - (size_t)y * q) << 2;
            const int ix = (y < t_y && t_x > 0) ? idx[y] : -1;
            OutT o[4];
#pragma unroll
            for (int e = 0; e < 4; e++) o[e] = (OutT)(x4 + e == ix ? 1 : 0);
            *reinterpret_cast<int4*>(out + (size_t)y * t_x_max + x4) = *reinterpret_cast<const int4*>(o);
        }
    } else {
        const size_t total = (size_t)t_y_max * t_x_max;
        for (size_t i = tid; i < total; i += nthr) {
            const int y = (int)(i / t_x_max), x = (int)(i - (size_t)y * t_x_max);
            const int ix = (y < t_y && t_x > 0) ? idx[y] : -1;
            out[i] = (OutT)(x == ix ? 1 : 0);
        }
    }
}

template <int NX, int P, typename OutT>
cudaError_t launch(void* paths, const float* values, const int* t_ys, const int* t_xs, int b, int t_y_max, int t_x_max, int wpr,
                   uint32_t* gbits, int* gidx, const MasSmem& sm, int threads, cudaStream_t st) {
    auto kern = k_mas<NX, P, OutT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return e;
    kern<<<b, threads, sm.bytes, st>>>(static_cast<OutT*>(paths), values, t_ys, t_xs, t_y_max, t_x_max, wpr, gbits, gidx, sm);
    return cudaGetLastError();
}

template <typename OutT>
cudaError_t dispatch(int nx, void* paths, const float* values, const int* t_ys, const int* t_xs, int b, int t_y_max, int t_x_max,
                     int wpr, uint32_t* gbits, int* gidx, const MasSmem& sm, int threads, cudaStream_t st) {
    switch (nx) {
        case 1: return launch<1, 16, OutT>(paths, values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st);
        case 2: return launch<2, 8, OutT>(paths, values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st);
        case 4: return launch<4, 4, OutT>(paths, values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st);
        case 8: return launch<8, 2, OutT>(paths, values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st);
        case 16: return launch<16, 1, OutT>(paths, values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st);
        default: return launch<24, 1, OutT>(paths, values, t_ys, t_xs, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st);
    }
}

}  // namespace

extern "C" {

const char* mas_last_error(void) { return g_err; }
float mas_last_ms(void) { return g_ms; }

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
    g_ms = 0.f;
    if (b < 0 || t_y_max < 0 || t_x_max < 0) return fail(MAS_E_INVALID, "mas_maximum_path: %s", "negative extent");
    if (b == 0 || t_y_max == 0 || t_x_max == 0) return MAS_OK;                  // nothing to write
    if (!paths || !values || !t_ys || !t_xs) return fail(MAS_E_INVALID, "mas_maximum_path: %s", "null pointer");
    if (t_x_max > kMaxTx) return fail(MAS_E_INVALID, "mas_maximum_path: %s", "t_x_max exceeds 24576 columns");
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
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // geometry: one thread per column up to 1024, then NX columns per thread
    int nx = 1;
    while (nx < 24 && (long)nx * kMaxThreads < t_x_max) nx = nx < 16 ? nx * 2 : 24;
    int threads = ((t_x_max + nx - 1) / nx + 31) / 32 * 32;
    if (threads < 64) threads = 64;
    if (threads > kMaxThreads) threads = kMaxThreads;
    const int wpr = (t_x_max + 31) / 32;
    MasSmem sm;
    sm.row_floats = ((nx * threads + 1) + 3) / 4 * 4;
    long used = 2L * sm.row_floats * 4;
    const long bits_bytes = (long)t_y_max * wpr * 4, idx_bytes = (long)t_y_max * 4;
    sm.idx_ints = (used + idx_bytes <= kSmemBudget) ? t_y_max : 0;
    used += (long)sm.idx_ints * 4;
    sm.bits_words = (used + bits_bytes <= kSmemBudget) ? t_y_max * wpr : 0;
    used += (long)sm.bits_words * 4;
    sm.bytes = (int)used;

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
        if (!sm.bits_words) MAS_CK(cudaMallocAsync(&gbits, (size_t)b * bits_bytes, st), "cudaMallocAsync(bits)");
        if (!sm.idx_ints) MAS_CK(cudaMallocAsync(&gidx, (size_t)b * idx_bytes, st), "cudaMallocAsync(idx)");
        if (timed) {
            MAS_CK(cudaEventCreate(&ev0), "cudaEventCreate");
            MAS_CK(cudaEventCreate(&ev1), "cudaEventCreate");
            MAS_CK(cudaEventRecord(ev0, st), "cudaEventRecord");
        }
        MAS_CK(f32 ? dispatch<float>(nx, p_out, v_in, ty_in, tx_in, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st)
                   : dispatch<int32_t>(nx, p_out, v_in, ty_in, tx_in, b, t_y_max, t_x_max, wpr, gbits, gidx, sm, threads, st),
               "k_mas launch");
        if (timed) MAS_CK(cudaEventRecord(ev1, st), "cudaEventRecord");
        if (!on_dev) MAS_CK(cudaMemcpyAsync(paths, d_path, cells * 4, cudaMemcpyDeviceToHost, st), "D2H paths");
        if (!on_dev || timed) MAS_CK(cudaStreamSynchronize(st), "cudaStreamSynchronize");
        if (timed) MAS_CK(cudaEventElapsedTime(&g_ms, ev0, ev1), "cudaEventElapsedTime");
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
    return rc;
}

}  // extern "C"
