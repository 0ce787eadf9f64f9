// Shared declarations for the B200 VITS engine kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <limits.h>

#define CONV_MAX_TAPS 16
#define CONV_MAX_SLICES 16

enum ConvEpi { EPI_STORE = 0, EPI_GATE = 1, EPI_SPLIT = 2, EPI_SUBFROM = 3 };
enum OutAct { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

// One 1-D convolution over a packed, channel-last varlen batch, expressed as an implicit
// GEMM:  out[t, n] = epi( sum_tap sum_ci act(x[t + toff[tap], ci]) * W[tap][ci][n] + bias[n] ).
// Rows of utterance b are [cu[b]*rate, cu[b+1]*rate); samples outside the utterance read 0
// (the zero padding of every Conv1d on the path; the decoder is unmasked, models.py:720,
// so B=1 semantics == zero padding at utterance edges).
// ---- programmatic dependent launch (PDL).  Every kernel of the engine starts with pdl_enter(): wait for the preceding grid of the
// stream to complete and flush (griddepcontrol.wait -- a no-op when the launch carries no programmatic dependency), then allow the
// NEXT grid to be launched (griddepcontrol.launch_dependents): its CTAs are scheduled while this grid runs and sit in their own
// pdl_enter() until this grid is done.  What overlaps is the launch latency only -- no kernel touches its inputs early -- which is what a
// one-utterance call consists of: ~150 dependent launches of a few CTAs each (bench C1).  Host side: launch_k() below.
// The two halves, for kernels with a prologue worth overlapping (conv_tc.cuh): trigger first, run everything that touches only
// constants (barriers, tensor memory, bias, the weight ring), and wait only in the warps that read or write activations.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

struct ConvArgs {
    const float* x;  int ldx;  int xcol;  int cin;
    int ntaps;  int toff[CONV_MAX_TAPS];
    const float* w;            // fp32 [ntaps][cin][npad]
    const __nv_bfloat16* wtc;  // bf16 tcgen05 layout [ntaps][cin/8][npad16][8] (may be null)
    int n;  int npad;          // npad: n rounded up to a multiple of 4 (fp32) ; npad16 to 16 (tc)
    int npad16;
    const float* bias;         // [npad] or null
    const float* utab;  const int* uidx;  int utab_ld;   // per-utterance bias row utab[uidx[b]*utab_ld + n] or null
    int in_act;  float in_slope;      // 1: leaky-relu(slope) applied to x on load
    int epi;
    const float* res;  int ldres;  int rescol;           // residual (added; SUBFROM: res - v)
    int out_act;  float out_div;  int accumulate;
    float* out;  int ldo;  int ocol;
    float* out2;  int ldo2;  int ocol2;  int split;  int accumulate2;   // EPI_SPLIT: n >= split -> out2[n - split]
    const int* cu;  const int* tile_cu;  int B;  int rate;  int ntiles;
    const int4* tdesc;          // tcgen05 path: per 128-row tile {utterance's first row, its rows, tile's first row in it, utterance} (k_tile_desc)
    // split3 (tcgen05 path only): fp32-faithful bf16x3 product  x*w ~= xh*wh + xh*wl + xl*wh  (xh = bf16(x), xl =
    // bf16(x - xh)).  The activation tile holds a hi and a lo plane; wtc then holds 3*cin input channels per tap
    // ([wh | wl | wh]).  Used for the text side, whose output feeds ceil(exp(logw)).
    int split3;
    // K slices (tcgen05 path): the kernel loops over nks slices of `cin` input channels each (activation columns xcol + ks * cin,
    // weights wtc_ks[ks]) and accumulates them in TMEM: one epilogue per tile, activation tiles double-buffered across slices
    int nks;  const __nv_bfloat16* wtc_ks[CONV_MAX_SLICES];
    // N-tiled copies of the same weights (n > 256 only): [n / wtile][tap][cin/8][wtile][8], so that an N tile's pieces are contiguous
    // and cross the weight ring as ONE bulk copy each; used when the launch plan's N tile equals wtile (0: none)
    int wtile;  const __nv_bfloat16* wtc_t_ks[CONV_MAX_SLICES];
    // per-slice tap lists (tcgen05 path, optional): slice ks uses taps toff[tap0_ks[ks] .. + ntaps_ks[ks]) (`ntaps` = the flattened
    // total); all zero = every slice uses toff[0 .. ntaps).  Lets ONE launch sum convolutions with different kernels / dilations.
    int ntaps_ks[CONV_MAX_SLICES];  int tap0_ks[CONV_MAX_SLICES];
    const __nv_bfloat16* xb;  int ldxb;       // tcgen05 path: the input as bf16 MMA-operand rows [rows, ldxb] (already activated) instead of fp32 `x`;
                                            // TMA'd (or cp.async'd) straight into the operand tile (xcol applies)
    long xb_rows;                           // host side: rows of the xb array (> 0 enables the TMA loader: its tensor map needs the extent)
    int nresb;  int resb_stride;            // > 1: the residual is the SUM of nresb such row groups, resb_stride columns apart (<= 3)
    const __nv_bfloat16* resb;  int ldresb;  float resb_slope;   // residual given as bf16 lrelu_{slope} rows: res = min(v, v / slope) (rescol applies)
    __nv_bfloat16* outb;  float outb_slope;   // EPI_STORE / EPI_GATE on the tcgen05 path: write bf16(lrelu_{slope}(v)) here instead of fp32 `out`
    unsigned long long* dbg;   // test-only phase timeline of CTA (0,0): [tile][16] clock64 stamps, or null
};

__device__ __forceinline__ int find_segment(const int* __restrict__ offs, int B, int v) {
    // largest b in [0, B) with offs[b] <= v   (offs is non-decreasing, offs[0] == 0)
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(offs + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ float leaky(float v, float slope) { return v >= 0.f ? v : v * slope; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- host: every launch of the engine goes through here (PDL attribute; engine option "pdl" = 0 launches plainly)
static int g_pdl = 1;
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
