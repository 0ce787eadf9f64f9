/* mas_b200.h -- C ABI of the monotonic alignment search kernel in libvits_b200.so (B200 / sm_100a only).
 *
 * SURVEY.md 8(f)-4, the training-side neighbour of the synthesis path: `monotonic_align.maximum_path`
 * (/root/reference/phoonnx_train/vits/monotonic_align/__init__.py:7-21), whose compiled core is the Cython function
 *     maximum_path_c(int[:,:,::1] paths, float[:,:,::1] values, int[::1] t_ys, int[::1] t_xs)      core.pyx:38-42
 * (one maximum_path_each per batch item, core.pyx:7-34), called once per training step from SynthesizerTrn.forward
 * (phoonnx_train/vits/models.py:646-650) after a device->host copy of neg_cent and followed by a host->device copy of the path.
 * mas_maximum_path is that call with the same arguments in the same order (plus the array extents a memoryview carries
 * implicitly), on tensors that stay on the device.
 *
 * Semantics kept bit for bit (tests/test_gpu_mas.py): float32 running sums in the reference's evaluation order, the band
 * max(0, t_x + y - t_y) <= x < min(t_x, y + 1), the -1e9 sentinel, the reference's `b > a ? b : a` maximum, and a backtrack that
 * compares the values AS STORED (cells outside the band keep their raw neg_cent value, as in the in-place Cython loop).
 * Differences, both deliberate: `values` is read-only here (the reference accumulates in place into its private numpy copy), and an
 * item with t_x > t_y (no monotonic path; the reference reads out of bounds there) yields the same greedy backtrack without the
 * out-of-bounds read.
 */
#ifndef MAS_B200_H
#define MAS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAS_OK 0
#define MAS_E_INVALID (-1)   /* null pointer, negative extent, t_x_max beyond the kernel's row buffer (24576) */
#define MAS_E_CUDA (-2)      /* a CUDA call failed or the device is not compute capability 10.x: mas_last_error() */

#define MAS_DEVICE_PTRS 0x1  /* paths / values / t_ys / t_xs are device pointers (else host: staged, and the call synchronises) */
#define MAS_PATH_F32 0x2     /* paths is float32 0.0 / 1.0 (what maximum_path() returns for fp32 models) instead of int32 */

/* paths  [b][t_y_max][t_x_max]  out: 1 on the alignment path of item i inside its t_ys[i] x t_xs[i] corner, 0 elsewhere
 * values [b][t_y_max][t_x_max]  in:  neg_cent (float32), not modified
 * t_ys, t_xs [b]                in:  valid rows (frames) / columns (text positions) of each item
 * stream: a cudaStream_t (null = the legacy default stream); with MAS_DEVICE_PTRS the call only enqueues work on it. */
int mas_maximum_path(void* paths, const float* values, const int32_t* t_ys, const int32_t* t_xs, int b, int t_y_max,
                     int t_x_max, int flags, void* stream);

/* last error message of the calling thread ("" if none) */
const char* mas_last_error(void);

/* device time of the last mas_maximum_path on this thread, in ms (0 if timing was not requested): set MAS_TIMED in flags and the call
 * brackets its kernels with CUDA events and synchronises the stream before returning */
#define MAS_TIMED 0x4
#define MAS_ROW_KERNEL 0x8   /* use the general one-barrier-per-row kernel (the only one above 1024 columns) also below: for tests / A-B */
float mas_last_ms(void);
float mas_last_forward_ms(void);   /* the forward + backtrack launch alone; the rest of mas_last_ms() is the output pass */

#ifdef __cplusplus
}
#endif
#endif
