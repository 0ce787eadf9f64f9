/*
 * vits_b200_test.h -- test-only entry points of libvits_b200.so.  NOT part of the drop-in boundary (include/vits_b200.h):
 * nothing on the reference side binds these; tests/ and tools/ use them to pin single kernels against numpy.
 */
#ifndef VITS_B200_TEST_H
#define VITS_B200_TEST_H

#include "vits_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Test hook: one convolution (single utterance of L rows, channel-last) through the production
 * launch path -- use_tc = 0: fp32 CUDA-core kernel, 1: tcgen05 kernel.  `out` is [L, out_cols]
 * and is read first when `accumulate` is set.  epi: 0 store, 1 gate (out_cols = n/2),
 * 2 split (first n/2 columns accumulate into out[:, :n/2], the rest into out[:, n/2:]), 3 res - v. */
int vits_test_conv(vits_handle* h, int use_tc, const float* x, int L, int cin, const int* taps, int ntaps,
                   const float* w32, const uint16_t* wtc, const float* bias, int n, int in_act, float in_slope,
                   int epi, const float* res, int accumulate, float out_div, int out_act, float* out, int out_cols);

/* Test hook: tcgen05.mma issue-rate probe (M=128, K=16, N=n; `nd` accumulators and `na` activation row offsets in
 * rotation, operand tile of `rows` rows).  Returns the average cycles per MMA (issue only / issue + completion). */
int vits_test_mma_probe(vits_handle* h, int n, int iters, int nd, int na, int rows, int nctas, double* issue_cycles,
                        double* total_cycles);
/* mode: 0 both operands in shared memory, K-major no-swizzle (== vits_test_mma_probe); 1 both SWIZZLE_128B; 2 A operand in tensor
 * memory, B no-swizzle; 3 A in tensor memory, B SWIZZLE_128B; 4 = 2 with a tcgen05.cp smem->tmem of the A operand before each MMA; 5 that
 * copy alone; 6 cta_group::2 (pairs of CTAs, nctas even, n a multiple of 32), operands in shared memory; 7 = 6 with A in tensor memory
 * (csrc/probe_tc.cuh). */
int vits_test_mma_probe_mode(vits_handle* h, int mode, int n, int iters, int nd, int na, int rows, int nctas, double* issue_cycles,
                             double* total_cycles);

/* Test hooks of the native voice loader (no GPU needed): the architecture vits_open would infer from `path`, and one packed blob
 * by name (bytes copied to `out` when it fits; returns its size in bytes, or a negative VITS_E_*; *dtype as in vits_upload).
 * name == NULL: returns the number of blobs; name "#<i>": copies the i-th blob's NAME into `out` instead. */
int vits_test_file_arch(const char* path, vits_arch* arch, char* err, size_t err_cap);
int64_t vits_test_file_blob(const char* path, const char* name, void* out, int64_t cap, int* dtype);

/* Host-side launch plans of the two tensor-core kernels, without a GPU (the planners are plain host code): the CPU tests check their
 * invariants (shared-memory budget, TMA box geometry, tile steps shared by the kernels of one stage).
 * vits_test_mrf3_plan: one fused stage / ResBlock1 pair of `C` channels with `nrb` resblocks of kernel sizes k[], first / second
 * dilations d1[], d2[]; up_cin > 0: with the fused ConvTranspose (u = 4) from up_cin channels.  out[16] receives
 * {nb, span, hmax, h1max, t_out, t_step, rx, rx1, smem_bytes, tmem_cols, nstages, resident, tma, nboxes, box_rows, u_rows}; returns 1 when
 * the shape can be planned, 0 when not.
 * vits_test_conv_plan: k_conv_tc for cin -> n channels with `ntaps` taps at offsets toff[], input as bf16 operand rows (xb != 0, xb_rows
 * rows) or fp32, split3 = the bf16x3 text-side form, over `ntiles` 128-row tiles on `num_sms` SMs.  out[16] receives
 * {ntile, rows_a, rows_need, a_bytes, slot_bytes, nstages, resident, nabuf, smem_bytes, tmem_cols, tma, nboxes, box_rows, nepi, nload, naccbuf}. */
int vits_test_mrf3_plan(int C, int nrb, const int* k, const int* d1, const int* d2, int rb1, int up_cin, int nb_pref, int fuse_post,
                        int use_tma, int min_hmax, int* out);
int vits_test_conv_plan(int cin, int n, int ntaps, const int* toff, int xb, long xb_rows, int split3, int ntiles, int num_sms, int* out);

#ifdef __cplusplus
}
#endif
#endif /* VITS_B200_TEST_H */
