/*
 * vits_b200.h -- C ABI of the B200-native VITS synthesis engine (libvits_b200.so).
 *
 * Drop-in boundary: this library replaces the third-party
 * onnxruntime.InferenceSession that phoonnx's TTSVoice holds
 * (reference: phoonnx/voice.py:105-107 field, :167-171 construction, :347 get_inputs(),
 * :374-377 run()).  ORT's own C API is NOT re-implemented; these entry points are what a
 * ctypes / cffi binding on the reference side calls instead, and what
 * phoonnx_b200/engine.py binds.  Plain pointers and sizes only -- no torch / C++ types.
 *
 * All functions return 0 on success, a negative VITS_E_* code on failure; the message is
 * available from vits_last_error().  There is NO CPU fallback: every entry point that
 * computes fails with VITS_E_CUDA when no sm_100 device is usable.
 *
 * Threading: one handle == one GPU + one stream.  Calls on one handle are serialised by
 * an internal mutex; different handles are fully concurrent (ORT's run() is re-entrant,
 * voice.py:374 is called serially by TTSVoice.synthesize).
 */
#ifndef VITS_B200_H
#define VITS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITS_B200_ABI_VERSION 2

enum {
    VITS_OK = 0,
    VITS_E_INVALID = -1,  /* bad argument (shape, id >= n_vocab, sid >= n_speakers, ...) -> ValueError  */
    VITS_E_CUDA = -2,     /* CUDA runtime failure / no device                              -> RuntimeError */
    VITS_E_STATE = -3,    /* call order violated (decode before prepare, missing tensor)   -> RuntimeError */
    VITS_E_NOMEM = -4
};

typedef struct vits_handle vits_handle;

/* Architecture of the voice.  The exported file carries no hyper-parameters
 * (export_onnx.py:335-345 stores only n_speakers/n_vocab/sample_rate), so the host-side
 * loader infers them from initializer shapes and passes them here.
 * Field meaning follows phoonnx_train/vits/models.py:527-615. */
#define VITS_MAX_UPS 8
#define VITS_MAX_RBK 8
#define VITS_MAX_DIL 4
#define VITS_MAX_FLOWS 8
typedef struct vits_arch {
    int32_t n_vocab, hidden, inter, filter, n_heads, n_layers, enc_kernel, window;
    int32_t n_speakers, gin;
    int32_t use_sdp, dp_filter, dp_kernel, dds_layers, n_cflows, cflows[VITS_MAX_FLOWS], num_bins;
    int32_t n_flow, flow_layers[VITS_MAX_FLOWS], wn_layers, wn_kernel, wn_dilation_rate;
    int32_t resblock_type;                 /* 1 = ResBlock1 (high), 2 = ResBlock2          */
    int32_t n_ups, up_rates[VITS_MAX_UPS], up_kernels[VITS_MAX_UPS], up_init;
    int32_t n_rbk, rb_kernels[VITS_MAX_RBK], rb_ndil[VITS_MAX_RBK], rb_dilations[VITS_MAX_RBK][VITS_MAX_DIL];
    int32_t sample_rate;
} vits_arch;

/* ABI / build information: returns VITS_B200_ABI_VERSION. */
int vits_abi_version(void);

/* What a caller can ask a loaded voice (the reference reads the same facts from the voice's JSON config and from
 * session.get_inputs(), voice.py:347: whether a `sid` input exists; phoonnx/config.py sample_rate / num_speakers). */
typedef struct vits_info {
    int32_t n_vocab, n_speakers, has_sid;      /* has_sid: the feed needs `sid` (multi-speaker voice, voice.py:370)      */
    int32_t hidden, inter;                     /* encoder width, latent channels (noise_z rows)                          */
    int32_t sample_rate, hop;                  /* audio samples per frame: utterance b has y_lengths[b] * hop samples    */
    int32_t resblock_type, use_sdp;
    int32_t precision;                         /* 0 fp32 CUDA cores, 1 bf16 tcgen05                                      */
    int32_t device, num_sms, finalized;
    int32_t has_scales, has_langid;            /* inputs the voice's own graph declares (voice.py:358, :369); vits_create handles: 1, 0 */
    int32_t reserved[6];
} vits_info;
int vits_describe(vits_handle* h, vits_info* info);

/* Replaces InferenceSession(str(model_path), ...) (voice.py:167-171) in ONE call: opens a voice file written by
 * phoonnx_train/export_onnx.py (plain or gzip-compressed .onnx), recovers the architecture from its tensors (the file carries no
 * hyper-parameters), packs the kernel layouts, uploads them to device `device_id` and finalizes the handle -- no host-language
 * helper involved (csrc/voice_file.h; pinned blob for blob against the Python loader by tests/test_native_loader.py).
 * precision: 0 fp32 CUDA cores (parity), 1 bf16 tcgen05.  On failure *out is NULL and, when `err` is given, a message of at most
 * err_cap bytes explains why (bad file -> VITS_E_INVALID, no sm_100 device -> VITS_E_CUDA). */
int vits_open(const char* path, int device_id, int precision, vits_handle** out, char* err, size_t err_cap);

/* Lower-level construction for hosts that parse and pack the file themselves (phoonnx_b200/weights.py + packing.py do, for
 * checkpoints and state dicts): binds device `device_id`, creates the stream and an empty weight table; then vits_upload per
 * blob, vits_set_option, vits_finalize. */
int vits_create(const vits_arch* arch, int device_id, vits_handle** out);

/* Upload one packed weight blob (already in kernel layout; packing is done by the
 * host-side loader phoonnx_b200/packing.py from the exported initializers).
 * dtype: 0 = float32, 1 = bfloat16 (raw uint16), 2 = int32. */
int vits_upload(vits_handle* h, const char* name, const void* data, size_t nbytes, int dtype);

/* Resolve every tensor the architecture needs; fails with VITS_E_STATE naming the first
 * missing blob. */
int vits_finalize(vits_handle* h);

/* Options.  "precision": 0 = fp32 CUDA cores everywhere (parity mode),
 *                        1 = bf16 tcgen05 tensor cores for the decoder / flow contractions
 *                            (fp32 accumulate, fp32 everywhere else).
 *           "max_chunk_frames": frame budget of one decoder pass (workspace sizing). */
int vits_set_option(vits_handle* h, const char* key, double value);

/* ---- the hot path: replaces session.run(None, feed) (voice.py:374-377) --------------
 *
 * Phase 1 (text side): embedding, text encoder, duration predictor, length regulation.
 *   ids          packed phoneme ids, sum(lengths) values      ("input", voice.py:350)
 *   lengths      [B]                                          ("input_lengths", voice.py:351)
 *   scales       {noise_scale, length_scale, noise_w}         ("scales", voice.py:364-367)
 *   sid          [B] or NULL (single-speaker)                 ("sid", voice.py:370)
 *   noise_dp     test-only: [B][2][dp_stride] N(0,1) samples replacing the
 *                RandomNormalLike of models.py:111, or NULL -> Philox(seed)
 *   logw_override test-only: packed [sum(lengths)] log-durations replacing the duration
 *                predictor output (bit-exactness tests of the integer path), or NULL
 * Outputs: total_frames and, per utterance, y_lengths[B] (frames).  Audio length of
 * utterance b is y_lengths[b] * hop samples.
 */
int vits_prepare(vits_handle* h, const int64_t* ids, const int64_t* lengths, int32_t B,
                 const float scales[3], const int64_t* sid, const float* noise_dp,
                 int64_t dp_stride, const float* logw_override, uint64_t seed,
                 int64_t* y_lengths, int64_t* total_frames);

/* Phase 2 (frame side): prior expansion + sampling, coupling flow (reverse), HiFi-GAN.
 *   noise_z      test-only: [B][inter][z_stride] N(0,1) samples replacing randn_like of
 *                models.py:718, or NULL -> Philox(seed)
 *   out_kind     0: keep audio on the device only (throughput timing)
 *                1: float32 to host buffer `out` (packed, utterance b at sample offset
 *                   hop * sum(y_lengths[:b]))
 *                2: int16 to host buffer after the caller-side post-processing of
 *                   voice.py:271-282 + AudioChunk.audio_int16_array voice.py:88-91
 *                   (peak-normalise per utterance, volume, clip, x32767) done on device
 *   out_capacity number of samples `out` can hold (>= hop * total_frames)
 */
int vits_decode(vits_handle* h, const float* noise_z, int64_t z_stride, int32_t out_kind,
                void* out, int64_t out_capacity, float volume, int32_t normalize);
/*   out_kind     3: float32 to a DEVICE buffer `out` of the caller (stream-ordered copy on the handle's stream, no host
 *                   synchronisation: pair it with vits_set_stream() to stay inside the caller's own CUDA stream)
 *   | VITS_OUT_ASYNC (kinds 1 and 2): return once the device->host transfer is enqueued on the copy stream; the result is complete
 *                   after vits_wait_ticket(h, vits_output_ticket(h)).  Per call, no session-wide mode: a blocking call made
 *                   while an asynchronous result is in flight neither disturbs it nor returns early itself. */
#define VITS_OUT_ASYNC 0x100
int64_t vits_output_ticket(vits_handle* h);                 /* ticket of the most recent vits_decode with host output (>= 1) */
int vits_wait_ticket(vits_handle* h, int64_t ticket);       /* returns when that call's host buffer is complete               */

/* Samples the output buffer of vits_decode must hold.  sum_ids <= 0: EXACT figure of the last successful vits_prepare
 * (total_frames * hop; VITS_E_STATE before any).  sum_ids > 0: a planning figure for that many phoneme ids at
 * `length_scale` BEFORE vits_prepare has run (durations are data: 16 frames per id is an allowance, not a bound --
 * size the real buffer from vits_prepare's total_frames). */
int64_t vits_max_output_samples(vits_handle* h, int64_t sum_ids, float length_scale);

/* Run every later call of this handle on the caller's CUDA stream (a cudaStream_t passed as void*; NULL restores the
 * handle's own stream).  Pending work on the previous stream is completed first. */
int vits_set_stream(vits_handle* h, void* cuda_stream);

/* Output buffers.  onnxruntime allocates the array run() returns (voice.py:374-377); here the caller does, and a
 * page-locked buffer from vits_host_alloc() makes the device->host transfer a true asynchronous DMA on the handle's
 * copy stream: with option "async_output" = 1, vits_decode(out_kind = 1) returns once the transfer is enqueued and
 * vits_wait_output() (or the second-next vits_decode on the same handle) completes it, so the transfer of one batch
 * overlaps the kernels of the next.  Without the option vits_decode() blocks until `out` is complete. */
/* Scatter output: the NEXT vits_decode with a host output (out_kind 1 or 2) writes utterance b's samples at out + sample_offsets[b]
 * instead of packing the utterances back to back -- one DMA per utterance, so the device's copy engine places every utterance
 * where the caller wants it (original order of a larger job, a result buffer shared by several processes, ...).  One-shot; B must
 * equal the prepared batch; capacity is checked per utterance.  vits_host_register page-locks memory the caller already owns
 * (e.g. a POSIX shared-memory segment) so that those DMAs are asynchronous. */
int vits_set_output_offsets(vits_handle* h, const int64_t* sample_offsets, int32_t B);
int vits_host_register(void* p, size_t nbytes);
int vits_host_unregister(void* p);
int vits_host_alloc(size_t nbytes, void** out);
int vits_host_free(void* p);
int vits_wait_output(vits_handle* h, int older_only);   /* older_only: leave the most recent vits_decode's transfer in flight */

/* Test / debugging: copy a stage tensor of the last prepare/decode to the host.
 * names: "x" [sumT,H], "stats" [sumT,2C] (m_p | logs_p), "logw" [sumT], "durations" / "cum" (int32
 * bits, [sumT]), "noise_dp" [2][sumT] (option debug_keep_noise_dp); per-chunk workspaces "frame_index"
 * (int32 bits), "z_p" / "z" [frames, C] -- available only when the last vits_decode ran as ONE chunk
 * (VITS_E_STATE otherwise).  Returns the element count written (<= capacity) or <0. */
int64_t vits_fetch(vits_handle* h, const char* name, void* out, int64_t capacity_elems);

/* Device-side timing of everything enqueued between start and stop on the handle's
 * stream (CUDA events on the launching stream). */
int vits_timer_start(vits_handle* h);
int vits_timer_stop(vits_handle* h, float* elapsed_ms);

/* Counters since creation: number of kernels of THIS library launched. */
int64_t vits_launch_count(vits_handle* h);

/* Device time (ms, CUDA events) of the decoder (HiFi-GAN) part of the last vits_decode
 * calls since the last vits_timer_start, for the roofline of the dominant kernel family. */
int vits_stage_ms(vits_handle* h, float* text_ms, float* flow_ms, float* dec_ms);

/* Device time (ms, CUDA events around the kernel launches alone), launch count and ALGORITHMIC multiply-accumulates (halo recompute
 * and padding excluded) of one fused decoder kernel since the last vits_timer_start: which = 0 the last generator stage
 * (ConvTranspose1d + multi-receptive-field ResBlocks + lrelu / conv_post / tanh, models.py:352-366, one kernel), 1 the other fused
 * multi-receptive-field stages.  The measurement behind bench.py's per-kernel roofline. */
int vits_kernel_ms(vits_handle* h, int which, float* ms, int64_t* launches, double* macs);

const char* vits_last_error(vits_handle* h);
void vits_destroy(vits_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* VITS_B200_H */
