/* A complete host of libvits_b200.so in plain C: no Python, no torch, no onnxruntime.
 *
 *   gcc -O2 -Iinclude examples/synth.c -Lphoonnx_b200 -lvits_b200 -Wl,-rpath,$PWD/phoonnx_b200 -o synth
 *   ./synth voice.onnx 1,0,20,0,59,0,24,0,2 out.pcm [noise_scale length_scale noise_w]
 *
 * What phoonnx's TTSVoice does around its onnxruntime session (voice.py:328-379 feed construction, :374 run, :271-282 + :88-91
 * int16 post-processing), through the C ABI: open the exported voice, ask it what it is, synthesise one utterance of phoneme ids
 * to 16-bit PCM. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vits_b200.h"

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s voice.onnx id,id,... out.pcm [noise_scale length_scale noise_w]\n", argv[0]); return 2; }
    char err[512] = "";
    vits_handle* h = NULL;
    int rc = vits_open(argv[1], 0, /*precision: bf16 tensor cores*/ 1, &h, err, sizeof err);
    if (rc != VITS_OK) { fprintf(stderr, "vits_open failed (%d): %s\n", rc, err); return 1; }
    vits_info info;
    vits_describe(h, &info);
    int64_t ids[4096]; int64_t n = 0;
    for (char* tok = strtok(argv[2], ","); tok && n < 4096; tok = strtok(NULL, ",")) ids[n++] = atoll(tok);
    float scales[3] = {0.667f, 1.0f, 0.8f};                 /* config.py:9-11 defaults; order noise, length, noise_w (voice.py:364-367) */
    for (int i = 0; i < 3 && 4 + i < argc; i++) scales[i] = (float)atof(argv[4 + i]);
    int64_t sid = 0, ylen = 0, frames = 0;
    rc = vits_prepare(h, ids, &n, 1, scales, info.has_sid ? &sid : NULL, NULL, 0, NULL, /*seed*/ 1, &ylen, &frames);
    if (rc != VITS_OK) { fprintf(stderr, "vits_prepare failed (%d): %s\n", rc, vits_last_error(h)); vits_destroy(h); return 1; }
    const int64_t samples = vits_max_output_samples(h, 0, 0.f);          /* exact: hop * frames of the prepare above */
    int16_t* pcm = (int16_t*)malloc((size_t)samples * sizeof(int16_t));
    rc = vits_decode(h, NULL, 0, /*int16, peak-normalised on the device*/ 2, pcm, samples, 1.0f, 1);
    if (rc != VITS_OK) { fprintf(stderr, "vits_decode failed (%d): %s\n", rc, vits_last_error(h)); vits_destroy(h); return 1; }
    FILE* f = fopen(argv[3], "wb");
    if (!f || fwrite(pcm, sizeof(int16_t), (size_t)samples, f) != (size_t)samples) { fprintf(stderr, "cannot write %s\n", argv[3]); return 1; }
    fclose(f);
    printf("%lld ids -> %lld frames -> %lld samples at %d Hz (%d speakers, hop %d, %lld kernels launched)\n", (long long)n, (long long)frames,
           (long long)samples, info.sample_rate, info.n_speakers, info.hop, (long long)vits_launch_count(h));
    free(pcm);
    vits_destroy(h);
    return 0;
}
