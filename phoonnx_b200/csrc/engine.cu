// libvits_b200: host-side orchestration of the VITS phoneme-ids -> audio hot path and its C ABI
// (include/vits_b200.h).  What it replaces: the onnxruntime InferenceSession.run() call at
// phoonnx/voice.py:374-377, whose arithmetic is SynthesizerTrn.infer
// (phoonnx_train/vits/models.py:681-722).
//
// Execution model: one handle = one GPU = one stream.  All activations are packed varlen,
// channel-last fp32 ([rows, C]); utterances never see each other's samples (B=1 semantics per
// utterance, zero padding at utterance edges).  Phase 1 (vits_prepare) runs the text side and
// ends with the only host sync of the path: the per-utterance frame counts.  Phase 2
// (vits_decode) runs the frame side in chunks bounded by a frame budget.
#include <cuda.h>
#include "../../include/vits_b200.h"
#include "../../include/vits_b200_test.h"
#include "common.cuh"
#include "kernels_f32.cuh"
#include "attention.cuh"
#include "conv_tc.cuh"
#include "mrf3_tc.cuh"
#include "probe_tc.cuh"
#include "voice_file.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

struct DevBlob { void* p = nullptr; size_t bytes = 0; int dtype = 0; };

struct ConvP {
    const float* w = nullptr; const __nv_bfloat16* wtc = nullptr; const float* b = nullptr;
    int cin = 0, n = 0, npad = 0, npad16 = 0, ntaps = 0; int toff[CONV_MAX_TAPS] = {0};
    // fp32-faithful bf16x3 operands (text side): one blob per K slice of slice_cin input channels
    const __nv_bfloat16* wtc3[CONV_MAX_SLICES] = {nullptr}; int nsl = 0, slice_cin = 0;
    // N-tiled copies (n > 256): same operands re-laid [n / wtile][tap][cin/8][wtile][8] at load time (k_retile_weights)
    int wtile = 0; const __nv_bfloat16* wtc_t = nullptr; const __nv_bfloat16* wtc3_t[CONV_MAX_SLICES] = {nullptr};
};

struct LnP { const float* g = nullptr; const float* b = nullptr; };
struct DdsP { const float* dw_w; const float* dw_b; LnP ln1; ConvP pw; LnP ln2; int dil; };

struct Buf {   // grow-only device buffer
    void* p = nullptr; size_t cap = 0;
};

#define SPLIT3_MAX_CIN 96      // packing.SPLIT3_MAX_CIN
struct StagePair { cudaEvent_t a, b; int stage; };

}  // namespace

struct vits_handle {
    vits_arch A;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::string err;
    std::map<std::string, DevBlob> blobs;
    std::map<std::string, double> opts;
    bool finalized = false;
    int precision = 0;
    int text_tc = 1;                 // bf16 mode: text-side GEMMs as bf16x3 on tcgen05 (0: fp32 CUDA cores)
    int64_t max_chunk_frames = 262144;   // frames per decode chunk (r01: larger chunks amortise tails; workspaces grow on demand)
    int64_t launches = 0;
    int num_sms = 148;

    // resolved weights
    const float* emb = nullptr;
    struct EncL { ConvP qkv, o, ffn1, ffn2; const float* rel_k; const float* rel_v; LnP ln1, ln2; };
    std::vector<EncL> enc;
    ConvP enc_proj;
    ConvP dp_pre, dp_proj; std::vector<DdsP> dp_dds; const float* dp_cond_tab = nullptr;
    struct CFlow { const float* pre_w; const float* pre_b; std::vector<DdsP> dds; ConvP proj; };
    std::vector<CFlow> cflows;
    float ea_m = 0.f, ea_logs = 0.f;
    ConvP dpd_c1, dpd_c2, dpd_proj; LnP dpd_n1, dpd_n2;
    struct FlowS { ConvP pre, post; std::vector<ConvP> in, rs; std::vector<const float*> cond_tab; int xcol, ocol;
                   std::vector<ConvP> rsr, mskip; const float* mskip_b = nullptr; bool v2 = false; };   // tensor-core tail (packing.py "mskip")
    std::vector<FlowS> flows;
    ConvP dec_pre; const float* dec_cond_tab = nullptr;
    struct UpS { ConvP A, B; int rate, cout; };
    std::vector<UpS> ups;
    std::vector<std::vector<ConvP>> rb_c1, rb_c2;   // [resblock][conv]
    const float* post_w = nullptr; int post_c = 0;

    // text-side state of the last prepare
    int B = 0; int R = 0;            // utterances, total ids
    std::vector<int> h_cu_t, h_ylen, h_cu_y; std::vector<int> h_sid;
    float scales[3] = {0.667f, 1.f, 0.8f};
    uint64_t seed = 0, utt_counter = 0, utt_base = 0;
    bool prepared = false;
    int64_t total_frames = 0;
    int last_chunk_frames = 0;
    int last_nchunks = 0;             // chunks of the last vits_decode (chunk tensors can only be fetched when it was one)
    cudaStream_t own_stream = nullptr;   // the stream vits_create made (h->stream is this one unless vits_set_stream gave another)
    std::vector<int64_t> out_offsets;   // one-shot per-utterance destination offsets of the next host-output vits_decode (vits_set_output_offsets)
    int graph_has_scales = 1, graph_has_langid = 0;   // inputs the file's graph declares (vits_open)
    int64_t ticket = 0;               // number of vits_decode calls that produced host output so far
    int64_t ticket_of[2] = {0, 0};    // ticket whose transfer ev_out[i] tracks

    Buf ids, cu_t, tile_t, sid, x, y, qkv, att, ffn, stats, d0, d1, gdp, hp, z0, z1, logw, dur, cum, ylen;
    Buf inj_dp, inj_z, dbg_dp, audio16_alt;
    Buf chunk_meta, tdesc, P, fh, facts, fskip, fidx, dpre, sX, sT1, sYa, sYb, sXSa, sXSb, audio, peaks, audio16, cu_y_dev, dbg_zp, mrf_dbg, conv_dbg;
    int conv_counter = 0;

    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    // output path: the audio of one vits_decode lands in one of two device buffers and leaves through a copy stream, so the
    // device->host DMA of call k overlaps the kernels of call k+1 (and chunk c's DMA those of chunk c+1)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_chunk = nullptr, ev_out[2] = {nullptr, nullptr};
    bool out_pending[2] = {false, false};
    int audio_sel = 0;
    Buf audio_alt, facts_b, tdesc_t, tdesc_c, sX1b, rowpos, fpos;
    std::vector<void*> owned;         // device allocations made at finalize time (re-tiled weight copies)
    std::vector<float*> rb_b2sum;     // per stage: sum over resblocks of the second conv's bias (ResBlock2; fused conv2 launch)
    std::vector<StagePair> stage_events;
    std::vector<cudaEvent_t> event_pool;
    float stage_ms[5] = {0, 0, 0, 0, 0};      // text, flow, decoder; [3] the fused last-stage kernel alone, [4] the other fused MRF stage kernels
    size_t open_stage = 0;
    int conv_text_counter = 0;
    int64_t kern_launches[2] = {0, 0};        // launches / algorithmic MACs behind stage_ms[3], [4] (vits_kernel_ms)
    double kern_macs[2] = {0, 0};
};

namespace {

int fail(vits_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define CK(h, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
    return fail(h, VITS_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)

int ensure(vits_handle* h, Buf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    size_t want = std::max(bytes, b.cap + b.cap / 2);
    want = (want + 255) & ~size_t(255);
    if (b.p) { CK(h, cudaStreamSynchronize(h->stream)); CK(h, cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    CK(h, cudaMalloc(&b.p, want));
    b.cap = want;
    return 0;
}
template <class T> T* ptr(Buf& b) { return reinterpret_cast<T*>(b.p); }

const DevBlob* find_blob(vits_handle* h, const std::string& name) {
    auto it = h->blobs.find(name);
    return it == h->blobs.end() ? nullptr : &it->second;
}

int need_f32(vits_handle* h, const std::string& name, const float** out, size_t min_elems) {
    const DevBlob* b = find_blob(h, name);
    if (!b) return fail(h, VITS_E_STATE, "missing weight blob '%s'", name.c_str());
    if (b->dtype != 0 || b->bytes < min_elems * sizeof(float))
        return fail(h, VITS_E_STATE, "weight blob '%s': wrong dtype/size (%zu bytes, need %zu)", name.c_str(), b->bytes,
                    min_elems * sizeof(float));
    *out = reinterpret_cast<const float*>(b->p);
    return 0;
}

int rup(int v, int m) { return (v + m - 1) / m * m; }

// Resolve a conv: blobs "<name>.w" [ntaps][cin][npad] fp32, optional "<name>.b" [npad],
// optional "<name>.wtc" bf16 [ntaps][cin/8][npad16][8].
int mkconv(vits_handle* h, ConvP& c, const std::string& name, int cin, int n, const std::vector<int>& toff, bool bias) {
    c.cin = cin; c.n = n; c.npad = rup(n, 4); c.npad16 = rup(n, 16); c.ntaps = (int)toff.size();
    if (c.ntaps > CONV_MAX_TAPS) return fail(h, VITS_E_INVALID, "conv '%s': %d taps > %d", name.c_str(), c.ntaps, CONV_MAX_TAPS);
    if (cin % 4) return fail(h, VITS_E_INVALID, "conv '%s': cin %d not a multiple of 4", name.c_str(), cin);
    for (int i = 0; i < c.ntaps; i++) c.toff[i] = toff[i];
    int rc = need_f32(h, name + ".w", &c.w, (size_t)c.ntaps * cin * c.npad);
    if (rc) return rc;
    c.b = nullptr;
    if (bias) { rc = need_f32(h, name + ".b", &c.b, c.npad); if (rc) return rc; }
    const DevBlob* t = find_blob(h, name + ".wtc");
    c.wtc = nullptr;
    if (t && t->dtype == 1 && cin % 16 == 0 && t->bytes >= (size_t)c.ntaps * cin * c.npad16 * 2)
        c.wtc = reinterpret_cast<const __nv_bfloat16*>(t->p);
    // K slices of the bf16x3 form: largest divisor of cin that is a multiple of 16 and <= 96 (packing.split3_slice): the hi + lo
    // activation planes of one slice, double-buffered, fit one CTA's shared memory
    c.nsl = 0; c.slice_cin = 0;
    if (cin % 16 == 0 && n % 16 == 0) {
        int sl = 0;
        for (int v = std::min(cin, SPLIT3_MAX_CIN); v >= 16; v--) if (cin % v == 0 && v % 16 == 0) { sl = v; break; }
        if (sl && cin / sl <= CONV_MAX_SLICES) {
            bool all = true;
            for (int j = 0; j < cin / sl; j++) {
                const DevBlob* t3 = find_blob(h, name + ".wtc3." + std::to_string(j));
                if (!t3 || t3->dtype != 1 || t3->bytes < (size_t)c.ntaps * 3 * sl * c.npad16 * 2) { all = false; break; }
                c.wtc3[j] = reinterpret_cast<const __nv_bfloat16*>(t3->p);
            }
            if (all) { c.nsl = cin / sl; c.slice_cin = sl; }
        }
    }
    // N-tiled copies for n > 256 (first-choice N tile of conv_tc_plan: the largest multiple-of-16 divisor of npad16 that is <= 256)
    c.wtile = 0; c.wtc_t = nullptr;
    for (auto& q : c.wtc3_t) q = nullptr;
    if (c.npad16 > 256) {
        int wt = 0;
        for (int cand = 256; cand >= 16; cand -= 16) if (c.npad16 % cand == 0) { wt = cand; break; }
        auto retile = [&](const __nv_bfloat16* src, long rows, const __nv_bfloat16** dst) -> int {
            __nv_bfloat16* d = nullptr;
            CK(h, cudaMalloc(&d, (size_t)rows * c.npad16 * 16));
            h->owned.push_back(d);
            const long units = rows * c.npad16;
            launch_k(k_retile_weights, (unsigned)((units + 255) / 256), 256, 0, h->stream, src, d, rows, c.npad16, wt);
            CK(h, cudaGetLastError());
            *dst = d;
            return 0;
        };
        if (wt) {
            c.wtile = wt;
            if (c.wtc && (rc = retile(c.wtc, (long)c.ntaps * (cin / 8), &c.wtc_t))) return rc;
            for (int j = 0; j < c.nsl; j++)
                if ((rc = retile(c.wtc3[j], (long)c.ntaps * 3 * (c.slice_cin / 8), &c.wtc3_t[j]))) return rc;
            CK(h, cudaStreamSynchronize(h->stream));
        }
    }
    return 0;
}

std::vector<int> sym_taps(int k, int d) {   // Conv1d with "same" padding (k*d - d)/2, commons.py:17-18
    std::vector<int> t(k);
    for (int i = 0; i < k; i++) t[i] = (i - (k - 1) / 2) * d;
    return t;
}

int mkln(vits_handle* h, LnP& l, const std::string& name, int C) {
    int rc = need_f32(h, name + ".g", &l.g, C); if (rc) return rc;
    return need_f32(h, name + ".b", &l.b, C);
}

int mkdds(vits_handle* h, std::vector<DdsP>& v, const std::string& name, int C) {
    const vits_arch& A = h->A;
    v.resize(A.dds_layers);
    int dil = 1;
    for (int i = 0; i < A.dds_layers; i++) {
        DdsP& d = v[i];
        std::string p = name + "." + std::to_string(i);
        int rc;
        if ((rc = need_f32(h, p + ".dw_w", &d.dw_w, (size_t)A.dp_kernel * C))) return rc;
        if ((rc = need_f32(h, p + ".dw_b", &d.dw_b, C))) return rc;
        if ((rc = mkln(h, d.ln1, p + ".ln1", C))) return rc;
        if ((rc = mkconv(h, d.pw, p + ".pw", C, C, {0}, true))) return rc;
        if ((rc = mkln(h, d.ln2, p + ".ln2", C))) return rc;
        d.dil = dil;
        dil *= A.dp_kernel;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------
struct Tiles { const int* cu; const int* t64; const int* t128; const int* t256; const int* tx; int n64, n128, n256, nx, tmx; int B; int rate;
               const int4* d128; long rows; };   // rows: all rows of the arrays these tiles cover (sum of the utterances' lengths x rate)   // d128: per-128-row-tile descriptors for the tcgen05 conv kernel (make_d128)

ConvArgs base_args(const ConvP& c, const float* x, int ldx, int xcol, float* out, int ldo, int ocol) {
    ConvArgs a;
    memset(&a, 0, sizeof a);
    a.x = x; a.ldx = ldx; a.xcol = xcol; a.cin = c.cin;
    a.ntaps = c.ntaps; for (int i = 0; i < c.ntaps; i++) a.toff[i] = c.toff[i];
    a.w = c.w; a.wtc = c.wtc; a.n = c.n; a.npad = c.npad; a.npad16 = c.npad16; a.bias = c.b;
    a.wtile = c.wtc_t ? c.wtile : 0; a.wtc_t_ks[0] = c.wtc_t;
    a.in_act = 0; a.in_slope = 1.f; a.epi = EPI_STORE; a.out_act = ACT_NONE; a.out_div = 1.f;
    a.out = out; a.ldo = ldo; a.ocol = ocol;
    return a;
}

int launch_conv(vits_handle* h, ConvArgs& a, const Tiles& T, bool allow_tc) {
    a.cu = T.cu; a.B = T.B; a.rate = T.rate;
    a.split3 = 0;
    if (a.nks <= 1) { a.nks = 1; a.wtc_ks[0] = a.wtc; }
    if (allow_tc && h->precision == 1 && conv_tc_supported(a)) {
        a.tile_cu = T.t128; a.ntiles = T.n128; a.tdesc = T.d128;
        a.xb_rows = a.xb ? T.rows : 0;          // the TMA loader's tensor map covers exactly the rows these tiles address
        if (T.n128 == 0) return 0;
        if (!T.d128) return fail(h, VITS_E_STATE, "conv_tc: tile descriptors missing for rate %d", T.rate);
        a.dbg = nullptr;
        if (h->opts.count("conv_dbg") && (int)h->opts["conv_dbg"] == ++h->conv_counter) {
            int rc = ensure(h, h->conv_dbg, (size_t)TC_DBG_TILES * 16 * 8);
            if (rc) return rc;
            CK(h, cudaMemsetAsync(h->conv_dbg.p, 0, (size_t)TC_DBG_TILES * 16 * 8, h->stream));
            a.dbg = ptr<unsigned long long>(h->conv_dbg);
        }
        cudaError_t e = conv_tc_launch(a, h->num_sms, h->stream, h->opts["conv_cluster"] != 0);
        if (e != cudaSuccess) return fail(h, VITS_E_CUDA, "conv_tc launch: %s", cudaGetErrorString(e));
        h->launches++;
        return 0;
    }
    if (a.xb || a.outb) return fail(h, VITS_E_STATE, "bf16 operand rows need the tcgen05 conv path (cin %d, n %d)", a.cin, a.n);
    a.tile_cu = T.t64; a.ntiles = T.n64;
    if (T.n64 == 0) return 0;
    dim3 grid(T.n64, (a.n + CF_TN - 1) / CF_TN);
    launch_k(k_conv_f32, grid, 256, 0, h->stream, a);
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

// Text-side convolution (encoder / duration predictor): in bf16 mode it runs on the tensor cores as an fp32-faithful
// bf16x3 product, K-sliced so the hi+lo activation planes fit shared memory; slices after the first accumulate.
int launch_conv_text(vits_handle* h, const ConvP& c, ConvArgs& a, const Tiles& T) {
    const bool want = h->precision == 1 && h->text_tc && c.nsl > 0 && a.epi == EPI_STORE && a.ldx % 4 == 0 && a.xcol % 4 == 0;
    if (!want) return launch_conv(h, a, T, false);
    a.cu = T.cu; a.B = T.B; a.rate = T.rate;
    a.tile_cu = T.t128; a.ntiles = T.n128; a.tdesc = T.d128;
    if (T.n128 == 0) return 0;
    if (!T.d128) return fail(h, VITS_E_STATE, "conv_tc: tile descriptors missing for rate %d", T.rate);
    // one launch: the kernel loops over the K slices and accumulates them in TMEM
    ConvArgs s = a;
    s.split3 = 1; s.cin = c.slice_cin; s.nks = c.nsl; s.wtc = c.wtc3[0];
    s.wtile = c.wtc3_t[0] ? c.wtile : 0;
    for (int j = 0; j < c.nsl; j++) { s.wtc_ks[j] = c.wtc3[j]; s.wtc_t_ks[j] = c.wtc3_t[j]; }
    s.dbg = nullptr;
    if (h->opts.count("conv_text_dbg") && (int)h->opts["conv_text_dbg"] == ++h->conv_text_counter) {
        int rc = ensure(h, h->conv_dbg, (size_t)TC_DBG_TILES * 16 * 8);
        if (rc) return rc;
        CK(h, cudaMemsetAsync(h->conv_dbg.p, 0, (size_t)TC_DBG_TILES * 16 * 8, h->stream));
        s.dbg = ptr<unsigned long long>(h->conv_dbg);
    }
    cudaError_t e = conv_tc_launch(s, h->num_sms, h->stream, h->opts["conv_cluster"] != 0);
    if (e != cudaSuccess) return fail(h, VITS_E_CUDA, "conv_tc (bf16x3) launch: %s", cudaGetErrorString(e));
    h->launches++;
    return 0;
}

int launch_ln(vits_handle* h, const float* in, float* out, const LnP& ln, int rows, int C, int mode,
              const DdsP* dw, const Tiles& T) {
    if (C % 32 || C > 32 * LN_MAXV) return fail(h, VITS_E_INVALID, "LayerNorm width %d unsupported (multiple of 32, <= %d)", C, 32 * LN_MAXV);
    if (rows == 0) return 0;
    if (C % 64 == 0 && C <= 256 && (mode != 1 || h->A.dp_kernel == 3) && h->opts["ln_scalar"] == 0) {
        // 64-bit accesses (kernels_f32.cuh k_layernorm_v2); every exported voice's widths (192, 256) qualify, x_low's 96 does not
#define LN2_LAUNCH(N_) launch_k(k_layernorm_v2<N_>, (rows + 3) / 4, 128, 0, h->stream, in, out, ln.g, ln.b, rows, mode, dw ? dw->dw_w : nullptr, \
                           dw ? dw->dw_b : nullptr, dw ? dw->dil : 1, ptr<int2>(h->rowpos))
        switch (C / 64) { case 1: LN2_LAUNCH(1); break; case 2: LN2_LAUNCH(2); break; case 3: LN2_LAUNCH(3); break; default: LN2_LAUNCH(4); break; }
#undef LN2_LAUNCH
        h->launches++;
        CK(h, cudaGetLastError());
        return 0;
    }
    const int rpw = h->opts.count("ln_rpw") ? (int)h->opts["ln_rpw"] : 1;
#define LN_LAUNCH(R_) launch_k(k_layernorm<R_>, (rows + 4 * R_ - 1) / (4 * R_), 128, 0, h->stream, in, out, ln.g, ln.b, rows, C, mode, \
                          dw ? dw->dw_w : nullptr, dw ? dw->dw_b : nullptr, h->A.dp_kernel, dw ? dw->dil : 1, ptr<int2>(h->rowpos))
    if (rpw >= 4) LN_LAUNCH(4); else if (rpw == 2) LN_LAUNCH(2); else LN_LAUNCH(1);
#undef LN_LAUNCH
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

// DDSConv (modules.py:117-129): x updated in place; tmp1/tmp2 scratch [rows, C]
int run_dds(vits_handle* h, std::vector<DdsP>& L, float* x, float* tmp1, float* tmp2, int rows, int C, const Tiles& T) {
    for (auto& d : L) {
        int rc;
        if ((rc = launch_ln(h, x, tmp1, d.ln1, rows, C, 1, &d, T))) return rc;
        ConvArgs a = base_args(d.pw, tmp1, C, 0, tmp2, C, 0);
        if ((rc = launch_conv_text(h, d.pw, a, T))) return rc;
        if ((rc = launch_ln(h, tmp2, x, d.ln2, rows, C, 2, nullptr, T))) return rc;
    }
    return 0;
}

cudaEvent_t get_event(vits_handle* h) {
    if (!h->event_pool.empty()) { cudaEvent_t e = h->event_pool.back(); h->event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
void stage_begin(vits_handle* h, int stage) {
    StagePair sp; sp.a = get_event(h); sp.b = get_event(h); sp.stage = stage;
    cudaEventRecord(sp.a, h->stream);
    h->stage_events.push_back(sp);
    h->open_stage = h->stage_events.size() - 1;
}
void stage_end(vits_handle* h) { cudaEventRecord(h->stage_events[h->open_stage].b, h->stream); }
// a timed span nested inside a stage (one kernel): returns its slot for sub_end
size_t sub_begin(vits_handle* h, int stage) {
    StagePair sp; sp.a = get_event(h); sp.b = get_event(h); sp.stage = stage;
    cudaEventRecord(sp.a, h->stream);
    h->stage_events.push_back(sp);
    return h->stage_events.size() - 1;
}
void sub_end(vits_handle* h, size_t slot) { cudaEventRecord(h->stage_events[slot].b, h->stream); }
void resolve_stage_events(vits_handle* h) {
    for (auto& sp : h->stage_events) {
        float ms = 0.f;
        if (cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess)
            h->stage_ms[sp.stage] += ms;
        h->event_pool.push_back(sp.a); h->event_pool.push_back(sp.b);
    }
    h->stage_events.clear();
}

// build per-rate tile tables on the host and upload them in one copy
struct TileBuilder {
    std::vector<int> host;
    struct Ent { int rate; size_t o64, o128, o256, ox; int n64, n128, n256, nx, tmx; };
    std::vector<Ent> ents;
    size_t cu_off = 0; int B = 0;
    void begin(const int* cu_local, int B_) {
        B = B_; host.assign(cu_local, cu_local + B + 1); cu_off = 0; ents.clear();
    }
    void add(int rate, int tm_extra = 0) {
        Ent e; e.rate = rate; e.ox = 0; e.nx = 0; e.tmx = tm_extra;
        auto build = [&](int tm, int& n) {
            size_t off = host.size();
            int acc = 0;
            for (int b = 0; b < B; b++) {
                host.push_back(acc);
                long rows = (long)(host[cu_off + b + 1] - host[cu_off + b]) * rate;
                acc += (int)((rows + tm - 1) / tm);
            }
            host.push_back(acc); n = acc;
            return off;
        };
        e.o64 = build(64, e.n64); e.o128 = build(128, e.n128); e.o256 = build(256, e.n256);
        if (tm_extra > 0) e.ox = build(tm_extra, e.nx);
        ents.push_back(e);
    }
    Tiles get(const int* dev, int rate) const {
        for (auto& e : ents) if (e.rate == rate) {
            Tiles t; t.cu = dev + cu_off; t.t64 = dev + e.o64; t.t128 = dev + e.o128; t.t256 = dev + e.o256;
            t.n64 = e.n64; t.n128 = e.n128; t.n256 = e.n256; t.B = B; t.rate = rate;
            t.tx = dev + e.ox; t.nx = e.nx; t.tmx = e.tmx; t.d128 = nullptr; t.rows = (long)host[cu_off + B] * rate; return t;
        }
        Tiles t; memset(&t, 0, sizeof t); return t;
    }
};

// the fused MRF kernel with CUDA events around the kernel proper (the tile-descriptor helper launch stays outside)
cudaError_t mrf3_launch_timed(vits_handle* h, const Mrf3Args& m, const Mrf3Cfg& c, int stage) {
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    if (c.tma) {
        const bool modeU = m.up_u != 0;
        if (!make_rows_tmap(&tm, modeU ? (const void*)m.hb : (const void*)m.xb, m.in_rows, modeU ? m.up_cin : m.C, c.box_rows))
            return cudaErrorInvalidValue;
    }
    cudaError_t e = mrf3_tiles_launch(m, c, h->stream);
    if (e != cudaSuccess) return e;
    const size_t slot = sub_begin(h, stage);
    e = mrf3_kernel_launch(m, c, tm, h->num_sms, h->stream);
    sub_end(h, slot);
    return e;
}

// expand T's 128-row tile table into per-tile descriptors at `dst` (device, T.n128 entries) on the handle's stream
int make_d128(vits_handle* h, Tiles& T, int4* dst) {
    T.d128 = dst;
    if (T.n128 <= 0) return 0;
    launch_k(k_tile_desc, (T.n128 + 255) / 256, 256, 0, h->stream, T.cu, T.t128, T.B, T.rate, T.n128, TC_M, dst);
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int vits_abi_version(void) { return VITS_B200_ABI_VERSION; }

int vits_create(const vits_arch* arch, int device_id, vits_handle** out) {
    if (!arch || !out) return VITS_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device_id < 0 || device_id >= ndev) return VITS_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return VITS_E_CUDA;
    if (prop.major != 10) return VITS_E_CUDA;   // sm_100a only: no other code path exists
    vits_handle* h = new vits_handle();
    h->A = *arch; h->device = device_id; h->num_sms = prop.multiProcessorCount;
    if (cudaSetDevice(device_id) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h; return VITS_E_CUDA;
    }
    h->own_stream = h->stream;
    cudaEventCreate(&h->ev_t0); cudaEventCreate(&h->ev_t1);
    cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_chunk, cudaEventDisableTiming);
    for (int i = 0; i < 2; i++) cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming);
    *out = h;
    return VITS_OK;
}

int vits_open(const char* path, int device_id, int precision, vits_handle** out, char* err, size_t err_cap) {
    auto say = [&](const std::string& m) { if (err && err_cap) { snprintf(err, err_cap, "%s", m.c_str()); } };
    if (out) *out = nullptr;
    if (!path || !out || (precision != 0 && precision != 1)) { say("vits_open: null argument or precision not 0/1"); return VITS_E_INVALID; }
    vf::Voice v;
    std::string e;
    if (!vf::load_voice(path, v, e)) { say(e); return VITS_E_INVALID; }
    vits_handle* h = nullptr;
    int rc = vits_create(&v.arch, device_id, &h);
    if (rc != VITS_OK) { say("vits_open: no usable sm_100 CUDA device (this engine has no CPU fallback)"); return rc; }
    for (auto& kv : v.blobs)
        if ((rc = vits_upload(h, kv.first.c_str(), kv.second.bytes.data(), kv.second.bytes.size(), kv.second.dtype)) != VITS_OK) break;
    for (auto& kv : v.opts) if (rc == VITS_OK) rc = vits_set_option(h, kv.first.c_str(), kv.second);
    if (rc == VITS_OK) rc = vits_set_option(h, "precision", (double)precision);
    if (rc == VITS_OK) rc = vits_finalize(h);
    if (rc != VITS_OK) { say(h->err); vits_destroy(h); return rc; }
    auto has = [&](const char* n) { return std::find(v.model.inputs.begin(), v.model.inputs.end(), n) != v.model.inputs.end(); };
    h->graph_has_scales = has("scales") ? 1 : 0; h->graph_has_langid = has("langid") ? 1 : 0;
    *out = h;
    return VITS_OK;
}

int vits_test_mrf3_plan(int C, int nrb, const int* k, const int* d1, const int* d2, int rb1, int up_cin, int nb_pref, int fuse_post,
                        int use_tma, int min_hmax, int* out) {
    if (!k || !d1 || !d2 || !out || nrb < 1 || nrb > MRF3_MAX_RB) return 0;
    Mrf3Args m; memset(&m, 0, sizeof m);
    m.C = C; m.nrb = nrb; m.rb1 = rb1; m.out_div = (float)nrb; m.slope = 0.1f;
    for (int j = 0; j < nrb; j++) { m.k[j] = k[j]; m.d1[j] = d1[j]; m.d2[j] = d2[j]; }
    if (up_cin > 0) { m.up_u = 4; m.up_cin = up_cin; }
    Mrf3Cfg c;
    if (!mrf3_plan(m, c, nb_pref, fuse_post != 0, use_tma != 0, min_hmax)) return 0;
    const int v[16] = {c.nb, c.span, c.hmax, c.h1max, c.t_out, c.t_step, c.rx, c.rx1, c.smem_bytes, c.tmem_cols, c.nstages, c.resident,
                       c.tma, c.nboxes, c.box_rows, c.u_rows};
    memcpy(out, v, sizeof v);
    return 1;
}

int vits_test_conv_plan(int cin, int n, int ntaps, const int* toff, int xb, long xb_rows, int split3, int ntiles, int num_sms, int* out) {
    if (!toff || !out || ntaps < 1 || ntaps > CONV_MAX_TAPS) return 0;
    ConvArgs a; memset(&a, 0, sizeof a);
    a.cin = cin; a.n = n; a.npad = rup(n, 4); a.npad16 = rup(n, 16); a.ntaps = ntaps;
    for (int i = 0; i < ntaps; i++) a.toff[i] = toff[i];
    a.nks = 1; a.split3 = split3; a.ntiles = ntiles; a.ldo = a.npad16; a.ldx = cin; a.ldxb = cin;
    static const __nv_bfloat16 dummy[8] = {};
    if (xb) { a.xb = dummy; a.xb_rows = xb_rows; }
    TcCfg c; memset(&c, 0, sizeof c);
    if (!conv_tc_plan(a, c, num_sms)) return 0;
    const int v[16] = {c.ntile, c.rows_a, c.rows_need, c.a_bytes, c.slot_bytes, c.nstages, c.resident, c.nabuf, c.smem_bytes, c.tmem_cols,
                       c.tma, c.nboxes, c.box_rows, c.nepi, c.nload, c.naccbuf};
    memcpy(out, v, sizeof v);
    return 1;
}

int vits_test_file_arch(const char* path, vits_arch* arch, char* err, size_t err_cap) {
    if (!path || !arch) return VITS_E_INVALID;
    vf::Voice v;
    std::string e;
    if (!vf::load_voice(path, v, e, false)) { if (err && err_cap) snprintf(err, err_cap, "%s", e.c_str()); return VITS_E_INVALID; }
    *arch = v.arch;
    return VITS_OK;
}

int64_t vits_test_file_blob(const char* path, const char* name, void* out, int64_t cap, int* dtype) {
    static std::string cached_path;                      // the tests walk every blob of one file: parse and pack it once
    static vf::Voice cached;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!path) return VITS_E_INVALID;
    if (cached_path != path) {
        cached = vf::Voice();
        std::string e;
        cached_path.clear();
        if (!vf::load_voice(path, cached, e)) return VITS_E_INVALID;
        cached_path = path;
    }
    if (!name) return (int64_t)(cached.blobs.size() + cached.opts.size());
    if (name[0] == '#') {
        int64_t i = atoll(name + 1);
        std::string nm;
        if (i < (int64_t)cached.blobs.size()) { auto it = cached.blobs.begin(); std::advance(it, i); nm = it->first; }
        else if (i < (int64_t)(cached.blobs.size() + cached.opts.size())) { auto it = cached.opts.begin(); std::advance(it, i - (int64_t)cached.blobs.size()); nm = "opt:" + it->first; }
        else return VITS_E_INVALID;
        if (out && cap > (int64_t)nm.size()) memcpy(out, nm.c_str(), nm.size() + 1);
        return (int64_t)nm.size();
    }
    if (!strncmp(name, "opt:", 4)) {
        auto it = cached.opts.find(name + 4);
        if (it == cached.opts.end()) return VITS_E_INVALID;
        if (out && cap >= 8) memcpy(out, &it->second, 8);
        if (dtype) *dtype = 3;
        return 8;
    }
    auto it = cached.blobs.find(name);
    if (it == cached.blobs.end()) return VITS_E_INVALID;
    if (dtype) *dtype = it->second.dtype;
    if (out && cap >= (int64_t)it->second.bytes.size()) memcpy(out, it->second.bytes.data(), it->second.bytes.size());
    return (int64_t)it->second.bytes.size();
}

int vits_upload(vits_handle* h, const char* name, const void* data, size_t nbytes, int dtype) {
    if (!h || !name || !data) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    CK(h, cudaSetDevice(h->device));
    DevBlob b; b.bytes = nbytes; b.dtype = dtype;
    CK(h, cudaMalloc(&b.p, std::max<size_t>(nbytes, 16)));
    CK(h, cudaMemcpy(b.p, data, nbytes, cudaMemcpyHostToDevice));
    auto it = h->blobs.find(name);
    if (it != h->blobs.end()) { cudaFree(it->second.p); it->second = b; } else h->blobs[name] = b;
    h->finalized = false;
    return VITS_OK;
}

int vits_set_option(vits_handle* h, const char* key, double value) {
    if (!h || !key) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    std::string k(key);
    if (k == "precision") {
        if (value != 0 && value != 1) return fail(h, VITS_E_INVALID, "precision must be 0 (fp32) or 1 (bf16 tensor cores)");
        h->precision = (int)value;
    } else if (k == "text_tc") {
        h->text_tc = value != 0;
    } else if (k == "num_sms") {
        if (value < 1) return fail(h, VITS_E_INVALID, "num_sms must be >= 1");
        h->num_sms = (int)value;                 // test hook: persistent grids use this many CTAs
    } else if (k == "pdl") {
        g_pdl = value != 0;                      // programmatic dependent launch of every kernel (common.cuh launch_k; process-wide)
    } else if (k == "conv_tma") {
        g_tc_tma = value != 0;                   // experiment switch of conv_tc.cuh (process-wide): TMA or cp.async activation loader
    } else if (k == "conv_tma_max_cin") {
        g_tc_tma_max_cin = (int)value;
    } else if (k == "conv_nepi") {
        if (value != 8 && value != 12) return fail(h, VITS_E_INVALID, "conv_nepi must be 8 or 12");
        g_tc_nepi_xb = (int)value;               // experiment switch of conv_tc.cuh (process-wide)
    } else if (k == "max_chunk_frames") {
        if (value < 1) return fail(h, VITS_E_INVALID, "max_chunk_frames must be >= 1");
        h->max_chunk_frames = (int64_t)value;
    } else {
        h->opts[k] = value;
    }
    return VITS_OK;
}

int vits_finalize(vits_handle* h) {
    if (!h) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    const vits_arch& A = h->A;
    const int H = A.hidden, C = A.inter, F = A.filter, Fd = A.dp_filter;
    if (H % 32 || H > 256 || H % A.n_heads || (H / A.n_heads) % 4 || (H / A.n_heads) > 32 * ATT_MAX_DKM)
        return fail(h, VITS_E_INVALID, "hidden=%d / heads=%d unsupported", H, A.n_heads);
    if (C % 8) return fail(h, VITS_E_INVALID, "inter_channels=%d must be a multiple of 8", C);
    if (A.num_bins != SPL_K && A.use_sdp) return fail(h, VITS_E_INVALID, "num_bins=%d unsupported (kernel is specialised for %d)", A.num_bins, SPL_K);
    int rc;
    for (void* pz : h->owned) cudaFree(pz);
    h->owned.clear();
    if ((rc = need_f32(h, "enc.emb", &h->emb, (size_t)A.n_vocab * H))) return rc;
    h->enc.resize(A.n_layers);
    std::vector<int> ffn_taps(A.enc_kernel);
    for (int i = 0; i < A.enc_kernel; i++) ffn_taps[i] = i - (A.enc_kernel - 1) / 2;   // attentions.py:419-427
    const int nrel = 2 * A.window + 1, dk = H / A.n_heads;
    for (int i = 0; i < A.n_layers; i++) {
        auto& L = h->enc[i];
        std::string p = "enc." + std::to_string(i);
        if ((rc = mkconv(h, L.qkv, p + ".qkv", H, 3 * H, {0}, true))) return rc;
        if ((rc = need_f32(h, p + ".rel_k", &L.rel_k, (size_t)nrel * dk))) return rc;
        if ((rc = need_f32(h, p + ".rel_v", &L.rel_v, (size_t)nrel * dk))) return rc;
        if ((rc = mkconv(h, L.o, p + ".o", H, H, {0}, true))) return rc;
        if ((rc = mkln(h, L.ln1, p + ".ln1", H))) return rc;
        if ((rc = mkconv(h, L.ffn1, p + ".ffn1", H, F, ffn_taps, true))) return rc;
        if ((rc = mkconv(h, L.ffn2, p + ".ffn2", F, H, ffn_taps, true))) return rc;
        if ((rc = mkln(h, L.ln2, p + ".ln2", H))) return rc;
    }
    if ((rc = mkconv(h, h->enc_proj, "enc.proj", H, 2 * C, {0}, true))) return rc;
    if (A.n_speakers > 1) {
        if ((rc = need_f32(h, "dp.cond_tab", &h->dp_cond_tab, (size_t)A.n_speakers * (A.use_sdp ? Fd : H)))) return rc;
        if ((rc = need_f32(h, "dec.cond_tab", &h->dec_cond_tab, (size_t)A.n_speakers * A.up_init))) return rc;
    }
    if (A.use_sdp) {
        if ((rc = mkconv(h, h->dp_pre, "dp.pre", H, Fd, {0}, true))) return rc;
        if ((rc = mkconv(h, h->dp_proj, "dp.proj", Fd, Fd, {0}, true))) return rc;
        if ((rc = mkdds(h, h->dp_dds, "dp.convs", Fd))) return rc;
        h->cflows.resize(A.n_cflows);
        for (int k = 0; k < A.n_cflows; k++) {
            auto& cf = h->cflows[k];
            std::string p = "dp.flows." + std::to_string(A.cflows[k]);
            if ((rc = need_f32(h, p + ".pre_w", &cf.pre_w, Fd))) return rc;
            if ((rc = need_f32(h, p + ".pre_b", &cf.pre_b, Fd))) return rc;
            if ((rc = mkdds(h, cf.dds, p + ".convs", Fd))) return rc;
            if ((rc = mkconv(h, cf.proj, p + ".proj", Fd, rup(3 * A.num_bins - 1, 16), {0}, true))) return rc;   // packed with zero padding channels
        }
        if (!h->opts.count("dp.ea_m") || !h->opts.count("dp.ea_logs"))
            return fail(h, VITS_E_STATE, "missing options dp.ea_m / dp.ea_logs (ElementwiseAffine of the duration flow)");
        h->ea_m = (float)h->opts["dp.ea_m"]; h->ea_logs = (float)h->opts["dp.ea_logs"];
    } else {
        if ((rc = mkconv(h, h->dpd_c1, "dp.conv_1", H, Fd, sym_taps(A.dp_kernel, 1), true))) return rc;
        if ((rc = mkln(h, h->dpd_n1, "dp.norm_1", Fd))) return rc;
        if ((rc = mkconv(h, h->dpd_c2, "dp.conv_2", Fd, Fd, sym_taps(A.dp_kernel, 1), true))) return rc;
        if ((rc = mkln(h, h->dpd_n2, "dp.norm_2", Fd))) return rc;
        if ((rc = mkconv(h, h->dpd_proj, "dp.proj", Fd, 1, {0}, true))) return rc;
    }
    // coupling flow, execution order; Flip folded into column offsets + permuted weights
    h->flows.resize(A.n_flow);
    const int half = C / 2;
    if (half % 4) return fail(h, VITS_E_INVALID, "inter_channels/2 = %d must be a multiple of 4", half);
    for (int s = 0; s < A.n_flow; s++) {
        auto& f = h->flows[s];
        const bool flipped = (s % 2 == 0);
        f.xcol = flipped ? half : 0; f.ocol = flipped ? 0 : half;
        std::string p = "flow." + std::to_string(s);
        if ((rc = mkconv(h, f.pre, p + ".pre", half, H, {0}, true))) return rc;
        f.in.resize(A.wn_layers); f.rs.resize(A.wn_layers); f.cond_tab.assign(A.wn_layers, nullptr);
        int dil = 1;
        for (int i = 0; i < A.wn_layers; i++) {
            if ((rc = mkconv(h, f.in[i], p + ".in." + std::to_string(i), H, 2 * H, sym_taps(A.wn_kernel, dil), true))) return rc;
            if ((rc = mkconv(h, f.rs[i], p + ".rs." + std::to_string(i), H, (i < A.wn_layers - 1) ? 2 * H : H, {0}, true))) return rc;
            if (A.n_speakers > 1)
                if ((rc = need_f32(h, p + ".cond_tab." + std::to_string(i), &f.cond_tab[i], (size_t)A.n_speakers * 2 * H))) return rc;
            dil *= A.wn_dilation_rate;
        }
        if ((rc = mkconv(h, f.post, p + ".post", H, half, {0}, true))) return rc;
        // optional tensor-core tail: residual-only res_skip convs + one composed GEMM for m (packing.py)
        f.v2 = find_blob(h, p + ".mskip.b") != nullptr && H % 16 == 0 && half % 16 == 0;
        if (f.v2) {
            f.rsr.resize(A.wn_layers); f.mskip.resize(A.wn_layers);
            for (int i = 0; i < A.wn_layers && f.v2; i++) {
                if (i < A.wn_layers - 1 && (mkconv(h, f.rsr[i], p + ".rsr." + std::to_string(i), H, H, {0}, true) || !f.rsr[i].wtc)) f.v2 = false;
                if (f.v2 && (mkconv(h, f.mskip[i], p + ".mskip." + std::to_string(i), H, half, {0}, false) || !f.mskip[i].wtc)) f.v2 = false;
            }
            if (f.v2 && need_f32(h, p + ".mskip.b", &f.mskip_b, half)) f.v2 = false;
            if (A.wn_layers > CONV_MAX_SLICES) f.v2 = false;
        }
    }
    // decoder
    if ((rc = mkconv(h, h->dec_pre, "dec.pre", C, A.up_init, sym_taps(7, 1), true))) return rc;
    h->ups.resize(A.n_ups);
    int ch = A.up_init;
    h->rb_c1.clear(); h->rb_c2.clear();
    for (int i = 0; i < A.n_ups; i++) {
        const int u = A.up_rates[i];
        if (u % 2 || A.up_kernels[i] != 2 * u) return fail(h, VITS_E_INVALID, "upsample %d: rate %d kernel %d unsupported (need even rate, kernel = 2*rate)", i, u, A.up_kernels[i]);
        auto& U = h->ups[i];
        U.rate = u; U.cout = ch / 2;
        std::string p = "dec.ups." + std::to_string(i);
        if ((rc = mkconv(h, U.A, p + ".A", ch, (u / 2) * (ch / 2), {-1, 0}, true))) return rc;
        if ((rc = mkconv(h, U.B, p + ".B", ch, (u / 2) * (ch / 2), {0, 1}, true))) return rc;
        ch /= 2;
        for (int j = 0; j < A.n_rbk; j++) {
            const int n = i * A.n_rbk + j;
            std::vector<ConvP> c1(A.rb_ndil[j]), c2;
            if (A.resblock_type == 1) c2.resize(A.rb_ndil[j]);
            for (int c = 0; c < A.rb_ndil[j]; c++) {
                std::string q = "dec.rb." + std::to_string(n);
                if (A.resblock_type == 1) {
                    if ((rc = mkconv(h, c1[c], q + ".c1." + std::to_string(c), ch, ch, sym_taps(A.rb_kernels[j], A.rb_dilations[j][c]), true))) return rc;
                    if ((rc = mkconv(h, c2[c], q + ".c2." + std::to_string(c), ch, ch, sym_taps(A.rb_kernels[j], 1), true))) return rc;
                } else {
                    if ((rc = mkconv(h, c1[c], q + ".c." + std::to_string(c), ch, ch, sym_taps(A.rb_kernels[j], A.rb_dilations[j][c]), true))) return rc;
                }
            }
            h->rb_c1.push_back(c1); h->rb_c2.push_back(c2);
        }
    }
    // per stage: sum over the resblocks of the second conv's bias (ResBlock2): ONE launch sums the second convs of a stage
    for (float* pz : h->rb_b2sum) if (pz) cudaFree(pz);
    h->rb_b2sum.assign(A.n_ups, nullptr);
    if (A.resblock_type == 2) {
        int chs = A.up_init;
        for (int i = 0; i < A.n_ups; i++) {
            chs /= 2;
            bool ok = true;
            for (int j = 0; j < A.n_rbk; j++) if (A.rb_ndil[j] != 2 || !h->rb_c1[i * A.n_rbk + j][1].b) ok = false;
            if (!ok) continue;
            const int np4 = rup(chs, 4), np16 = rup(chs, 16);
            std::vector<float> sum(np16, 0.f), tmp(np4);
            for (int j = 0; j < A.n_rbk; j++) {
                CK(h, cudaMemcpy(tmp.data(), h->rb_c1[i * A.n_rbk + j][1].b, (size_t)np4 * 4, cudaMemcpyDeviceToHost));
                for (int q = 0; q < chs; q++) sum[q] += tmp[q];
            }
            CK(h, cudaMalloc(&h->rb_b2sum[i], (size_t)np16 * 4));
            CK(h, cudaMemcpy(h->rb_b2sum[i], sum.data(), (size_t)np16 * 4, cudaMemcpyHostToDevice));
        }
    }
    h->post_c = ch;
    if (ch > 64) return fail(h, VITS_E_INVALID, "conv_post input width %d > 64 unsupported", ch);
    if ((rc = need_f32(h, "dec.post_w", &h->post_w, (size_t)7 * ch))) return rc;
    h->finalized = true;
    return VITS_OK;
}

int vits_prepare(vits_handle* h, const int64_t* ids, const int64_t* lengths, int32_t B, const float scales[3],
                 const int64_t* sid, const float* noise_dp, int64_t dp_stride, const float* logw_override,
                 uint64_t seed, int64_t* y_lengths, int64_t* total_frames) {
    if (!h) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->finalized) return fail(h, VITS_E_STATE, "vits_finalize() has not succeeded");
    if (!ids || !lengths || !scales || B <= 0 || !y_lengths) return fail(h, VITS_E_INVALID, "null argument or B <= 0");
    const vits_arch& A = h->A;
    h->prepared = false;
    // ---- validation (what ORT's Gather would raise on)
    long R = 0;
    h->h_cu_t.assign(B + 1, 0);
    for (int b = 0; b < B; b++) {
        if (lengths[b] < 1 || lengths[b] > (1 << 20)) return fail(h, VITS_E_INVALID, "input_lengths[%d] = %lld out of range", b, (long long)lengths[b]);
        R += lengths[b]; h->h_cu_t[b + 1] = (int)R;
        if (R > (1l << 30)) return fail(h, VITS_E_INVALID, "batch too large");
    }
    std::vector<int> h_ids(R);
    for (long i = 0; i < R; i++) {
        if (ids[i] < 0 || ids[i] >= A.n_vocab) return fail(h, VITS_E_INVALID, "phoneme id %lld at position %ld outside [0, %d)", (long long)ids[i], i, A.n_vocab);
        h_ids[i] = (int)ids[i];
    }
    h->h_sid.assign(B, 0);
    if (A.n_speakers > 1) {
        if (!sid) return fail(h, VITS_E_INVALID, "multi-speaker voice: 'sid' is required");
        for (int b = 0; b < B; b++) {
            if (sid[b] < 0 || sid[b] >= A.n_speakers) return fail(h, VITS_E_INVALID, "sid[%d] = %lld outside [0, %d)", b, (long long)sid[b], A.n_speakers);
            h->h_sid[b] = (int)sid[b];
        }
    }
    if (!(scales[1] > 0.f) || scales[0] < 0.f || scales[2] < 0.f) return fail(h, VITS_E_INVALID, "scales must be >= 0 (length_scale > 0)");
    if (noise_dp) for (int b = 0; b < B; b++) if (dp_stride < lengths[b]) return fail(h, VITS_E_INVALID, "noise_dp stride %lld < length %lld", (long long)dp_stride, (long long)lengths[b]);
    h->scales[0] = scales[0]; h->scales[1] = scales[1]; h->scales[2] = scales[2];
    h->seed = seed; h->B = B; h->R = (int)R;
    h->utt_base = h->utt_counter; h->utt_counter += (uint64_t)B;
    h->conv_text_counter = 0;
    CK(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int H = A.hidden, C = A.inter, F = A.filter, Fd = A.dp_filter;
    int rc;
    // ---- buffers
    TileBuilder tb; tb.begin(h->h_cu_t.data(), B); tb.add(1);
    if ((rc = ensure(h, h->ids, R * 4)) || (rc = ensure(h, h->tile_t, tb.host.size() * 4)) || (rc = ensure(h, h->sid, B * 4)) ||
        (rc = ensure(h, h->x, R * H * 4)) || (rc = ensure(h, h->y, R * H * 4)) || (rc = ensure(h, h->qkv, R * 3 * H * 4)) ||
        (rc = ensure(h, h->att, R * H * 4)) || (rc = ensure(h, h->ffn, R * F * 4)) || (rc = ensure(h, h->stats, R * 2 * C * 4)) ||
        (rc = ensure(h, h->d0, R * Fd * 4)) || (rc = ensure(h, h->d1, R * Fd * 4)) || (rc = ensure(h, h->y, R * std::max(H, Fd) * 4)) ||
        (rc = ensure(h, h->gdp, R * Fd * 4)) || (rc = ensure(h, h->hp, R * 32 * 4)) || (rc = ensure(h, h->z0, R * 4)) ||
        (rc = ensure(h, h->z1, R * 4)) || (rc = ensure(h, h->logw, R * 4)) || (rc = ensure(h, h->dur, R * 4)) ||
        (rc = ensure(h, h->cum, R * 4)) || (rc = ensure(h, h->ylen, B * 4)))
        return rc;
    CK(h, cudaMemcpyAsync(h->ids.p, h_ids.data(), R * 4, cudaMemcpyHostToDevice, st));
    CK(h, cudaMemcpyAsync(h->tile_t.p, tb.host.data(), tb.host.size() * 4, cudaMemcpyHostToDevice, st));
    CK(h, cudaMemcpyAsync(h->sid.p, h->h_sid.data(), B * 4, cudaMemcpyHostToDevice, st));
    const float* d_inj = nullptr;
    if (noise_dp && A.use_sdp && !logw_override) {
        size_t nb = (size_t)B * 2 * dp_stride * 4;
        if ((rc = ensure(h, h->inj_dp, nb))) return rc;
        CK(h, cudaMemcpyAsync(h->inj_dp.p, noise_dp, nb, cudaMemcpyHostToDevice, st));
        d_inj = ptr<float>(h->inj_dp);
    }
    Tiles T = tb.get(ptr<int>(h->tile_t), 1);
    if ((rc = ensure(h, h->tdesc_t, (size_t)std::max(T.n128 + T.n64, 1) * sizeof(int4))) || (rc = make_d128(h, T, ptr<int4>(h->tdesc_t)))) return rc;
    if (T.n64 > 0) {     // 64-row tile descriptors (attention q tiles) behind the 128-row ones
        launch_k(k_tile_desc, (T.n64 + 255) / 256, 256, 0, st, T.cu, T.t64, T.B, 1, T.n64, 64, ptr<int4>(h->tdesc_t) + T.n128);
        h->launches++;
    }
    if ((rc = ensure(h, h->rowpos, (size_t)R * sizeof(int2)))) return rc;
    launch_k(k_row_pos, (unsigned)((R + 255) / 256), 256, 0, st, T.cu, B, (int)R, ptr<int2>(h->rowpos));
    h->launches++;
    const int* d_sid = ptr<int>(h->sid);
    float *x = ptr<float>(h->x), *y = ptr<float>(h->y), *qkv = ptr<float>(h->qkv), *att = ptr<float>(h->att),
          *ffn = ptr<float>(h->ffn), *stats = ptr<float>(h->stats);
    stage_begin(h, 0);
    // ---- text encoder (models.py:198-209, attentions.py:60-74)
    {
        long n4 = R * (H / 4);
        launch_k(k_embed, (unsigned)((n4 + 255) / 256), 256, 0, st, ptr<int>(h->ids), h->emb, x, (int)R, H, sqrtf((float)H));
        h->launches++;
        const int dk = H / A.n_heads, nrel = 2 * A.window + 1;
        for (int i = 0; i < A.n_layers; i++) {
            auto& L = h->enc[i];
            ConvArgs a = base_args(L.qkv, x, H, 0, qkv, 3 * H, 0);
            if ((rc = launch_conv_text(h, L.qkv, a, T))) return rc;
            cudaError_t ae = cudaSuccess;
            // bf16 mode: both contractions on the tensor cores as fp32-faithful bf16x3 products; fp32 mode: CUDA cores
            if (h->precision == 1 && h->text_tc && h->opts["attention_v1"] == 0 && h->opts["attention_fp32"] == 0 &&
                attention_mma_launch(qkv, L.rel_k, L.rel_v, att, ptr<int4>(h->tdesc_t) + T.n128, T.n64, H, A.n_heads, dk, A.window, st, &ae)) {
                if (ae != cudaSuccess) return fail(h, VITS_E_CUDA, "attention (mma) launch: %s", cudaGetErrorString(ae));
            } else if (h->opts["attention_v1"] == 0 &&
                attention_tiled_launch(qkv, L.rel_k, L.rel_v, att, T.cu, T.t64, T.n64, B, H, A.n_heads, dk, A.window, st, &ae)) {
                if (ae != cudaSuccess) return fail(h, VITS_E_CUDA, "attention launch: %s", cudaGetErrorString(ae));
            } else {
                launch_k(k_rel_attention, dim3((R + 3) / 4, A.n_heads), 128, 4 * (dk + nrel) * sizeof(float), st, 
                    qkv, L.rel_k, L.rel_v, att, T.cu, B, (int)R, H, A.n_heads, dk, A.window);
            }
            h->launches++;
            a = base_args(L.o, att, H, 0, y, H, 0); a.res = x; a.ldres = H;
            if ((rc = launch_conv_text(h, L.o, a, T))) return rc;
            if ((rc = launch_ln(h, y, x, L.ln1, (int)R, H, 0, nullptr, T))) return rc;
            a = base_args(L.ffn1, x, H, 0, ffn, F, 0); a.out_act = ACT_RELU;
            if ((rc = launch_conv_text(h, L.ffn1, a, T))) return rc;
            a = base_args(L.ffn2, ffn, F, 0, y, H, 0); a.res = x; a.ldres = H;
            if ((rc = launch_conv_text(h, L.ffn2, a, T))) return rc;
            if ((rc = launch_ln(h, y, x, L.ln2, (int)R, H, 0, nullptr, T))) return rc;
        }
        ConvArgs a = base_args(h->enc_proj, x, H, 0, stats, 2 * C, 0);
        if ((rc = launch_conv_text(h, h->enc_proj, a, T))) return rc;
    }
    // ---- duration predictor
    float* logw = ptr<float>(h->logw);
    if (logw_override) {
        CK(h, cudaMemcpyAsync(logw, logw_override, R * 4, cudaMemcpyHostToDevice, st));
    } else if (A.use_sdp) {
        float *d0 = ptr<float>(h->d0), *d1 = ptr<float>(h->d1), *gdp = ptr<float>(h->gdp), *hp = ptr<float>(h->hp);
        float *z0 = ptr<float>(h->z0), *z1 = ptr<float>(h->z1);
        ConvArgs a = base_args(h->dp_pre, x, H, 0, d0, Fd, 0);
        if (A.n_speakers > 1) { a.utab = h->dp_cond_tab; a.uidx = d_sid; a.utab_ld = Fd; }
        if ((rc = launch_conv_text(h, h->dp_pre, a, T))) return rc;
        if ((rc = run_dds(h, h->dp_dds, d0, d1, y, (int)R, Fd, T))) return rc;
        a = base_args(h->dp_proj, d0, Fd, 0, gdp, Fd, 0);
        if ((rc = launch_conv_text(h, h->dp_proj, a, T))) return rc;
        launch_k(k_noise_dp, (R + 255) / 256, 256, 0, st, z0, z1, d_inj, dp_stride, T.cu, B, (int)R, h->scales[2], seed, h->utt_base);
        h->launches++;
        if (h->opts.count("debug_keep_noise_dp") && h->opts["debug_keep_noise_dp"] != 0) {
            // test hook: the duration predictor's noise (models.py:111) as drawn, [z0 | z1], before the flows overwrite it
            if ((rc = ensure(h, h->dbg_dp, (size_t)R * 2 * 4))) return rc;
            CK(h, cudaMemcpyAsync(h->dbg_dp.p, z0, (size_t)R * 4, cudaMemcpyDeviceToDevice, st));
            CK(h, cudaMemcpyAsync(ptr<float>(h->dbg_dp) + R, z1, (size_t)R * 4, cudaMemcpyDeviceToDevice, st));
        }
        for (int k = 0; k < A.n_cflows; k++) {
            std::swap(z0, z1);                                   // Flip (modules.py:386)
            auto& cf = h->cflows[k];
            long n = R * Fd;
            launch_k(k_cf_pre, (unsigned)((n + 255) / 256), 256, 0, st, z0, cf.pre_w, cf.pre_b, gdp, d0, (int)R, Fd);
            h->launches++;
            if ((rc = run_dds(h, cf.dds, d0, d1, y, (int)R, Fd, T))) return rc;
            a = base_args(cf.proj, d0, Fd, 0, hp, 32, 0);
            if ((rc = launch_conv_text(h, cf.proj, a, T))) return rc;
            launch_k(k_spline_inverse, (R + 127) / 128, 128, 0, st, hp, 32, z1, (int)R, 1.f / sqrtf((float)Fd));
            h->launches++;
        }
        std::swap(z0, z1);                                       // final Flip, then EA on channel 0
        launch_k(k_ea_logw, (R + 255) / 256, 256, 0, st, z0, h->ea_m, expf(-h->ea_logs), logw, (int)R);
        h->launches++;
    } else {
        float *d0 = ptr<float>(h->d0), *d1 = ptr<float>(h->d1);
        const float* xin = x;
        if (A.n_speakers > 1) {
            CK(h, cudaMemcpyAsync(y, x, R * H * 4, cudaMemcpyDeviceToDevice, st));
            long n = R * H;
            launch_k(k_add_rowbias, (unsigned)((n + 255) / 256), 256, 0, st, y, h->dp_cond_tab, d_sid, T.cu, B, (int)R, H);
            h->launches++;
            xin = y;
        }
        ConvArgs a = base_args(h->dpd_c1, xin, H, 0, d0, Fd, 0); a.out_act = ACT_RELU;
        if ((rc = launch_conv_text(h, h->dpd_c1, a, T))) return rc;
        if ((rc = launch_ln(h, d0, d0, h->dpd_n1, (int)R, Fd, 0, nullptr, T))) return rc;
        a = base_args(h->dpd_c2, d0, Fd, 0, d1, Fd, 0); a.out_act = ACT_RELU;
        if ((rc = launch_conv_text(h, h->dpd_c2, a, T))) return rc;
        if ((rc = launch_ln(h, d1, d1, h->dpd_n2, (int)R, Fd, 0, nullptr, T))) return rc;
        a = base_args(h->dpd_proj, d1, Fd, 0, logw, 1, 0);
        if ((rc = launch_conv(h, a, T, false))) return rc;
    }
    // ---- length regulation (integer path)
    launch_k(k_durations, B, 256, 0, st, logw, h->scales[1], T.cu, ptr<int>(h->dur), ptr<int>(h->cum), ptr<int>(h->ylen));
    h->launches++;
    CK(h, cudaGetLastError());
    stage_end(h);
    h->h_ylen.assign(B, 0);
    CK(h, cudaMemcpyAsync(h->h_ylen.data(), h->ylen.p, B * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));          // the one host sync of the path: frame counts
    h->h_cu_y.assign(B + 1, 0);
    int64_t tot = 0;
    for (int b = 0; b < B; b++) {
        // k_durations saturates an utterance's frame count at INT_MAX instead of wrapping
        if (h->h_ylen[b] < 1 || h->h_ylen[b] > (1 << 30)) return fail(h, VITS_E_INVALID, "durations overflow: utterance %d has more than 2^30 frames", b);
        tot += h->h_ylen[b];
        if (tot > (1ll << 30)) return fail(h, VITS_E_INVALID, "durations overflow: more than 2^30 frames");
        h->h_cu_y[b + 1] = (int)tot;
        y_lengths[b] = h->h_ylen[b];
    }
    h->total_frames = tot;
    if (total_frames) *total_frames = tot;
    h->prepared = true;
    return VITS_OK;
}

int vits_decode(vits_handle* h, const float* noise_z, int64_t z_stride, int32_t out_kind, void* out,
                int64_t out_capacity, float volume, int32_t normalize) {
    if (!h) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->prepared) return fail(h, VITS_E_STATE, "vits_decode() before a successful vits_prepare()");
    const vits_arch& A = h->A;
    const int B = h->B, H = A.hidden, C = A.inter;
    h->conv_counter = 0;
    int hop = 1; for (int i = 0; i < A.n_ups; i++) hop *= A.up_rates[i];
    const int64_t total_samples = h->total_frames * hop;
    const bool async_flag = (out_kind & VITS_OUT_ASYNC) != 0;
    out_kind &= ~VITS_OUT_ASYNC;
    if (out_kind < 0 || out_kind > 3) return fail(h, VITS_E_INVALID, "out_kind %d", out_kind);
    std::vector<int64_t> offs;                           // consumed by this call whatever happens next
    offs.swap(h->out_offsets);
    const bool scatter = !offs.empty();
    if (scatter) {
        if (out_kind != 1 && out_kind != 2) return fail(h, VITS_E_INVALID, "output offsets apply to host outputs (out_kind 1 or 2)");
        if ((int)offs.size() != B) return fail(h, VITS_E_INVALID, "output offsets given for %d utterances, the prepared batch has %d", (int)offs.size(), B);
        if (!out) return fail(h, VITS_E_INVALID, "null output buffer");
        for (int b = 0; b < B; b++)
            if (offs[b] < 0 || offs[b] + (int64_t)h->h_ylen[b] * hop > out_capacity)
                return fail(h, VITS_E_INVALID, "utterance %d: offset %lld + %lld samples exceeds the output capacity %lld", b, (long long)offs[b],
                            (long long)h->h_ylen[b] * hop, (long long)out_capacity);
    } else
    if (out_kind != 0 && (!out || out_capacity < total_samples)) return fail(h, VITS_E_INVALID, "output buffer too small: %lld < %lld samples", (long long)out_capacity, (long long)total_samples);
    if (noise_z) for (int b = 0; b < B; b++) if (z_stride < h->h_ylen[b]) return fail(h, VITS_E_INVALID, "noise_z stride %lld < frames %d of utterance %d", (long long)z_stride, h->h_ylen[b], b);
    CK(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    int rc;
    const float* d_injz = nullptr;
    if (noise_z) {
        size_t nb = (size_t)B * C * z_stride * 4;
        if ((rc = ensure(h, h->inj_z, nb))) return rc;
        CK(h, cudaMemcpyAsync(h->inj_z.p, noise_z, nb, cudaMemcpyHostToDevice, st));
        d_injz = ptr<float>(h->inj_z);
    }
    // alternate between two device audio buffers; the one chosen must have finished leaving for the host
    h->audio_sel ^= 1;
    const int asel = h->audio_sel;
    if (h->out_pending[asel]) { CK(h, cudaEventSynchronize(h->ev_out[asel])); h->out_pending[asel] = false; }
    Buf& abuf = asel ? h->audio_alt : h->audio;
    if ((rc = ensure(h, abuf, std::max<int64_t>(total_samples, 1) * 4))) return rc;
    float* audio = ptr<float>(abuf);
    const bool async_out = async_flag || (h->opts.count("async_output") && h->opts["async_output"] != 0);
    // int16 results (out_kind 2) leave per chunk like fp32 ones: chunks are whole utterances, so the per-utterance peak is known
    // when the chunk's kernels finish; two device buffers alternate like the fp32 ones
    int16_t* audio16 = nullptr;
    if (out_kind == 2) {
        Buf& a16 = asel ? h->audio16_alt : h->audio16;
        if ((rc = ensure(h, a16, std::max<int64_t>(total_samples, 1) * 2 + 16)) || (rc = ensure(h, h->peaks, (size_t)B * 4)) ||
            (rc = ensure(h, h->cu_y_dev, (size_t)(B + 1) * 4)))
            return rc;
        audio16 = ptr<int16_t>(a16);
        CK(h, cudaMemcpyAsync(h->cu_y_dev.p, h->h_cu_y.data(), (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
        CK(h, cudaMemsetAsync(h->peaks.p, 0, (size_t)B * 4, st));
    }
    h->last_nchunks = 0;
    // per-stage geometry
    std::vector<int> rates(A.n_ups + 1), chans(A.n_ups + 1);
    rates[0] = 1; chans[0] = A.up_init;
    for (int i = 0; i < A.n_ups; i++) { rates[i + 1] = rates[i] * A.up_rates[i]; chans[i + 1] = chans[i] / 2; }
    size_t stage_elems_per_frame = 0;
    for (int i = 1; i <= A.n_ups; i++) stage_elems_per_frame = std::max(stage_elems_per_frame, (size_t)rates[i] * chans[i]);

    int b_lo = 0;
    while (b_lo < B) {
        int b_hi = b_lo + 1;
        int64_t fr = h->h_ylen[b_lo];
        while (b_hi < B && fr + h->h_ylen[b_hi] <= h->max_chunk_frames) { fr += h->h_ylen[b_hi]; b_hi++; }
        const int nB = b_hi - b_lo;
        const int Fr = (int)fr;
        const int f_lo = h->h_cu_y[b_lo];
        h->last_nchunks++;
        std::vector<int> cu_local(nB + 1);
        for (int i = 0; i <= nB; i++) cu_local[i] = h->h_cu_y[b_lo + i] - f_lo;
        // fused MRF stage kernel (mrf3_tc.cuh) where the stage qualifies: bf16 mode, ResBlock2, 32 / 64 channels.  The last stage also
        // absorbs its ConvTranspose1d and lrelu -> conv_post -> tanh.  Stages that do not qualify run conv by conv on bf16 operand rows
        // (below) -- which is also the reference the fused kernel is tested against (option no_fused_mrf).
        // input tiles of the fused kernels by TMA (cp.async.bulk.tensor) unless option mrf_tma = 0 or the driver has no encoder
        const bool use_tma = (h->opts.count("mrf_tma") ? h->opts["mrf_tma"] != 0 : true) && tmap_encoder() != nullptr;
        std::vector<Mrf3Args> mrf3_args(A.n_ups + 1);
        std::vector<Mrf3Cfg> mrf3_cfg(A.n_ups + 1);
        std::vector<int> mrf_on(A.n_ups + 1, 0);      // 0: conv by conv, 3: fused (bf16 inter-stage rows, optional fused ConvTranspose)
        std::vector<int> up_fused(A.n_ups + 2, 0);    // stage's ConvTranspose runs inside its fused kernel
        bool post_fused = false;
        for (int i = 0; i < A.n_ups; i++) {
            const int co = chans[i + 1];
            if (h->precision != 1 || A.resblock_type != 2 || A.n_rbk > MRF_MAX_RB || h->opts["no_fused_mrf"] != 0) continue;
            // the stage's input arrives as bf16 lrelu rows; when the previous stage is fused as well and the geometry allows (u = 4),
            // the ConvTranspose runs inside the kernel on the previous stage's bf16 output
            Mrf3Args& m3 = mrf3_args[i + 1];
            memset(&m3, 0, sizeof m3);
            m3.C = co; m3.nrb = A.n_rbk; m3.out_div = (float)A.n_rbk; m3.slope = 0.1f;
            m3.interleave = h->opts.count("mrf_interleave") ? (int)h->opts["mrf_interleave"] : 0;
            bool ok = true;
            for (int j = 0; j < A.n_rbk; j++) {
                const auto& cv = h->rb_c1[i * A.n_rbk + j];
                if (A.rb_ndil[j] != 2 || !cv[0].wtc || !cv[1].wtc) { ok = false; break; }
                m3.k[j] = A.rb_kernels[j]; m3.d1[j] = A.rb_dilations[j][0]; m3.d2[j] = A.rb_dilations[j][1];
                m3.w[j][0] = cv[0].wtc; m3.w[j][1] = cv[1].wtc; m3.b[j][0] = cv[0].b; m3.b[j][1] = cv[1].b;
            }
            if (!ok) continue;
            const bool want_post = (i == A.n_ups - 1) && h->opts["no_fused_post"] == 0 && co <= 64;
            // TMA input tiles only when the launch has more tiles than half the SMs (latency mode keeps the cp.async loader: conv_tc.cuh)
            const bool many_tiles = (long)Fr * rates[i + 1] / 512 * 2 > (long)h->num_sms;
            const int nbp = (int)(h->opts.count("mrf_nb") ? h->opts["mrf_nb"] : (co == 32 ? 4 : 2));
            const auto& U = h->ups[i];
            const bool can_up = i >= 1 && mrf_on[i] == 3 && U.rate == 4 && U.A.wtc && U.B.wtc && h->opts["no_fused_ups"] == 0;
            bool done = false;
            for (int tryu = can_up ? 1 : 0; tryu >= 0 && !done; tryu--) {
                m3.up_u = tryu ? U.rate : 0; m3.up_cin = tryu ? U.A.cin : 0;
                for (int tryp = want_post ? 1 : 0; tryp >= 0 && !done; tryp--)
                    if (mrf3_plan(m3, mrf3_cfg[i + 1], nbp, tryp != 0, use_tma && many_tiles)) {
                        mrf_on[i + 1] = 3; up_fused[i + 1] = tryu; if (tryp) post_fused = true; done = true;
                    }
            }
            if (!done) { m3.up_u = 0; m3.up_cin = 0; }
        }
        // unfused ResBlock2 stages (e.g. the 128-channel first stage) in bf16 mode: bf16 operand rows between the convs
        std::vector<int> rb_bf16(A.n_ups + 2, 0);
        for (int i = 0; i < A.n_ups; i++) {
            const int co = chans[i + 1];
            // (ResBlock1 stages too, round 2: `high` ran conv by conv on fp32 rows -- 12 B per value per conv through HBM and an fp32 ->
            // bf16 conversion pass in every loader; C4 decoder 0.25 of tensor peak, profiles/r02e)
            bool ok = h->precision == 1 && mrf_on[i + 1] == 0 && h->opts["no_stage_bf16"] == 0 && co % 16 == 0 &&
                      h->ups[i].A.wtc && h->ups[i].B.wtc && (h->ups[i].rate * co) % 8 == 0;
            for (int j = 0; j < A.n_rbk && ok; j++)
                for (int c2 = 0; c2 < A.rb_ndil[j]; c2++) {
                    if (!h->rb_c1[i * A.n_rbk + j][c2].wtc) ok = false;
                    if (A.resblock_type == 1 && !h->rb_c2[i * A.n_rbk + j][c2].wtc) ok = false;
                }
            rb_bf16[i + 1] = ok;
        }
        // ResBlock1 stages of 32 / 64 channels (`high`): every (conv_{k,d} -> conv_{k,1}) pair of modules.py:301-314 as ONE fused launch
        // (mrf3_tc.cuh with n_r = 1, rb1): the intermediate never leaves shared memory -- conv by conv these stages move ~1.3 GB per
        // launch and sit at 50-67 % of HBM bandwidth (profiles/r02l_launch_list_C4.txt).  All pairs of a stage share one tile table
        // (the halo of the widest second conv).
        std::vector<int> rb1_fused(A.n_ups + 2, 0), rb1_post(A.n_ups + 2, 0);
        std::vector<std::vector<Mrf3Args>> rb1_args(A.n_ups + 2);
        std::vector<std::vector<Mrf3Cfg>> rb1_cfg(A.n_ups + 2);
        for (int i = 0; i < A.n_ups; i++) {
            const int co = chans[i + 1];
            if (A.resblock_type != 1 || !rb_bf16[i + 1] || (co != 32 && co != 64 && co != 128) || h->opts["no_fused_rb1"] != 0) continue;
            if (co == 128 && h->opts["no_fused_rb1_128"] != 0) continue;
            int hmax = 0;
            bool shape_ok = true;
            for (int j = 0; j < A.n_rbk; j++) { hmax = std::max(hmax, (A.rb_kernels[j] - 1) / 2); if (A.rb_kernels[j] % 2 == 0) shape_ok = false; }
            if (!shape_ok) continue;
            const bool many_tiles = (long)Fr * rates[i + 1] / 512 * 2 > (long)h->num_sms;
            // last stage: every pair is planned with conv_post's geometry (tiles 6 rows shorter) so that the LAST pair of the last
            // resblock can run lrelu -> conv_post -> tanh at its tail, on the stage output it has just combined (no fp32 stage output,
            // no separate conv_post pass)
            // OPT-IN (option rb1_fused_post): conv_post then reads a bf16 operand instead of the fp32 stage output -- +6.5 % on C4 (decoder
            // 0.46 -> 0.50 of tensor peak) for 7 dB of the preset's margin against the oracle (worst utterance of the bench's spot check
            // 51.3 -> 44.5 dB, gate 40 dB; profiles/r02zl...): not the default.
            const bool want_post = (i == A.n_ups - 1) && co <= 64 && A.n_rbk >= 2 && A.n_rbk <= 3 && h->opts["no_fused_post"] == 0 &&
                                   h->opts["no_rb1_rows_out"] == 0 && h->opts["rb1_fused_post"] != 0;
            for (int tryp = want_post ? 1 : 0; tryp >= 0 && !rb1_fused[i + 1]; tryp--)
            for (int nbp = (co == 32 ? 4 : (co == 64 ? 2 : 1)); nbp >= 1 && !rb1_fused[i + 1]; nbp--) {
                std::vector<Mrf3Args> as; std::vector<Mrf3Cfg> cs;
                bool all = true;
                for (int j = 0; j < A.n_rbk && all; j++)
                    for (int c2 = 0; c2 < A.rb_ndil[j] && all; c2++) {
                        const ConvP &c1v = h->rb_c1[i * A.n_rbk + j][c2], &c2v = h->rb_c2[i * A.n_rbk + j][c2];
                        Mrf3Args m; memset(&m, 0, sizeof m);
                        m.C = co; m.nrb = 1; m.rb1 = 1; m.out_div = 1.f; m.slope = 0.1f;
                        m.k[0] = A.rb_kernels[j]; m.d1[0] = A.rb_dilations[j][c2]; m.d2[0] = 1;
                        m.w[0][0] = c1v.wtc; m.w[0][1] = c2v.wtc; m.b[0][0] = c1v.b; m.b[0][1] = c2v.b;
                        Mrf3Cfg cf;
                        if (!c1v.wtc || !c2v.wtc || !c1v.b || !c2v.b || c1v.ntaps != m.k[0] || c2v.ntaps != m.k[0] ||
                            !mrf3_plan(m, cf, nbp, tryp != 0, use_tma && many_tiles, hmax) || cf.nb != nbp) { all = false; break; }
                        as.push_back(m); cs.push_back(cf);
                    }
                if (all && !as.empty()) { rb1_fused[i + 1] = 1; rb1_post[i + 1] = tryp; rb1_args[i + 1] = as; rb1_cfg[i + 1] = cs; if (tryp) post_fused = true; }
            }
        }
        // ... of which: stages whose second convs run as one summed launch (needs a consumer that takes bf16 rows or fp32)
        std::vector<int> stage_sum2(A.n_ups + 2, 0);
        for (int i = 0; i < A.n_ups; i++) {
            int tt = 0; bool ok = rb_bf16[i + 1] && h->opts["no_stage_sum2"] == 0 && A.n_rbk <= 3 && A.n_rbk <= CONV_MAX_SLICES && h->rb_b2sum[i] != nullptr;
            for (int j = 0; j < A.n_rbk && ok; j++) { if (A.rb_ndil[j] != 2) ok = false; tt += A.rb_kernels[j]; }
            // the consumer of bf16 rows is the next stage's ConvTranspose launch; a fused-ups v3 kernel also reads them; the last stage stays fp32
            stage_sum2[i + 1] = ok && tt <= CONV_MAX_TAPS;
        }
        TileBuilder tb; tb.begin(cu_local.data(), nB);
        for (int i = 0; i <= A.n_ups; i++)
            tb.add(rates[i], mrf_on[i] == 3 ? mrf3_cfg[i].t_step : (rb1_fused[i] ? rb1_cfg[i][0].t_step : 0));
        if ((rc = ensure(h, h->chunk_meta, tb.host.size() * 4)) || (rc = ensure(h, h->P, (size_t)Fr * C * 4)) ||
            (rc = ensure(h, h->fh, (size_t)Fr * H * 4)) || (rc = ensure(h, h->facts, (size_t)Fr * H * 4)) ||
            (rc = ensure(h, h->fskip, (size_t)Fr * H * 4)) || (rc = ensure(h, h->fidx, (size_t)Fr * 4)) || (rc = ensure(h, h->fpos, (size_t)Fr * 8)) ||
            (rc = ensure(h, h->dpre, (size_t)Fr * A.up_init * 4)) ||
            (rc = ensure(h, h->sX, Fr * stage_elems_per_frame * 4)) || (rc = ensure(h, h->sT1, Fr * stage_elems_per_frame * 4)) ||
            (rc = ensure(h, h->sXSa, Fr * stage_elems_per_frame * 4)) || (rc = ensure(h, h->sXSb, Fr * stage_elems_per_frame * 4)) ||
            (rc = ensure(h, h->sYa, Fr * stage_elems_per_frame * 4)))
            return rc;
        if (A.resblock_type == 1)
            if ((rc = ensure(h, h->sYb, Fr * stage_elems_per_frame * 4))) return rc;
        CK(h, cudaMemcpyAsync(h->chunk_meta.p, tb.host.data(), tb.host.size() * 4, cudaMemcpyHostToDevice, st));
        const int* meta = ptr<int>(h->chunk_meta);
        // per-rate tile descriptors for the tcgen05 conv kernel
        std::vector<Tiles> TR(A.n_ups + 1);
        {
            size_t ntl = 0;
            for (int i = 0; i <= A.n_ups; i++) { TR[i] = tb.get(meta, rates[i]); ntl += (size_t)TR[i].n128; }
            if ((rc = ensure(h, h->tdesc_c, std::max<size_t>(ntl, 1) * sizeof(int4)))) return rc;
            int4* dp = ptr<int4>(h->tdesc_c);
            for (int i = 0; i <= A.n_ups; i++) { if ((rc = make_d128(h, TR[i], dp))) return rc; dp += TR[i].n128; }
        }
        const Tiles T1 = TR[0];
        const int* d_sid = ptr<int>(h->sid) + b_lo;
        float *P = ptr<float>(h->P), *fh = ptr<float>(h->fh), *facts = ptr<float>(h->facts), *fskip = ptr<float>(h->fskip);
        // ---- prior expansion + sampling (models.py:705-718)
        stage_begin(h, 1);
        launch_k(k_frame_index, (Fr + 255) / 256, 256, 0, st, ptr<int>(h->cum), ptr<int>(h->tile_t), T1.cu, b_lo, nB, Fr, ptr<int>(h->fidx), ptr<int2>(h->fpos));
        h->launches++;
        {
            long n = (long)Fr * (C / 4);
            // test hook "debug_eps": m_p = logs_p = 0, so z_p IS the noise draw (times noise_scale) -- the statistical test of the device RNG
            const float* stats_in = (h->opts.count("debug_eps") && h->opts["debug_eps"] != 0) ? nullptr : ptr<float>(h->stats);
            launch_k(k_expand_sample, (unsigned)((n + 255) / 256), 256, 0, st, stats_in, ptr<int>(h->fidx), ptr<int2>(h->fpos),
                                                                         d_injz, z_stride, h->scales[0], h->seed, h->utt_base, P, Fr, C);
            h->launches++;
        }
        h->last_chunk_frames = Fr;
        if (h->opts.count("debug_keep_zp") && h->opts["debug_keep_zp"] != 0) {
            if ((rc = ensure(h, h->dbg_zp, (size_t)Fr * C * 4))) return rc;
            CK(h, cudaMemcpyAsync(h->dbg_zp.p, P, (size_t)Fr * C * 4, cudaMemcpyDeviceToDevice, st));
        }
        // ---- coupling flow, reverse (models.py:247-254, modules.py:447-466, 184-209)
        const bool tc_flow = true;
        const bool flow_v2 = h->precision == 1 && h->opts["no_flow_v2"] == 0;
        for (int s = 0; s < A.n_flow; s++) {
            auto& f = h->flows[s];
            ConvArgs a = base_args(f.pre, P, C, f.xcol, fh, H, 0);
            if ((rc = launch_conv(h, a, T1, tc_flow))) return rc;
            if (flow_v2 && f.v2) {
                // bf16 tail: gate outputs of all layers side by side as MMA-operand rows [Fr, layers * H]; res_skip keeps its
                // residual half; m = one GEMM over all of them (K slices = layers), subtracted from the coupled half in place
                const int L = A.wn_layers;
                if ((rc = ensure(h, h->facts_b, (size_t)Fr * L * H * 2))) return rc;
                __nv_bfloat16* fb = ptr<__nv_bfloat16>(h->facts_b);
                for (int i = 0; i < L; i++) {
                    a = base_args(f.in[i], fh, H, 0, facts, L * H, i * H); a.epi = EPI_GATE;
                    a.outb = fb; a.outb_slope = 1.f;
                    if (A.n_speakers > 1) { a.utab = f.cond_tab[i]; a.uidx = d_sid; a.utab_ld = 2 * H; }
                    if ((rc = launch_conv(h, a, T1, true))) return rc;
                    if (i < L - 1) {
                        a = base_args(f.rsr[i], nullptr, 0, i * H, fh, H, 0); a.accumulate = 1;
                        a.xb = fb; a.ldxb = L * H;
                        if ((rc = launch_conv(h, a, T1, true))) return rc;
                    }
                }
                a = base_args(f.mskip[0], nullptr, 0, 0, P, C, f.ocol); a.epi = EPI_SUBFROM; a.res = P; a.ldres = C; a.rescol = f.ocol;
                a.bias = f.mskip_b; a.xb = fb; a.ldxb = L * H;
                a.nks = L; for (int i = 0; i < L; i++) a.wtc_ks[i] = f.mskip[i].wtc;
                if ((rc = launch_conv(h, a, T1, true))) return rc;
                continue;
            }
            for (int i = 0; i < A.wn_layers; i++) {
                a = base_args(f.in[i], fh, H, 0, facts, H, 0); a.epi = EPI_GATE;
                if (A.n_speakers > 1) { a.utab = f.cond_tab[i]; a.uidx = d_sid; a.utab_ld = 2 * H; }
                if ((rc = launch_conv(h, a, T1, tc_flow))) return rc;
                if (i < A.wn_layers - 1) {
                    a = base_args(f.rs[i], facts, H, 0, fh, H, 0); a.epi = EPI_SPLIT; a.split = H; a.accumulate = 1;
                    a.out2 = fskip; a.ldo2 = H; a.ocol2 = 0; a.accumulate2 = (i > 0);
                } else {
                    a = base_args(f.rs[i], facts, H, 0, fskip, H, 0); a.accumulate = (i > 0);
                }
                if ((rc = launch_conv(h, a, T1, tc_flow))) return rc;
            }
            a = base_args(f.post, fskip, H, 0, P, C, f.ocol); a.epi = EPI_SUBFROM; a.res = P; a.ldres = C; a.rescol = f.ocol;
            if ((rc = launch_conv(h, a, T1, tc_flow))) return rc;
        }
        stage_end(h);
        // ---- HiFi-GAN generator (models.py:348-368)
        stage_begin(h, 2);
        float *dpre = ptr<float>(h->dpre), *X = ptr<float>(h->sX), *T1b = ptr<float>(h->sT1);
        float *XSab[2] = {ptr<float>(h->sXSa), ptr<float>(h->sXSb)};
        float *Ya = ptr<float>(h->sYa), *Yb = ptr<float>(h->sYb);
        {
            ConvArgs a = base_args(h->dec_pre, P, C, 0, dpre, A.up_init, 0);
            if (A.n_speakers > 1) { a.utab = h->dec_cond_tab; a.uidx = d_sid; a.utab_ld = A.up_init; }
            if ((rc = launch_conv(h, a, T1, true))) return rc;
        }
        const float* cur = dpre; int cur_c = A.up_init;
        const __nv_bfloat16* cur_b = nullptr;          // previous stage's output as bf16 lrelu rows (feeds a fused ConvTranspose)
        bool cur_is_b = false;                         // ... and ONLY in that form (`cur` is not valid)
        for (int i = 0; i < A.n_ups; i++) {
            auto& U = h->ups[i];
            const Tiles Tin = TR[i];
            const Tiles Tout = TR[i + 1];
            const int co = U.cout, u = U.rate;
            float* XS = XSab[i & 1];      // stage output; the next stage reads it while writing the other one
            // polyphase ConvTranspose1d (models.py:320-332): output row-block q holds u*co contiguous floats
            // == rows q*u .. q*u+u-1 of the [rows*u, co] result; phases [0,u/2) use taps {-1,0}, the rest {0,+1}
            ConvArgs a;
            __nv_bfloat16* Xb = reinterpret_cast<__nv_bfloat16*>(X);          // v3: the stage input as bf16 lrelu rows
            if (!up_fused[i + 1]) {
                for (int half = 0; half < 2; half++) {
                    a = base_args(half ? U.B : U.A, cur, cur_c, 0, X, u * co, half * (u / 2) * co); a.in_act = 1; a.in_slope = 0.1f;
                    if (cur_is_b) { a.x = nullptr; a.ldx = 0; a.in_act = 0; a.xb = cur_b; a.ldxb = cur_c; }   // already lrelu'd bf16 operand rows
                    if (mrf_on[i + 1] == 3 || rb_bf16[i + 1]) { a.outb = Xb; a.outb_slope = 0.1f; }
                    if ((rc = launch_conv(h, a, Tin, true))) return rc;
                }
            }
            bool out_is_b = false;
            if (mrf_on[i + 1] == 3) {
                Mrf3Args& m = mrf3_args[i + 1];
                m.cu = Tout.cu; m.tile_cu = Tout.tx; m.B = Tout.B; m.rate = Tout.rate; m.ntiles = Tout.nx;
                if (up_fused[i + 1]) {
                    m.hb = cur_b; m.up_w[0] = U.A.wtc; m.up_w[1] = U.B.wtc; m.up_b = U.A.b;     // bias tiled per phase: first C entries
                    m.in_rows = (long)Fr * rates[i];
                } else { m.xb = Xb; m.in_rows = (long)Fr * rates[i + 1]; }
                m.out = nullptr; m.outb = nullptr; m.post_w = nullptr; m.audio = nullptr;
                if (post_fused && i == A.n_ups - 1) { m.post_w = h->post_w; m.post_slope = 0.01f; m.audio = audio + (int64_t)f_lo * hop; }
                else if (i + 1 < A.n_ups && up_fused[i + 2]) { m.outb = reinterpret_cast<__nv_bfloat16*>(XS); m.outb_slope = 0.1f; }
                else m.out = XS;
                m.dbg = nullptr;
                if (h->opts.count("mrf_dbg") && (int)h->opts["mrf_dbg"] == i + 1) {
                    if ((rc = ensure(h, h->mrf_dbg, (size_t)MRF3_DBG_TILES * 48 * 8))) return rc;
                    CK(h, cudaMemsetAsync(h->mrf_dbg.p, 0, (size_t)MRF3_DBG_TILES * 48 * 8, st));
                    m.dbg = ptr<unsigned long long>(h->mrf_dbg);
                }
                if (m.ntiles > 0) {
                    if ((rc = ensure(h, h->tdesc, (size_t)m.ntiles * sizeof(int4)))) return rc;
                    m.tdesc = ptr<int4>(h->tdesc);
                    const int kw = (i == A.n_ups - 1) ? 0 : 1;
                    double mac = 0.0;
                    for (int j = 0; j < A.n_rbk; j++) mac += 2.0 * A.rb_kernels[j] * co * co;
                    mac *= rates[i + 1];
                    if (up_fused[i + 1]) mac += (double)rates[i] * U.A.cin * co * (2 * u);
                    if (m.post_w) mac += (double)rates[i + 1] * co * 7;
                    h->kern_macs[kw] += mac * Fr; h->kern_launches[kw]++;
                    cudaError_t e = mrf3_launch_timed(h, m, mrf3_cfg[i + 1], 3 + kw);
                    if (e != cudaSuccess) return fail(h, VITS_E_CUDA, "mrf3 launch: %s", cudaGetErrorString(e));
                    h->launches += 2;
                }
            } else if (stage_sum2[i + 1]) {
                // ResBlock2 stage (modules.py:355-364) in n_r + 1 launches: x1_r = x + conv_{k_r,d_r1}(lrelu x) for every resblock, written
                // side by side as bf16 lrelu rows [rows, n_r * C]; then ONE launch sums the second convs of all resblocks in TMEM
                // (K slices = resblocks, each with its own taps), adds sum_r x1_r recovered from the same rows and divides by n_r.
                // No fp32 read-modify-write of the stage output, 4 epilogues instead of 6.
                const int nr = A.n_rbk;
                const size_t rows = (size_t)Fr * rates[i + 1];
                if ((rc = ensure(h, h->sX1b, rows * nr * co * 2))) return rc;
                __nv_bfloat16* X1 = ptr<__nv_bfloat16>(h->sX1b);
                for (int j = 0; j < nr; j++) {
                    a = base_args(h->rb_c1[i * nr + j][0], nullptr, 0, 0, T1b, nr * co, j * co);
                    a.xb = Xb; a.ldxb = co; a.resb = Xb; a.ldresb = co; a.resb_slope = 0.1f;
                    a.outb = X1; a.outb_slope = 0.1f;
                    if ((rc = launch_conv(h, a, Tout, true))) return rc;
                }
                a = base_args(h->rb_c1[i * nr][1], nullptr, 0, 0, XS, co, 0);
                a.ntaps = 0;
                for (int j = 0; j < nr; j++) {
                    const ConvP& cv = h->rb_c1[i * nr + j][1];
                    a.tap0_ks[j] = a.ntaps; a.ntaps_ks[j] = cv.ntaps; a.wtc_ks[j] = cv.wtc;
                    for (int q = 0; q < cv.ntaps; q++) a.toff[a.ntaps++] = cv.toff[q];
                }
                a.nks = nr; a.bias = h->rb_b2sum[i]; a.out_div = (float)nr;
                a.xb = X1; a.ldxb = nr * co; a.resb = X1; a.ldresb = nr * co; a.resb_slope = 0.1f; a.nresb = nr; a.resb_stride = co;
                if (i + 1 < A.n_ups) { a.outb = reinterpret_cast<__nv_bfloat16*>(XS); a.outb_slope = 0.1f; out_is_b = true; }
                if ((rc = launch_conv(h, a, Tout, true))) return rc;
            } else if (rb1_fused[i + 1]) {
                // one fused launch per (conv_{k,d} -> conv_{k,1}) pair; x travels between pairs as bf16 lrelu rows, the last pair of
                // each resblock adds its result into the fp32 stage output (divided by n_r by the last one): the same dataflow as the
                // conv-by-conv path below, minus the intermediate's round trip and one launch per pair
                // stages followed by another ConvTranspose hand over bf16 rows (below); the last stage feeds conv_post in fp32
                // (the last stage without the opt-in conv_post fusion: the same hand-over between the resblocks, but the combined result is
                // written as fp32 -- conv_post's input stays what it was)
                // OPT-IN like the conv_post fusion (option rb1_last_rows_out): +2.9 % on C4 for 3 dB of the worst utterance's SNR (51.3 ->
                // 48.2 dB, profiles/r02zq...) -- nothing filters the last stage's rounding noise before conv_post.
                const bool last_rows = (i + 1 == A.n_ups) && !rb1_post[i + 1] && h->opts["rb1_last_rows_out"] != 0;
                const bool rows_out = ((i + 1 < A.n_ups) || rb1_post[i + 1] || last_rows) && A.n_rbk >= 2 && A.n_rbk <= 3 && h->opts["no_rb1_rows_out"] == 0;
                const bool fp32_out = rows_out && last_rows;
                __nv_bfloat16 *R0 = reinterpret_cast<__nv_bfloat16*>(T1b), *R1 = nullptr;
                if (rows_out) {
                    if ((rc = ensure(h, h->sX1b, (size_t)Fr * rates[i + 1] * co * 2))) return rc;
                    R1 = ptr<__nv_bfloat16>(h->sX1b);
                    out_is_b = !fp32_out;
                }
                if (Tout.nx > 0) {
                    if ((rc = ensure(h, h->tdesc, (size_t)Tout.nx * sizeof(int4)))) return rc;
                    size_t q = 0;
                    for (int j = 0; j < A.n_rbk; j++) {
                        const bool first = (j == 0), last = (j == A.n_rbk - 1);
                        const int nd = A.rb_ndil[j];
                        for (int c2 = 0; c2 < nd; c2++, q++) {
                            const bool fin = (c2 == nd - 1);
                            Mrf3Args m = rb1_args[i + 1][q];
                            const Mrf3Cfg& cf = rb1_cfg[i + 1][q];
                            m.cu = Tout.cu; m.tile_cu = Tout.tx; m.B = Tout.B; m.rate = Tout.rate; m.ntiles = Tout.nx;
                            m.tdesc = ptr<int4>(h->tdesc);
                            m.xb = (c2 == 0) ? Xb : reinterpret_cast<const __nv_bfloat16*>((c2 & 1) ? Ya : Yb);
                            m.in_rows = (long)Fr * rates[i + 1];
                            if (fin && rows_out) {
                                // the per-resblock results leave as plain bf16 rows; the last resblock's last pair adds them, divides by n_r
                                // and writes the next ConvTranspose's bf16 lrelu operand rows -- 10 bytes per value of stage output instead
                                // of 24 (three fp32 read-modify-writes + the fp32 read of the consumer; r02zj launch list: ~0.45 ms of
                                // every accumulating launch)
                                if (!last) { m.outb = (j == 0) ? R0 : R1; m.outb_slope = 1.f; }
                                else {
                                    m.addb[0] = R0; m.addb[1] = R1; m.naddb = A.n_rbk - 1; m.out_div = (float)A.n_rbk;
                                    if (rb1_post[i + 1]) { m.post_w = h->post_w; m.post_slope = 0.01f; m.audio = audio + (int64_t)f_lo * hop; }
                                    else if (fp32_out) m.out = XS;
                                    else { m.outb = reinterpret_cast<__nv_bfloat16*>(XS); m.outb_slope = 0.1f; }
                                }
                            } else if (fin) { m.out = XS; m.accumulate = !first; m.out_div = last ? (float)A.n_rbk : 1.f; }
                            else { m.outb = reinterpret_cast<__nv_bfloat16*>((c2 & 1) ? Yb : Ya); m.outb_slope = 0.1f; }
                            CUtensorMap tm; memset(&tm, 0, sizeof tm);
                            if (cf.tma && !make_rows_tmap(&tm, m.xb, m.in_rows, co, cf.box_rows)) return fail(h, VITS_E_CUDA, "rb1 pair: tensor map");
                            cudaError_t e = (q == 0) ? mrf3_tiles_launch(m, cf, st) : cudaSuccess;      // one tile table for the whole stage
                            if (e == cudaSuccess) e = mrf3_kernel_launch(m, cf, tm, h->num_sms, st);
                            if (e != cudaSuccess) return fail(h, VITS_E_CUDA, "rb1 pair launch: %s", cudaGetErrorString(e));
                            h->launches += (q == 0) ? 2 : 1;
                        }
                    }
                }
            } else
            for (int j = 0; j < A.n_rbk; j++) {
                const int n = i * A.n_rbk + j;
                const bool first = (j == 0), last = (j == A.n_rbk - 1);
                const int nd = A.rb_ndil[j];
                const float* in = X;
                for (int c = 0; c < nd; c++) {
                    const bool fin = (c == nd - 1);
                    if (A.resblock_type == 2 && rb_bf16[i + 1]) {
                        // modules.py:355-364 with bf16 operand rows between the convs: the input, the residual (recovered from
                        // the operand, lrelu is invertible) and the intermediate x1 are 2 bytes per value; the per-resblock
                        // results still accumulate in fp32
                        __nv_bfloat16* inb = (c == 0) ? Xb : reinterpret_cast<__nv_bfloat16*>((c & 1) ? T1b : Ya);
                        float* dstf = (c & 1) ? Ya : T1b;
                        a = base_args(h->rb_c1[n][c], nullptr, 0, 0, fin ? XS : dstf, co, 0);
                        a.xb = inb; a.ldxb = co; a.resb = inb; a.ldresb = co; a.resb_slope = 0.1f;
                        if (fin) { a.accumulate = !first; if (last) a.out_div = (float)A.n_rbk; }
                        else { a.outb = reinterpret_cast<__nv_bfloat16*>(dstf); a.outb_slope = 0.1f; }
                        if ((rc = launch_conv(h, a, Tout, true))) return rc;
                    } else if (A.resblock_type == 2) {
                        // modules.py:355-364: x = conv_d(lrelu(x)) + x
                        float* dst = fin ? XS : ((c & 1) ? Ya : T1b);
                        a = base_args(h->rb_c1[n][c], in, co, 0, dst, co, 0); a.in_act = 1; a.in_slope = 0.1f;
                        a.res = in; a.ldres = co;
                        if (fin) { a.accumulate = !first; if (last) a.out_div = (float)A.n_rbk; }
                        if ((rc = launch_conv(h, a, Tout, true))) return rc;
                        in = dst;
                    } else if (rb_bf16[i + 1]) {
                        // modules.py:301-314 on bf16 operand rows: xt = c1(lrelu x) leaves as bf16 lrelu(xt) -- exactly c2's MMA operand;
                        // c2 adds the residual x recovered from ITS operand rows (lrelu is invertible) and writes the next x the same
                        // way; only the per-resblock result accumulates in fp32.  2 B per value per conv instead of 12, no conversion pass.
                        __nv_bfloat16* inb = (c == 0) ? Xb : reinterpret_cast<__nv_bfloat16*>((c & 1) ? Ya : Yb);
                        __nv_bfloat16* tb = reinterpret_cast<__nv_bfloat16*>(T1b);
                        a = base_args(h->rb_c1[n][c], nullptr, 0, 0, T1b, co, 0);
                        a.xb = inb; a.ldxb = co; a.outb = tb; a.outb_slope = 0.1f;
                        if ((rc = launch_conv(h, a, Tout, true))) return rc;
                        float* dstf = (c & 1) ? Yb : Ya;
                        a = base_args(h->rb_c2[n][c], nullptr, 0, 0, fin ? XS : dstf, co, 0);
                        a.xb = tb; a.ldxb = co; a.resb = inb; a.ldresb = co; a.resb_slope = 0.1f;
                        if (fin) { a.accumulate = !first; if (last) a.out_div = (float)A.n_rbk; }
                        else { a.outb = reinterpret_cast<__nv_bfloat16*>(dstf); a.outb_slope = 0.1f; }
                        if ((rc = launch_conv(h, a, Tout, true))) return rc;
                    } else {
                        // modules.py:301-314: xt = c1(lrelu(x)); xt = c2(lrelu(xt)); x = xt + x
                        a = base_args(h->rb_c1[n][c], in, co, 0, T1b, co, 0); a.in_act = 1; a.in_slope = 0.1f;
                        if ((rc = launch_conv(h, a, Tout, true))) return rc;
                        float* dst = fin ? XS : ((c & 1) ? Yb : Ya);
                        a = base_args(h->rb_c2[n][c], T1b, co, 0, dst, co, 0); a.in_act = 1; a.in_slope = 0.1f;
                        a.res = in; a.ldres = co;
                        if (fin) { a.accumulate = !first; if (last) a.out_div = (float)A.n_rbk; }
                        if ((rc = launch_conv(h, a, Tout, true))) return rc;
                        in = dst;
                    }
                }
            }
            cur = XS; cur_c = co; cur_b = reinterpret_cast<const __nv_bfloat16*>(XS);
            cur_is_b = out_is_b;
        }
        // ---- lrelu(0.01) -> conv_post -> tanh (models.py:364-366)
        {
            const Tiles Tl = TR[A.n_ups];
            if (Tl.n256 > 0 && !post_fused) {
                size_t smem = (size_t)(CP_TILE + 6) * (h->post_c + 1) * sizeof(float);
                launch_k(k_conv_post, Tl.n256, 256, smem, st, cur, h->post_c, h->post_w, Tl.cu, Tl.t256, nB, Tl.rate, 0.01f,
                                                        audio + (int64_t)f_lo * hop);
                h->launches++;
            }
        }
        CK(h, cudaGetLastError());
        stage_end(h);
        if (out_kind == 2) {
            // caller-side post-processing on device (voice.py:271-282, 88-91) for this chunk's utterances
            dim3 g(32, (unsigned)std::min(nB, 65535));
            launch_k(k_absmax, g, 256, 0, st, audio, ptr<int>(h->cu_y_dev) + b_lo, nB, hop, ptr<unsigned int>(h->peaks) + b_lo);
            launch_k(k_to_int16, g, 256, 0, st, audio, ptr<int>(h->cu_y_dev) + b_lo, nB, hop, ptr<unsigned int>(h->peaks) + b_lo, normalize, volume, audio16);
            h->launches += 2;
            CK(h, cudaGetLastError());
        }
        if (out_kind == 1 || out_kind == 2) {
            // this chunk's audio leaves on the copy stream while the next chunk computes
            CK(h, cudaEventRecord(h->ev_chunk, st));
            CK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_chunk, 0));
            if (scatter) {
                // every utterance of the chunk straight to ITS place in the caller's buffer (e.g. a job-wide result buffer shared by
                // several processes, in original utterance order): the DMA engine does the gather, no host memcpy afterwards
                for (int b = b_lo; b < b_hi; b++) {
                    const int64_t s0 = (int64_t)h->h_cu_y[b] * hop, ns = (int64_t)h->h_ylen[b] * hop;
                    if (out_kind == 1) CK(h, cudaMemcpyAsync(static_cast<float*>(out) + offs[b], audio + s0, (size_t)ns * 4, cudaMemcpyDeviceToHost, h->copy_stream));
                    else CK(h, cudaMemcpyAsync(static_cast<int16_t*>(out) + offs[b], audio16 + s0, (size_t)ns * 2, cudaMemcpyDeviceToHost, h->copy_stream));
                }
            } else if (out_kind == 1)
                CK(h, cudaMemcpyAsync(static_cast<float*>(out) + (int64_t)f_lo * hop, audio + (int64_t)f_lo * hop, (size_t)Fr * hop * 4,
                                      cudaMemcpyDeviceToHost, h->copy_stream));
            else
                CK(h, cudaMemcpyAsync(static_cast<int16_t*>(out) + (int64_t)f_lo * hop, audio16 + (int64_t)f_lo * hop, (size_t)Fr * hop * 2,
                                      cudaMemcpyDeviceToHost, h->copy_stream));
        } else if (out_kind == 3) {
            // device-resident result: `out` is a device pointer of the caller; stream-ordered, no host synchronisation
            CK(h, cudaMemcpyAsync(static_cast<float*>(out) + (int64_t)f_lo * hop, audio + (int64_t)f_lo * hop, (size_t)Fr * hop * 4,
                                  cudaMemcpyDeviceToDevice, st));
        }
        b_lo = b_hi;
    }
    // ---- output
    if (out_kind == 1 || out_kind == 2) {
        CK(h, cudaEventRecord(h->ev_out[asel], h->copy_stream));
        h->out_pending[asel] = true;
        h->ticket_of[asel] = ++h->ticket;
        if (!async_out) { CK(h, cudaEventSynchronize(h->ev_out[asel])); h->out_pending[asel] = false; }
    }
    return VITS_OK;
}

int64_t vits_fetch(vits_handle* h, const char* name, void* out, int64_t capacity_elems) {
    if (!h || !name || !out) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->prepared) return fail(h, VITS_E_STATE, "nothing to fetch: no successful vits_prepare()");
    const vits_arch& A = h->A;
    const std::string k(name);
    const void* src = nullptr; int64_t n = 0;
    const int64_t R = h->R, Fr = h->last_chunk_frames;
    if (k == "x") { src = h->x.p; n = R * A.hidden; }
    else if (k == "stats") { src = h->stats.p; n = R * 2 * A.inter; }
    else if (k == "logw") { src = h->logw.p; n = R; }
    else if (k == "durations") { src = h->dur.p; n = R; }
    else if (k == "cum") { src = h->cum.p; n = R; }
    else if (k == "noise_dp") { src = h->dbg_dp.p; n = h->dbg_dp.p ? 2 * R : 0; }
    else if (k == "frame_index" || k == "z_p" || k == "z") {
        // per-chunk workspaces: after a multi-chunk decode they hold the LAST chunk only -- refuse instead of returning a fragment
        if (h->last_nchunks != 1)
            return fail(h, VITS_E_STATE, "stage tensor '%s' is per chunk and the last vits_decode used %d chunks (raise max_chunk_frames or fetch after a single-chunk call)", name, h->last_nchunks);
        if (k == "frame_index") { src = h->fidx.p; n = Fr; }
        else if (k == "z_p") { src = h->dbg_zp.p; n = h->dbg_zp.p ? Fr * A.inter : 0; }
        else { src = h->P.p; n = Fr * A.inter; }
    }
    else if (k == "conv_dbg") { src = h->conv_dbg.p; n = h->conv_dbg.p ? TC_DBG_TILES * 16 * 2 : 0; }
    else if (k == "mrf_dbg") { src = h->mrf_dbg.p; n = h->mrf_dbg.p ? MRF3_DBG_TILES * 48 * 2 : 0; }
    else return fail(h, VITS_E_INVALID, "unknown stage tensor '%s'", name);
    if (!src || n == 0) return fail(h, VITS_E_STATE, "stage tensor '%s' is not available", name);
    if (n > capacity_elems) return fail(h, VITS_E_INVALID, "fetch '%s': capacity %lld < %lld", name, (long long)capacity_elems, (long long)n);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaMemcpy(out, src, n * 4, cudaMemcpyDeviceToHost));
    return n;
}

int vits_timer_start(vits_handle* h) {
    if (!h) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    resolve_stage_events(h);
    for (float& v : h->stage_ms) v = 0.f;
    h->kern_launches[0] = h->kern_launches[1] = 0; h->kern_macs[0] = h->kern_macs[1] = 0.0;
    CK(h, cudaEventRecord(h->ev_t0, h->stream));
    return VITS_OK;
}

int vits_timer_stop(vits_handle* h, float* elapsed_ms) {
    if (!h || !elapsed_ms) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaEventRecord(h->ev_t1, h->stream));
    CK(h, cudaEventSynchronize(h->ev_t1));
    CK(h, cudaEventElapsedTime(elapsed_ms, h->ev_t0, h->ev_t1));
    resolve_stage_events(h);
    return VITS_OK;
}

int vits_stage_ms(vits_handle* h, float* text_ms, float* flow_ms, float* dec_ms) {
    if (!h) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    resolve_stage_events(h);
    if (text_ms) *text_ms = h->stage_ms[0];
    if (flow_ms) *flow_ms = h->stage_ms[1];
    if (dec_ms) *dec_ms = h->stage_ms[2];
    return VITS_OK;
}

int vits_kernel_ms(vits_handle* h, int which, float* ms, int64_t* launches, double* macs) {
    if (!h || which < 0 || which > 1) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    resolve_stage_events(h);
    if (ms) *ms = h->stage_ms[3 + which];
    if (launches) *launches = h->kern_launches[which];
    if (macs) *macs = h->kern_macs[which];
    return VITS_OK;
}

// Test hook: one convolution through the production launch path (fp32 or tcgen05 kernel) on
// host data, single utterance of L rows.  Used by tests/ to pin both conv kernels against numpy.
int vits_test_conv(vits_handle* h, int use_tc, const float* x, int L, int cin, const int* taps, int ntaps,
                   const float* w32, const uint16_t* wtc, const float* bias, int n, int in_act, float in_slope,
                   int epi, const float* res, int accumulate, float out_div, int out_act, float* out, int out_cols) {
    if (!h || !x || !taps || !w32 || !out || ntaps < 1 || ntaps > CONV_MAX_TAPS) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    CK(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int npad = rup(n, 4), npad16 = rup(n, 16);
    float *dx = nullptr, *dw = nullptr, *db = nullptr, *dres = nullptr, *dout = nullptr; void* dwtc = nullptr; int* dmeta = nullptr;
    const size_t out_elems = (size_t)L * out_cols;
    CK(h, cudaMalloc(&dx, (size_t)L * cin * 4)); CK(h, cudaMalloc(&dw, (size_t)ntaps * cin * npad * 4));
    CK(h, cudaMalloc(&dout, out_elems * 4)); CK(h, cudaMalloc(&dmeta, 64));
    CK(h, cudaMemcpy(dx, x, (size_t)L * cin * 4, cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(dw, w32, (size_t)ntaps * cin * npad * 4, cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(dout, out, out_elems * 4, cudaMemcpyHostToDevice));
    if (bias) { CK(h, cudaMalloc(&db, npad * 4)); CK(h, cudaMemcpy(db, bias, npad * 4, cudaMemcpyHostToDevice)); }
    if (res) { CK(h, cudaMalloc(&dres, (size_t)L * n * 4)); CK(h, cudaMemcpy(dres, res, (size_t)L * n * 4, cudaMemcpyHostToDevice)); }
    const size_t wtc_bytes = (size_t)ntaps * cin * npad16 * 2 * (use_tc == 2 ? 3 : 1);
    if (wtc) { CK(h, cudaMalloc(&dwtc, wtc_bytes)); CK(h, cudaMemcpy(dwtc, wtc, wtc_bytes, cudaMemcpyHostToDevice)); }
    int cu[2] = {0, L};
    TileBuilder tb; tb.begin(cu, 1); tb.add(1);
    CK(h, cudaMemcpy(dmeta, tb.host.data(), tb.host.size() * 4, cudaMemcpyHostToDevice));
    Tiles T = tb.get(dmeta, 1);
    { int rc0; if ((rc0 = ensure(h, h->tdesc_t, (size_t)std::max(T.n128, 1) * sizeof(int4))) || (rc0 = make_d128(h, T, ptr<int4>(h->tdesc_t)))) return rc0; }
    ConvP c; c.w = dw; c.wtc = (const __nv_bfloat16*)dwtc; c.b = db; c.cin = cin; c.n = n; c.npad = npad; c.npad16 = npad16; c.ntaps = ntaps;
    if (use_tc == 2) {
        // `wtc` holds the K slices back to back ("<n>.wtc3.0", ".1", ...), slice width as in mkconv
        int sl = 0;
        for (int v = std::min(cin, SPLIT3_MAX_CIN); v >= 16; v--) if (cin % v == 0 && v % 16 == 0) { sl = v; break; }
        if (!sl || cin / sl > CONV_MAX_SLICES) { cudaFree(dx); cudaFree(dw); cudaFree(dout); cudaFree(dmeta); return fail(h, VITS_E_INVALID, "bf16x3: unsupported cin %d", cin); }
        c.nsl = cin / sl; c.slice_cin = sl;
        for (int j = 0; j < c.nsl; j++) c.wtc3[j] = (const __nv_bfloat16*)dwtc + (size_t)j * ntaps * 3 * sl * npad16;
    }
    for (int i = 0; i < ntaps; i++) c.toff[i] = taps[i];
    ConvArgs a = base_args(c, dx, cin, 0, dout, out_cols, 0);
    a.in_act = in_act; a.in_slope = in_slope; a.epi = epi; a.res = dres; a.ldres = n; a.accumulate = accumulate;
    a.out_div = out_div; a.out_act = out_act;
    if (epi == EPI_SPLIT) { a.split = n / 2; a.out2 = dout + n / 2; a.ldo2 = out_cols; a.ocol2 = 0; a.accumulate2 = accumulate; }
    const int saved = h->precision;
    h->precision = use_tc ? 1 : 0;
    int rc = 0;
    if (use_tc && !conv_tc_supported(a)) rc = fail(h, VITS_E_INVALID, "shape not supported by the tcgen05 conv kernel");
    if (!rc) rc = (use_tc == 2) ? launch_conv_text(h, c, a, T) : launch_conv(h, a, T, use_tc != 0);
    h->precision = saved;
    cudaError_t e = cudaStreamSynchronize(st);
    if (!rc && e != cudaSuccess) rc = fail(h, VITS_E_CUDA, "test conv failed: %s", cudaGetErrorString(e));
    if (!rc) cudaMemcpy(out, dout, out_elems * 4, cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dw); cudaFree(dout); cudaFree(dmeta); if (db) cudaFree(db); if (dres) cudaFree(dres); if (dwtc) cudaFree(dwtc);
    return rc;
}

int vits_describe(vits_handle* h, vits_info* info) {
    if (!h || !info) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    const vits_arch& A = h->A;
    memset(info, 0, sizeof *info);
    int hop = 1; for (int i = 0; i < A.n_ups; i++) hop *= A.up_rates[i];
    info->n_vocab = A.n_vocab; info->n_speakers = A.n_speakers; info->has_sid = A.n_speakers > 1;
    info->hidden = A.hidden; info->inter = A.inter; info->sample_rate = A.sample_rate; info->hop = hop;
    info->resblock_type = A.resblock_type; info->use_sdp = A.use_sdp; info->precision = h->precision;
    info->device = h->device; info->num_sms = h->num_sms; info->finalized = h->finalized ? 1 : 0;
    info->has_scales = h->graph_has_scales; info->has_langid = h->graph_has_langid;
    return VITS_OK;
}

int64_t vits_max_output_samples(vits_handle* h, int64_t sum_ids, float length_scale) {
    if (!h) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    const vits_arch& A = h->A;
    int64_t hop = 1; for (int i = 0; i < A.n_ups; i++) hop *= A.up_rates[i];
    // after a successful vits_prepare the frame count is known exactly (the durations are data): that is what vits_decode needs
    if (sum_ids <= 0) return h->prepared ? h->total_frames * hop : (int64_t)VITS_E_STATE;
    // before it, only a planning figure exists: durations are ceil(exp(logw) * length_scale) per id, unbounded in principle
    // (the kernel clamps a single id at 10^6 frames); 16 frames per id at length_scale 1 has never been exceeded by a voice we have
    // seen (mean 3.5, p99 6, max 12) -- callers must still size the real buffer from vits_prepare's total_frames
    const double per_id = 16.0 * (length_scale > 0.f ? (double)length_scale : 1.0);
    return (int64_t)std::ceil((double)sum_ids * per_id) * hop;
}

int vits_set_stream(vits_handle* h, void* cuda_stream) {
    if (!h) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));          // nothing of ours is left behind on the old stream
    resolve_stage_events(h);
    h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    return VITS_OK;
}

int64_t vits_output_ticket(vits_handle* h) {
    if (!h) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    return h->ticket;
}

int vits_wait_ticket(vits_handle* h, int64_t ticket) {
    if (!h) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    CK(h, cudaSetDevice(h->device));
    // a transfer older than the two tracked ones finished before its device buffer was reused (vits_decode waits for that)
    for (int i = 0; i < 2; i++)
        if (h->ticket_of[i] == ticket && h->out_pending[i]) { CK(h, cudaEventSynchronize(h->ev_out[i])); h->out_pending[i] = false; }
    if (ticket > h->ticket) return fail(h, VITS_E_INVALID, "ticket %lld has not been issued (last: %lld)", (long long)ticket, (long long)h->ticket);
    return VITS_OK;
}

int vits_wait_output(vits_handle* h, int older_only) {
    if (!h) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    CK(h, cudaSetDevice(h->device));
    for (int i = 0; i < 2; i++) {
        if (older_only && i == h->audio_sel) continue;      // the most recent vits_decode keeps running
        if (h->out_pending[i]) { CK(h, cudaEventSynchronize(h->ev_out[i])); h->out_pending[i] = false; }
    }
    return VITS_OK;
}

int vits_set_output_offsets(vits_handle* h, const int64_t* sample_offsets, int32_t B) {
    if (!h || B < 0 || (B > 0 && !sample_offsets)) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    h->out_offsets.assign(sample_offsets, sample_offsets + B);
    return VITS_OK;
}

int vits_host_register(void* p, size_t nbytes) {
    if (!p || nbytes == 0) return VITS_E_INVALID;
    if (cudaHostRegister(p, nbytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return VITS_E_CUDA; }
    return VITS_OK;
}

int vits_host_unregister(void* p) {
    if (!p) return VITS_OK;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return VITS_E_CUDA; }
    return VITS_OK;
}

int vits_host_alloc(size_t nbytes, void** out) {
    if (!out || nbytes == 0) return VITS_E_INVALID;
    void* p = nullptr;
    if (cudaHostAlloc(&p, nbytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return VITS_E_NOMEM; }
    *out = p;
    return VITS_OK;
}

int vits_host_free(void* p) {
    if (!p) return VITS_OK;
    return cudaFreeHost(p) == cudaSuccess ? VITS_OK : VITS_E_CUDA;
}

int vits_test_mma_probe(vits_handle* h, int n, int iters, int nd, int na, int rows, int nctas, double* issue_cycles, double* total_cycles) {
    return vits_test_mma_probe_mode(h, 0, n, iters, nd, na, rows, nctas, issue_cycles, total_cycles);
}

int vits_test_mma_probe_mode(vits_handle* h, int mode, int n, int iters, int nd, int na, int rows, int nctas, double* issue_cycles, double* total_cycles) {
    if (!h || n < 16 || n > 256 || n % 16 || iters < 1 || nd < 1 || nd * n > ((mode >= 2 && mode != 6) ? 384 : 512) || rows < 128 + 8 * na || rows > 700 || nctas < 1 ||
        mode < 0 || mode > 7 || (mode >= 6 && (n % 32 || nctas % 2))) return VITS_E_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    CK(h, cudaSetDevice(h->device));
    unsigned long long* d = nullptr;
    CK(h, cudaMalloc(&d, (size_t)nctas * 16));
    const int smem = 8 * rows * 16 + 8 * 256 * 16 + 2048;
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        kern<<<nctas, 128, smem, h->stream>>>(n, iters, nd, na, rows, d);
    };
    switch (mode) {
        case 0: go(k_mma_probe<0>); break;
        case 1: go(k_mma_probe<1>); break;
        case 2: go(k_mma_probe<2>); break;
        case 3: go(k_mma_probe<3>); break;
        case 4: go(k_mma_probe<4>); break;
        case 5: go(k_mma_probe<5>); break;
        case 6: go(k_mma_probe2<0>); break;       // clusters of two CTAs (__cluster_dims__)
        default: go(k_mma_probe2<1>); break;
    }
    if (mode >= 6) nctas /= 2;                     // one result per pair
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { cudaFree(d); return fail(h, VITS_E_CUDA, "mma probe: %s", cudaGetErrorString(e)); }
    std::vector<unsigned long long> r(2 * (size_t)nctas);
    cudaMemcpy(r.data(), d, r.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    double a = 0, b = 0;
    for (int i = 0; i < nctas; i++) { a += (double)r[2 * i]; b += (double)r[2 * i + 1]; }
    if (issue_cycles) *issue_cycles = a / nctas / iters;
    if (total_cycles) *total_cycles = b / nctas / iters;
    return VITS_OK;
}

int64_t vits_launch_count(vits_handle* h) { return h ? h->launches : 0; }

const char* vits_last_error(vits_handle* h) { return h ? h->err.c_str() : "null handle"; }

void vits_destroy(vits_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->own_stream && h->own_stream != h->stream) cudaStreamSynchronize(h->own_stream);
    resolve_stage_events(h);
    for (auto& kv : h->blobs) cudaFree(kv.second.p);
    Buf* bufs[] = {&h->ids, &h->tile_t, &h->sid, &h->x, &h->y, &h->qkv, &h->att, &h->ffn, &h->stats, &h->d0, &h->d1,
                   &h->gdp, &h->hp, &h->z0, &h->z1, &h->logw, &h->dur, &h->cum, &h->ylen, &h->inj_dp, &h->inj_z,
                   &h->chunk_meta, &h->tdesc, &h->P, &h->fh, &h->facts, &h->fskip, &h->fidx, &h->dpre, &h->sX, &h->sT1, &h->sYa,
                   &h->sYb, &h->sXSa, &h->sXSb, &h->audio, &h->peaks, &h->audio16, &h->cu_y_dev, &h->dbg_zp, &h->mrf_dbg, &h->conv_dbg, &h->audio_alt, &h->facts_b, &h->tdesc_t, &h->tdesc_c, &h->sX1b, &h->rowpos, &h->fpos, &h->dbg_dp, &h->audio16_alt};
    for (Buf* b : bufs) if (b->p) cudaFree(b->p);
    for (float* pz : h->rb_b2sum) if (pz) cudaFree(pz);
    for (void* pz : h->owned) cudaFree(pz);
    for (auto e : h->event_pool) cudaEventDestroy(e);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->ev_chunk) cudaEventDestroy(h->ev_chunk);
    for (int i = 0; i < 2; i++) if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

}  // extern "C"
