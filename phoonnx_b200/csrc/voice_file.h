// Host-side voice loader of libvits_b200.so: exporter-format file -> canonical tensors -> architecture -> kernel-layout blobs.
//
// This is what lets the C ABI open a voice by itself (vits_open, include/vits_b200.h) -- the entry point SURVEY.md 8b sketched as
// vits_create(path, ...).  It re-states, in C++, the three Python modules a Python host uses (and that stay the reference
// implementation, pinned blob for blob by tests/test_native_loader.py):
//   * phoonnx_b200/onnx_reader.py  -- a minimal protobuf wire reader for the ModelProto / GraphProto / TensorProto / NodeProto
//     fields a file written by phoonnx_train/export_onnx.py:318-350 carries (no onnx / protobuf dependency);
//   * phoonnx_b200/weights.py      -- the exporter's quirks (anonymous weight-normed flow convs recovered through their bias names,
//     Identity aliases of de-duplicated initializers, dp.flows.0.logs surviving only as the folded constant feeding Exp) and the
//     architecture inferred from tensor shapes and Conv attributes;
//   * phoonnx_b200/packing.py      -- kernel layouts: fp32 [tap][C_in][N4], bf16 tcgen05 chunks [tap][C_in/8][N16][8], bf16x3 (hi | lo
//     | hi) slices, Flip folded into the coupling flow's weights, interleaved gate channels, polyphase ConvTranspose halves,
//     per-speaker conditioning tables, the composed m = post(sum skip) GEMM.
// No CUDA in this header: it is exercised without a GPU.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/vits_b200.h"

namespace vf {

struct Tensor { std::vector<int64_t> dims; std::vector<float> v; int64_t dim(int i) const { return i < (int)dims.size() ? dims[i] : 1; } };
struct Node { std::string op, name; std::vector<std::string> in, out; std::map<std::string, std::vector<int64_t>> ints; };
struct Model { std::vector<std::string> inputs; std::map<std::string, std::string> meta; std::map<std::string, Tensor> inits; std::vector<Node> nodes; };
struct Blob { std::vector<uint8_t> bytes; int dtype = 0; };      // dtype: 0 float32, 1 bfloat16 bits (vits_upload)
typedef std::map<std::string, Tensor> Canon;
typedef std::map<std::string, std::map<std::string, std::vector<int64_t>>> ConvAttrs;

// ------------------------------------------------------------------------------------------ protobuf wire format
struct Span { const uint8_t* p; size_t n; };
struct Field { uint32_t no; int wt; uint64_t val; Span sub; };

inline bool varint(const uint8_t*& p, const uint8_t* e, uint64_t& v) {
    v = 0;
    for (int sh = 0; sh < 64 && p < e; sh += 7) { const uint8_t b = *p++; v |= (uint64_t)(b & 0x7F) << sh; if (!(b & 0x80)) return true; }
    return false;
}
// iterate the fields of a message; returns false on malformed input
template <class F> inline bool fields(Span s, F&& fn) {
    const uint8_t* p = s.p; const uint8_t* e = s.p + s.n;
    while (p < e) {
        uint64_t key;
        if (!varint(p, e, key)) return false;
        Field f; f.no = (uint32_t)(key >> 3); f.wt = (int)(key & 7); f.val = 0; f.sub = {nullptr, 0};
        if (f.wt == 0) { if (!varint(p, e, f.val)) return false; }
        else if (f.wt == 1) { if (e - p < 8) return false; memcpy(&f.val, p, 8); f.sub = {p, 8}; p += 8; }
        else if (f.wt == 2) { uint64_t ln; if (!varint(p, e, ln) || (uint64_t)(e - p) < ln) return false; f.sub = {p, (size_t)ln}; p += ln; }
        else if (f.wt == 5) { if (e - p < 4) return false; uint32_t t; memcpy(&t, p, 4); f.val = t; f.sub = {p, 4}; p += 4; }
        else return false;
        if (!fn(f)) return false;
    }
    return true;
}
inline std::string str(Span s) { return std::string(reinterpret_cast<const char*>(s.p), s.n); }
inline void packed_ints(const Field& f, std::vector<int64_t>& out) {
    if (f.wt == 0) { out.push_back((int64_t)f.val); return; }
    const uint8_t* p = f.sub.p; const uint8_t* e = p + f.sub.n; uint64_t v;
    while (p < e && varint(p, e, v)) out.push_back((int64_t)v);
}

// TensorProto: dims (1), data_type (2), float_data (4), int64_data (7), name (8), raw_data (9).  Only float32 tensors carry data out.
inline bool parse_tensor(Span s, std::string& name, Tensor& t, int& dtype) {
    dtype = 1; Span raw = {nullptr, 0}; bool has_raw = false; std::vector<float> fl;
    const bool ok = fields(s, [&](const Field& f) {
        if (f.no == 1) packed_ints(f, t.dims);
        else if (f.no == 2) dtype = (int)f.val;
        else if (f.no == 8) name = str(f.sub);
        else if (f.no == 9) { raw = f.sub; has_raw = true; }
        else if (f.no == 4) {
            if (f.wt == 2) { const size_t n = f.sub.n / 4; const size_t o = fl.size(); fl.resize(o + n); memcpy(fl.data() + o, f.sub.p, n * 4); }
            else { float x; const uint32_t u = (uint32_t)f.val; memcpy(&x, &u, 4); fl.push_back(x); }
        }
        return true;
    });
    if (!ok) return false;
    if (dtype != 1) return true;                       // non-float initializers (shape constants ...) are not weights
    int64_t count = 1; for (int64_t d : t.dims) count *= d;
    if (has_raw) { if ((int64_t)(raw.n / 4) != count) return false; t.v.resize((size_t)count); memcpy(t.v.data(), raw.p, (size_t)count * 4); }
    else { if ((int64_t)fl.size() != count) return false; t.v.swap(fl); }
    return true;
}

inline bool parse_node(Span s, Node& n) {
    std::vector<Span> attrs;
    if (!fields(s, [&](const Field& f) {
            if (f.no == 1) n.in.push_back(str(f.sub)); else if (f.no == 2) n.out.push_back(str(f.sub));
            else if (f.no == 3) n.name = str(f.sub); else if (f.no == 4) n.op = str(f.sub); else if (f.no == 5) attrs.push_back(f.sub);
            return true; })) return false;
    if (n.op != "Conv" && n.op != "ConvTranspose") return true;
    for (Span a : attrs) {                             // AttributeProto: name (1), i (3), ints (8)
        std::string an; std::vector<int64_t> iv; bool has_i = false; int64_t i1 = 0;
        if (!fields(a, [&](const Field& f) {
                if (f.no == 1) an = str(f.sub); else if (f.no == 3) { has_i = true; i1 = (int64_t)f.val; } else if (f.no == 8) packed_ints(f, iv);
                return true; })) return false;
        if (iv.empty() && has_i) iv.push_back(i1);
        if (!iv.empty()) n.ints[an] = iv;
    }
    return true;
}

inline bool parse_model(const std::vector<uint8_t>& data, Model& m, std::string& err) {
    Span graph = {nullptr, 0};
    Span all = {data.data(), data.size()};
    if (!fields(all, [&](const Field& f) {
            if (f.no == 7 && f.wt == 2) graph = f.sub;
            else if (f.no == 14 && f.wt == 2) {        // metadata_props: StringStringEntryProto
                std::string k, v;
                fields(f.sub, [&](const Field& g) { if (g.no == 1) k = str(g.sub); else if (g.no == 2) v = str(g.sub); return true; });
                m.meta[k] = v;
            }
            return true; }) || !graph.p) { err = "not an ONNX ModelProto (no graph)"; return false; }
    std::vector<std::string> g_in;
    bool bad = false;
    if (!fields(graph, [&](const Field& f) {
            if (f.wt != 2) return true;
            if (f.no == 1) { Node n; if (!parse_node(f.sub, n)) { bad = true; return false; } m.nodes.push_back(std::move(n)); }
            else if (f.no == 5) { std::string nm; Tensor t; int dt; if (!parse_tensor(f.sub, nm, t, dt)) { bad = true; return false; } if (dt == 1) m.inits[nm] = std::move(t); else m.inits[nm]; }
            else if (f.no == 11) { std::string nm; fields(f.sub, [&](const Field& g) { if (g.no == 1 && nm.empty()) nm = str(g.sub); return true; }); g_in.push_back(nm); }
            return true; }) || bad) { err = "malformed ONNX graph"; return false; }
    for (auto& n : g_in) if (!m.inits.count(n)) m.inputs.push_back(n);
    return true;
}

inline bool read_file(const char* path, std::vector<uint8_t>& out, std::string& err) {
    gzFile f = gzopen(path, "rb");                     // transparently reads plain files too
    if (!f) { err = std::string("cannot open ") + path; return false; }
    uint8_t buf[1 << 16]; int n;
    while ((n = gzread(f, buf, sizeof buf)) > 0) out.insert(out.end(), buf, buf + n);
    gzclose(f);
    if (n < 0 || out.empty()) { err = std::string("cannot read ") + path; return false; }
    return true;
}

// ------------------------------------------------------------------------------------------ canonical tensors (weights.py)
inline bool ends_with(const std::string& s, const char* suf) { const size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }
inline bool starts_with(const std::string& s, const char* pre) { return s.compare(0, strlen(pre), pre) == 0; }
inline bool is_named(const std::string& k) { return k.find("::") == std::string::npos && !(k.size() && k[0] == '/'); }

inline bool canonical(const Model& m, Canon& W, ConvAttrs& attrs, std::string& err) {
    std::map<std::string, std::string> alias;
    for (auto& n : m.nodes) if (n.op == "Identity" && !n.in.empty() && !n.out.empty() && m.inits.count(n.in[0])) alias[n.out[0]] = n.in[0];
    auto resolve = [&](std::string name) -> const Tensor* {
        for (int i = 0; i < 8 && alias.count(name); i++) name = alias[name];
        auto it = m.inits.find(name);
        return (it == m.inits.end() || it->second.v.empty()) ? nullptr : &it->second;
    };
    for (auto& kv : m.inits) if (is_named(kv.first) && !kv.second.v.empty()) W[kv.first] = kv.second;
    for (auto& kv : alias) if (is_named(kv.first)) { const Tensor* r = resolve(kv.first); if (r) W[kv.first] = *r; }
    for (auto& n : m.nodes) {
        if ((n.op == "Conv" || n.op == "ConvTranspose") && n.in.size() >= 2) {
            const std::string& wname = n.in[1];
            std::string canon;
            if (is_named(wname)) canon = wname;
            else if (n.in.size() >= 3 && ends_with(n.in[2], ".bias")) canon = n.in[2].substr(0, n.in[2].size() - 5) + ".weight";   // quirk (i)
            if (canon.empty()) continue;
            const Tensor* w = resolve(wname);
            if (!w) { err = "Conv node " + n.name + ": weight " + wname + " is not an initializer"; return false; }
            W[canon] = *w;
            if (n.in.size() >= 3) { const Tensor* b = resolve(n.in[2]); if (b) W[n.in[2]] = *b; }
            attrs[canon] = n.ints;
        } else if (n.op == "Exp" && (n.name + "/").find("/dp/flows.0/") != std::string::npos && !n.in.empty()) {
            const Tensor* r = resolve(n.in[0]);                                            // quirk (iii): -logs
            if (r && !W.count("dp.flows.0.logs")) { Tensor t = *r; for (float& x : t.v) x = -x; W["dp.flows.0.logs"] = t; }
        }
    }
    if (W.count("dp.flows.0.m") && !W.count("dp.flows.0.logs")) { err = "cannot recover dp.flows.0.logs from the exported graph"; return false; }
    return true;
}

// indices i for which "<pre><i><suf>" is a key
inline std::vector<int> count_idx(const Canon& W, const std::string& pre, const std::string& suf) {
    std::vector<int> out;
    for (auto& kv : W) {
        const std::string& k = kv.first;
        if (k.size() <= pre.size() + suf.size() || k.compare(0, pre.size(), pre) || !ends_with(k, suf.c_str())) continue;
        const std::string mid = k.substr(pre.size(), k.size() - pre.size() - suf.size());
        if (mid.empty() || mid.find_first_not_of("0123456789") != std::string::npos) continue;
        out.push_back(atoi(mid.c_str()));
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return out;
}

inline bool infer_arch(const Canon& W, const ConvAttrs& attrs, const std::map<std::string, std::string>& meta, vits_arch& a, std::string& err) {
    memset(&a, 0, sizeof a);
    auto need = [&](const std::string& k) -> const Tensor* { auto it = W.find(k); if (it == W.end()) { err = "not a phoonnx VITS export: missing tensor " + k; return nullptr; } return &it->second; };
    auto attr0 = [&](const std::string& k, const char* an, int64_t& v) { auto it = attrs.find(k); if (it == attrs.end()) return false; auto jt = it->second.find(an); if (jt == it->second.end() || jt->second.empty()) return false; v = jt->second[0]; return true; };
    const Tensor *emb, *rel, *f1, *proj, *pre;
    if (!(emb = need("enc_p.emb.weight")) || !(rel = need("enc_p.encoder.attn_layers.0.emb_rel_k")) || !(f1 = need("enc_p.encoder.ffn_layers.0.conv_1.weight")) ||
        !(proj = need("enc_p.proj.weight")) || !(pre = need("dec.conv_pre.weight"))) return false;
    a.n_vocab = (int)emb->dim(0); a.hidden = (int)emb->dim(1);
    a.n_layers = (int)count_idx(W, "enc_p.encoder.attn_layers.", ".conv_q.weight").size();
    if (rel->dim(0) != 1) { err = "per-head relative embeddings (heads_share=False) are not supported"; return false; }
    a.window = (int)(rel->dim(1) - 1) / 2; a.n_heads = a.hidden / (int)rel->dim(2);
    a.filter = (int)f1->dim(0); a.enc_kernel = (int)f1->dim(2);
    a.inter = (int)proj->dim(0) / 2;
    if (W.count("emb_g.weight")) { a.n_speakers = (int)W.at("emb_g.weight").dim(0); a.gin = (int)W.at("emb_g.weight").dim(1); } else { a.n_speakers = 1; a.gin = 0; }
    a.use_sdp = 0;
    for (auto& kv : W) if (starts_with(kv.first, "dp.flows.")) { a.use_sdp = 1; break; }
    a.num_bins = 10;
    if (a.use_sdp) {
        const Tensor* p = need("dp.pre.weight"); const Tensor* sep = need("dp.convs.convs_sep.0.weight");
        if (!p || !sep) return false;
        a.dp_filter = (int)p->dim(0);
        a.dds_layers = (int)count_idx(W, "dp.convs.convs_sep.", ".weight").size();
        a.dp_kernel = (int)sep->dim(2);
        std::vector<int> cf = count_idx(W, "dp.flows.", ".proj.weight");
        if (cf.empty() || cf.size() > VITS_MAX_FLOWS) { err = "unsupported number of duration flows"; return false; }
        a.num_bins = ((int)W.at("dp.flows." + std::to_string(cf[0]) + ".proj.weight").dim(0) + 1) / 3;
        std::sort(cf.rbegin(), cf.rend());
        a.n_cflows = (int)cf.size();
        for (size_t i = 0; i < cf.size(); i++) a.cflows[i] = cf[i];
    } else {
        const Tensor* c1 = need("dp.conv_1.weight");
        if (!c1) return false;
        a.dp_filter = (int)c1->dim(0); a.dp_kernel = (int)c1->dim(2); a.n_cflows = 0; a.dds_layers = 3;
    }
    std::vector<int> fl = count_idx(W, "flow.flows.", ".pre.weight");
    if (fl.empty() || fl.size() > VITS_MAX_FLOWS) { err = "unsupported number of coupling layers"; return false; }
    const std::string f0 = "flow.flows." + std::to_string(fl[0]) + ".enc.in_layers.";
    a.wn_layers = (int)count_idx(W, f0, ".bias").size();
    const Tensor* wn = need(f0 + "0.weight");
    if (!wn) return false;
    a.wn_kernel = (int)wn->dim(2); a.wn_dilation_rate = 1;
    int64_t dv;
    if (a.wn_layers > 1 && attr0(f0 + "1.weight", "dilations", dv)) a.wn_dilation_rate = (int)dv;
    std::sort(fl.rbegin(), fl.rend());
    a.n_flow = (int)fl.size();
    for (size_t i = 0; i < fl.size(); i++) a.flow_layers[i] = fl[i];
    a.up_init = (int)pre->dim(0);
    std::vector<int> ups = count_idx(W, "dec.ups.", ".weight");
    if (ups.empty() || ups.size() > VITS_MAX_UPS) { err = "unsupported number of upsampling stages"; return false; }
    a.n_ups = (int)ups.size();
    for (size_t i = 0; i < ups.size(); i++) {
        const std::string k = "dec.ups." + std::to_string(ups[i]) + ".weight";
        const int kk = (int)W.at(k).dim(2);
        int64_t sv; const int s = attr0(k, "strides", sv) ? (int)sv : kk / 2;
        if (kk != 2 * s) { err = "dec.ups: only kernel == 2 * stride is supported"; return false; }
        a.up_kernels[i] = kk; a.up_rates[i] = s;
    }
    std::vector<int> rb1 = count_idx(W, "dec.resblocks.", ".convs1.0.weight"), rb2 = count_idx(W, "dec.resblocks.", ".convs.0.weight");
    a.resblock_type = rb1.empty() ? 2 : 1;
    const std::vector<int>& rbs = rb1.empty() ? rb2 : rb1;
    if (rbs.empty()) { err = "no resblocks in the decoder"; return false; }
    const int nk = (int)rbs.size() / a.n_ups;
    if (nk < 1 || nk > VITS_MAX_RBK) { err = "unsupported number of resblock kernels"; return false; }
    a.n_rbk = nk;
    const std::string key = rb1.empty() ? "convs" : "convs1";
    for (int j = 0; j < nk; j++) {
        const std::string base = "dec.resblocks." + std::to_string(j) + "." + key + ".";
        a.rb_kernels[j] = (int)W.at(base + "0.weight").dim(2);
        const int nconv = (int)count_idx(W, base, ".weight").size();
        if (nconv > VITS_MAX_DIL) { err = "too many dilations per resblock"; return false; }
        int got = 0;
        for (int c = 0; c < nconv; c++) if (attr0(base + std::to_string(c) + ".weight", "dilations", dv)) a.rb_dilations[j][got++] = (int)dv;
        if (got != nconv) {                            // no graph attributes: the reference presets (train.py:106-120)
            static const int d1[3] = {1, 3, 5};
            for (int c = 0; c < nconv; c++) {
                if (a.resblock_type == 1) a.rb_dilations[j][c] = d1[c % 3];
                else { const int kk = a.rb_kernels[j]; const int t[2] = {kk == 3 ? 1 : kk == 5 ? 2 : kk == 7 ? 3 : 1, kk == 3 ? 2 : kk == 5 ? 6 : kk == 7 ? 12 : 3}; a.rb_dilations[j][c] = t[c % 2]; }
            }
        }
        a.rb_ndil[j] = nconv;
    }
    a.sample_rate = 22050;
    auto sr = meta.find("sample_rate");
    if (sr != meta.end() && atoi(sr->second.c_str()) > 0) a.sample_rate = atoi(sr->second.c_str());
    return true;
}

// ------------------------------------------------------------------------------------------ kernel layouts (packing.py)
inline int rup(int v, int m) { return (v + m - 1) / m * m; }
inline uint16_t bf16_bits(float x) { uint32_t u; memcpy(&u, &x, 4); const uint64_t r = ((uint64_t)u + 0x7FFFu + ((u >> 16) & 1u)) >> 16; return (uint16_t)r; }
inline float bf16_round(float x) { const uint32_t u = (uint32_t)bf16_bits(x) << 16; float y; memcpy(&y, &u, 4); return y; }
inline int split3_slice(int cin) { for (int s = std::min(cin, 96); s >= 16; s--) if (cin % s == 0 && s % 16 == 0) return s; return 0; }

struct Conv3 { int taps = 0, cin = 0, n = 0; std::vector<float> w; std::vector<float> b; bool has_b = false;   // w: [taps][cin][n]
               float& at(int t, int c, int o) { return w[((size_t)t * cin + c) * n + o]; } float at(int t, int c, int o) const { return w[((size_t)t * cin + c) * n + o]; } };

inline void put_f32(std::map<std::string, Blob>& o, const std::string& name, const std::vector<float>& v) {
    Blob b; b.dtype = 0; b.bytes.resize(v.size() * 4); memcpy(b.bytes.data(), v.data(), v.size() * 4); o[name] = std::move(b);
}
// [taps][cin][n16] fp32 -> [taps][cin/8][n16][8] bf16 bits
inline void put_tc(std::map<std::string, Blob>& o, const std::string& name, const float* w, int taps, int cin, int n16) {
    Blob b; b.dtype = 1; b.bytes.resize((size_t)taps * cin * n16 * 2);
    uint16_t* d = reinterpret_cast<uint16_t*>(b.bytes.data());
    for (int t = 0; t < taps; t++) for (int c8 = 0; c8 < cin / 8; c8++) for (int n = 0; n < n16; n++) for (int e = 0; e < 8; e++)
        d[(((size_t)t * (cin / 8) + c8) * n16 + n) * 8 + e] = bf16_bits(w[((size_t)t * cin + c8 * 8 + e) * n16 + n]);
    o[name] = std::move(b);
}
inline void pack_conv(std::map<std::string, Blob>& o, const std::string& name, const Conv3& c, bool tc, bool tc3) {
    const int n4 = rup(c.n, 4), n16 = rup(c.n, 16);
    std::vector<float> w((size_t)c.taps * c.cin * n4, 0.f);
    for (int t = 0; t < c.taps; t++) for (int ci = 0; ci < c.cin; ci++) memcpy(&w[((size_t)t * c.cin + ci) * n4], &c.w[((size_t)t * c.cin + ci) * c.n], (size_t)c.n * 4);
    put_f32(o, name + ".w", w);
    if (c.has_b) { std::vector<float> b(n4, 0.f); memcpy(b.data(), c.b.data(), (size_t)c.n * 4); put_f32(o, name + ".b", b); }
    if (tc && c.cin % 16 == 0 && c.n % 16 == 0) put_tc(o, name + ".wtc", c.w.data(), c.taps, c.cin, n16);
    const int sl = split3_slice(c.cin);
    if (tc3 && c.cin % 16 == 0 && c.n % 16 == 0 && sl) {
        std::vector<float> seg((size_t)c.taps * 3 * sl * c.n);
        for (int j = 0; j < c.cin / sl; j++) {
            for (int t = 0; t < c.taps; t++) for (int k = 0; k < 3 * sl; k++) for (int n = 0; n < c.n; n++) {
                const float x = c.at(t, j * sl + (k % sl), n), hi = bf16_round(x);
                seg[((size_t)t * 3 * sl + k) * c.n + n] = (k >= sl && k < 2 * sl) ? bf16_round(x - hi) : hi;
            }
            put_tc(o, name + ".wtc3." + std::to_string(j), seg.data(), c.taps, 3 * sl, c.n);
        }
    }
}
// Conv1d weight [N, C_in, k] -> [k][C_in][N] (+ bias when present)
inline bool conv_std(const Canon& W, const std::string& name, Conv3& c, std::string& err) {
    auto it = W.find(name + ".weight");
    if (it == W.end() || it->second.dims.size() != 3) { err = "missing conv weight " + name + ".weight"; return false; }
    const Tensor& w = it->second;
    c.n = (int)w.dims[0]; c.cin = (int)w.dims[1]; c.taps = (int)w.dims[2];
    c.w.resize((size_t)c.taps * c.cin * c.n);
    for (int n = 0; n < c.n; n++) for (int ci = 0; ci < c.cin; ci++) for (int t = 0; t < c.taps; t++) c.at(t, ci, n) = w.v[((size_t)n * c.cin + ci) * c.taps + t];
    auto bt = W.find(name + ".bias");
    c.has_b = bt != W.end();
    if (c.has_b) c.b = bt->second.v;
    return true;
}

inline bool pack_model(const Canon& W, const vits_arch& a, std::map<std::string, Blob>& o, std::map<std::string, double>& opts, std::string& err) {
    const int H = a.hidden, C = a.inter;
    auto T = [&](const std::string& k) -> const Tensor* { auto it = W.find(k); if (it == W.end()) { err = "missing tensor " + k; return nullptr; } return &it->second; };
    auto copy = [&](const std::string& dst, const std::string& src) { const Tensor* t = T(src); if (!t) return false; put_f32(o, dst, t->v); return true; };
    auto std_conv = [&](const std::string& dst, const std::string& src, bool tc, bool tc3) { Conv3 c; if (!conv_std(W, src, c, err)) return false; pack_conv(o, dst, c, tc, tc3); return true; };
    if (!copy("enc.emb", "enc_p.emb.weight")) return false;
    for (int i = 0; i < a.n_layers; i++) {
        const std::string p = "enc_p.encoder.attn_layers." + std::to_string(i), e = "enc." + std::to_string(i);
        Conv3 q, k, v;
        if (!conv_std(W, p + ".conv_q", q, err) || !conv_std(W, p + ".conv_k", k, err) || !conv_std(W, p + ".conv_v", v, err)) return false;
        Conv3 qkv; qkv.taps = q.taps; qkv.cin = q.cin; qkv.n = q.n + k.n + v.n; qkv.has_b = true;
        qkv.w.resize((size_t)qkv.taps * qkv.cin * qkv.n);
        for (int t = 0; t < q.taps; t++) for (int ci = 0; ci < q.cin; ci++) {
            float* d = &qkv.w[((size_t)t * qkv.cin + ci) * qkv.n];
            memcpy(d, &q.w[((size_t)t * q.cin + ci) * q.n], (size_t)q.n * 4); memcpy(d + q.n, &k.w[((size_t)t * k.cin + ci) * k.n], (size_t)k.n * 4);
            memcpy(d + q.n + k.n, &v.w[((size_t)t * v.cin + ci) * v.n], (size_t)v.n * 4);
        }
        qkv.b = q.b; qkv.b.insert(qkv.b.end(), k.b.begin(), k.b.end()); qkv.b.insert(qkv.b.end(), v.b.begin(), v.b.end());
        pack_conv(o, e + ".qkv", qkv, false, true);
        const Tensor *rk = T(p + ".emb_rel_k"), *rv = T(p + ".emb_rel_v");
        if (!rk || !rv) return false;
        put_f32(o, e + ".rel_k", rk->v); put_f32(o, e + ".rel_v", rv->v);            // [1, 2w+1, dk]: the leading 1 drops out
        if (!std_conv(e + ".o", p + ".conv_o", false, true)) return false;
        const std::string f = "enc_p.encoder.ffn_layers." + std::to_string(i);
        if (!std_conv(e + ".ffn1", f + ".conv_1", false, true) || !std_conv(e + ".ffn2", f + ".conv_2", false, true)) return false;
        for (int j = 1; j <= 2; j++) {
            const std::string nl = "enc_p.encoder.norm_layers_" + std::to_string(j) + "." + std::to_string(i);
            if (!copy(e + ".ln" + std::to_string(j) + ".g", nl + ".gamma") || !copy(e + ".ln" + std::to_string(j) + ".b", nl + ".beta")) return false;
        }
    }
    if (!std_conv("enc.proj", "enc_p.proj", false, true)) return false;
    const Tensor* emb_g = a.n_speakers > 1 ? T("emb_g.weight") : nullptr;
    if (a.n_speakers > 1 && !emb_g) return false;
    // 1x1 conv of g = emb_g[s] -> [n_spk][N], accumulated in double like the numpy reference
    auto cond_table = [&](const std::string& name, std::vector<float>& tab, int& N) {
        const Tensor *w = T(name + ".weight"), *b = T(name + ".bias");
        if (!w || !b) return false;
        N = (int)w->dim(0); const int gin = (int)w->dim(1);
        tab.assign((size_t)a.n_speakers * N, 0.f);
        for (int s = 0; s < a.n_speakers; s++) for (int n = 0; n < N; n++) {
            double acc = 0.0;
            for (int g = 0; g < gin; g++) acc += (double)emb_g->v[(size_t)s * gin + g] * (double)w->v[(size_t)n * gin + g];
            tab[(size_t)s * N + n] = (float)(acc + (double)b->v[n]);
        }
        return true;
    };
    auto pack_dds = [&](const std::string& dst, const std::string& src) {
        for (int i = 0; i < a.dds_layers; i++) {
            const std::string si = std::to_string(i);
            const Tensor* sw = T(src + ".convs_sep." + si + ".weight");               // [C, 1, k] -> [k][C]
            if (!sw) return false;
            const int Cc = (int)sw->dim(0), k = (int)sw->dim(2);
            std::vector<float> dw((size_t)k * Cc);
            for (int c = 0; c < Cc; c++) for (int t = 0; t < k; t++) dw[(size_t)t * Cc + c] = sw->v[(size_t)c * k + t];
            put_f32(o, dst + "." + si + ".dw_w", dw);
            if (!copy(dst + "." + si + ".dw_b", src + ".convs_sep." + si + ".bias")) return false;
            if (!std_conv(dst + "." + si + ".pw", src + ".convs_1x1." + si, false, true)) return false;
            if (!copy(dst + "." + si + ".ln1.g", src + ".norms_1." + si + ".gamma") || !copy(dst + "." + si + ".ln1.b", src + ".norms_1." + si + ".beta") ||
                !copy(dst + "." + si + ".ln2.g", src + ".norms_2." + si + ".gamma") || !copy(dst + "." + si + ".ln2.b", src + ".norms_2." + si + ".beta")) return false;
        }
        return true;
    };
    if (a.use_sdp) {
        if (!std_conv("dp.pre", "dp.pre", false, true) || !std_conv("dp.proj", "dp.proj", false, true) || !pack_dds("dp.convs", "dp.convs")) return false;
        for (int q = 0; q < a.n_cflows; q++) {
            const std::string f = "dp.flows." + std::to_string(a.cflows[q]);
            if (!copy(f + ".pre_w", f + ".pre.weight") || !copy(f + ".pre_b", f + ".pre.bias") || !pack_dds(f + ".convs", f + ".convs")) return false;
            // 3 * num_bins - 1 spline parameters padded with zero output channels to a multiple of 16: the projection then runs on the
            // tensor cores (bf16x3) like the rest of the text side (packing.py)
            Conv3 pc;
            if (!conv_std(W, f + ".proj", pc, err)) return false;
            Conv3 pp; pp.taps = pc.taps; pp.cin = pc.cin; pp.n = rup(pc.n, 16); pp.has_b = true;
            pp.w.assign((size_t)pp.taps * pp.cin * pp.n, 0.f); pp.b.assign(pp.n, 0.f);
            for (int t = 0; t < pc.taps; t++) for (int ci = 0; ci < pc.cin; ci++) for (int n = 0; n < pc.n; n++) pp.at(t, ci, n) = pc.at(t, ci, n);
            if (pc.has_b) for (int n = 0; n < pc.n; n++) pp.b[n] = pc.b[n];
            pack_conv(o, f + ".proj", pp, false, true);
        }
        const Tensor *m0 = T("dp.flows.0.m"), *l0 = T("dp.flows.0.logs");
        if (!m0 || !l0) return false;
        opts["dp.ea_m"] = (double)m0->v[0]; opts["dp.ea_logs"] = (double)l0->v[0];
    } else {
        if (!std_conv("dp.conv_1", "dp.conv_1", false, true) || !std_conv("dp.conv_2", "dp.conv_2", false, true) || !std_conv("dp.proj", "dp.proj", false, false)) return false;
        for (int j = 1; j <= 2; j++)
            if (!copy("dp.norm_" + std::to_string(j) + ".g", "dp.norm_" + std::to_string(j) + ".gamma") || !copy("dp.norm_" + std::to_string(j) + ".b", "dp.norm_" + std::to_string(j) + ".beta")) return false;
    }
    if (emb_g) {
        std::vector<float> tab; int N;
        if (!cond_table("dp.cond", tab, N)) return false; put_f32(o, "dp.cond_tab", tab);
        if (!cond_table("dec.cond", tab, N)) return false; put_f32(o, "dec.cond_tab", tab);
    }
    const int half = C / 2;
    std::vector<int> inter(2 * H);                     // (tanh_c, sigmoid_c) interleaved
    for (int c = 0; c < H; c++) { inter[2 * c] = c; inter[2 * c + 1] = c + H; }
    for (int s = 0; s < a.n_flow; s++) {
        const std::string p = "flow.flows." + std::to_string(a.flow_layers[s]), d = "flow." + std::to_string(s);
        const bool flipped = (s % 2 == 0);
        Conv3 pre;
        if (!conv_std(W, p + ".pre", pre, err)) return false;                          // [1][half][H]
        if (flipped) { Conv3 t = pre; for (int ci = 0; ci < pre.cin; ci++) for (int n = 0; n < pre.n; n++) t.at(0, ci, n) = pre.at(0, pre.cin - 1 - ci, n); pre = t; }
        pack_conv(o, d + ".pre", pre, true, false);
        for (int i = 0; i < a.wn_layers; i++) {
            Conv3 in, rs;
            if (!conv_std(W, p + ".enc.in_layers." + std::to_string(i), in, err) || !conv_std(W, p + ".enc.res_skip_layers." + std::to_string(i), rs, err)) return false;
            Conv3 ip = in;
            for (int t = 0; t < in.taps; t++) for (int ci = 0; ci < in.cin; ci++) for (int n = 0; n < 2 * H; n++) ip.at(t, ci, n) = in.at(t, ci, inter[n]);
            for (int n = 0; n < 2 * H; n++) ip.b[n] = in.b[inter[n]];
            pack_conv(o, d + ".in." + std::to_string(i), ip, true, false);
            pack_conv(o, d + ".rs." + std::to_string(i), rs, true, false);
        }
        if (emb_g) {
            std::vector<float> tab; int N;
            if (!cond_table(p + ".enc.cond_layer", tab, N)) return false;               // [n_spk][2H * layers]
            for (int i = 0; i < a.wn_layers; i++) {
                std::vector<float> t((size_t)a.n_speakers * 2 * H);
                for (int sp = 0; sp < a.n_speakers; sp++) for (int n = 0; n < 2 * H; n++) t[(size_t)sp * 2 * H + n] = tab[(size_t)sp * N + i * 2 * H + inter[n]];
                put_f32(o, d + ".cond_tab." + std::to_string(i), t);
            }
        }
        Conv3 post;
        if (!conv_std(W, p + ".post", post, err)) return false;                        // [1][H][half]
        if (flipped) { Conv3 t = post; for (int ci = 0; ci < H; ci++) for (int n = 0; n < half; n++) t.at(0, ci, n) = post.at(0, ci, half - 1 - n); for (int n = 0; n < half; n++) t.b[n] = post.b[half - 1 - n]; post = t; }
        pack_conv(o, d + ".post", post, true, false);
        // tensor-core form of the WN tail: m = post(sum_l skip_l(acts_l)) as ONE composed GEMM (packing.py), in double
        std::vector<double> bsum(H, 0.0);
        for (int i = 0; i < a.wn_layers; i++) {
            Conv3 rs;
            if (!conv_std(W, p + ".enc.res_skip_layers." + std::to_string(i), rs, err)) return false;   // [1][H][2H] (last: [1][H][H])
            const bool last = (i == a.wn_layers - 1);
            const int off = last ? 0 : H;
            if (!last) {
                Conv3 r; r.taps = 1; r.cin = H; r.n = H; r.has_b = true; r.w.resize((size_t)H * H); r.b.assign(rs.b.begin(), rs.b.begin() + H);
                for (int ci = 0; ci < H; ci++) for (int n = 0; n < H; n++) r.at(0, ci, n) = rs.at(0, ci, n);
                pack_conv(o, d + ".rsr." + std::to_string(i), r, true, false);
            }
            Conv3 mk; mk.taps = 1; mk.cin = H; mk.n = half; mk.has_b = false; mk.w.resize((size_t)H * half);
            for (int ci = 0; ci < H; ci++) for (int n = 0; n < half; n++) {
                double acc = 0.0;
                for (int k = 0; k < H; k++) acc += (double)rs.at(0, ci, off + k) * (double)post.at(0, k, n);
                mk.at(0, ci, n) = (float)acc;
            }
            pack_conv(o, d + ".mskip." + std::to_string(i), mk, true, false);
            for (int k = 0; k < H; k++) bsum[k] += (double)rs.b[off + k];
        }
        std::vector<float> mb(half);
        for (int n = 0; n < half; n++) { double acc = 0.0; for (int k = 0; k < H; k++) acc += bsum[k] * (double)post.at(0, k, n); mb[n] = (float)(acc + (double)post.b[n]); }
        put_f32(o, d + ".mskip.b", mb);
    }
    if (!std_conv("dec.pre", "dec.conv_pre", true, false)) return false;
    for (int i = 0; i < a.n_ups; i++) {
        const int u = a.up_rates[i], k = a.up_kernels[i];
        const Tensor *w = T("dec.ups." + std::to_string(i) + ".weight"), *b = T("dec.ups." + std::to_string(i) + ".bias");   // [C_in, C_out, k]
        if (!w || !b) return false;
        const int cin = (int)w->dim(0), cout = (int)w->dim(1), pad = (k - u) / 2, hu = u / 2;
        if (k != 2 * u || u % 2 || pad != u / 2) { err = "upsample stage: kernel / stride unsupported"; return false; }
        Conv3 wa, wb; wa.taps = wb.taps = 2; wa.cin = wb.cin = cin; wa.n = wb.n = hu * cout; wa.has_b = wb.has_b = true;
        wa.w.assign((size_t)2 * cin * hu * cout, 0.f); wb.w = wa.w;
        auto wk = [&](int ci, int co, int t) { return w->v[((size_t)ci * cout + co) * k + t]; };
        for (int p_ = 0; p_ < hu; p_++) {
            const int r = p_ + pad, r2 = p_ + hu + pad;
            for (int ci = 0; ci < cin; ci++) for (int co = 0; co < cout; co++) {
                wa.at(0, ci, p_ * cout + co) = wk(ci, co, r + u); wa.at(1, ci, p_ * cout + co) = wk(ci, co, r);
                wb.at(0, ci, p_ * cout + co) = wk(ci, co, r2);    wb.at(1, ci, p_ * cout + co) = wk(ci, co, r2 - u);
            }
        }
        wa.b.resize((size_t)hu * cout);
        for (int p_ = 0; p_ < hu; p_++) memcpy(&wa.b[(size_t)p_ * cout], b->v.data(), (size_t)cout * 4);
        wb.b = wa.b;
        pack_conv(o, "dec.ups." + std::to_string(i) + ".A", wa, true, false);
        pack_conv(o, "dec.ups." + std::to_string(i) + ".B", wb, true, false);
        for (int j = 0; j < a.n_rbk; j++) {
            const int n = i * a.n_rbk + j;
            for (int c = 0; c < a.rb_ndil[j]; c++) {
                const std::string r = "dec.resblocks." + std::to_string(n), dd = "dec.rb." + std::to_string(n), sc = std::to_string(c);
                if (a.resblock_type == 1) { if (!std_conv(dd + ".c1." + sc, r + ".convs1." + sc, true, false) || !std_conv(dd + ".c2." + sc, r + ".convs2." + sc, true, false)) return false; }
                else if (!std_conv(dd + ".c." + sc, r + ".convs." + sc, true, false)) return false;
            }
        }
    }
    const Tensor* pw = T("dec.conv_post.weight");     // [1, C, 7] -> [7][C]
    if (!pw) return false;
    const int pc = (int)pw->dim(1), pk = (int)pw->dim(2);
    std::vector<float> post_w((size_t)pk * pc);
    for (int c = 0; c < pc; c++) for (int t = 0; t < pk; t++) post_w[(size_t)t * pc + c] = pw->v[(size_t)c * pk + t];
    put_f32(o, "dec.post_w", post_w);
    return true;
}

// the whole chain for one file
struct Voice { Model model; Canon W; ConvAttrs attrs; vits_arch arch; std::map<std::string, Blob> blobs; std::map<std::string, double> opts; };
inline bool load_voice(const char* path, Voice& v, std::string& err, bool pack = true) {
    std::vector<uint8_t> data;
    if (!read_file(path, data, err) || !parse_model(data, v.model, err) || !canonical(v.model, v.W, v.attrs, err)) return false;
    data.clear(); data.shrink_to_fit();
    if (!infer_arch(v.W, v.attrs, v.model.meta, v.arch, err)) return false;
    const bool has_sid = std::find(v.model.inputs.begin(), v.model.inputs.end(), "sid") != v.model.inputs.end();
    if (has_sid != (v.arch.n_speakers > 1)) { err = "graph inputs and emb_g disagree about multi-speaker support"; return false; }
    for (auto& kv : v.W) if (starts_with(kv.first, "emb_l.") || starts_with(kv.first, "emb_lang.")) { err = "multi-lingual voices (language embedding) are not supported"; return false; }
    return !pack || pack_model(v.W, v.arch, v.blobs, v.opts, err);
}

}  // namespace vf
