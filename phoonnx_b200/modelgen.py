"""Synthetic (random-init) Piper-style VITS voices in the exporter's file format.

There are no pretrained voices offline and the reference checkout does not travel to the
GPU box, so benchmarks and GPU tests need a way to mint a model of a named architecture
(BASELINE.json configs) without importing phoonnx_train.  ``synth_weights`` draws every
tensor of the canonical state_dict layout with the reference's init distributions
(models.py:190-191, attentions.py:197-209, commons.py:11-14, torch Conv1d default) plus the
"de-zero" recipe of SURVEY.md 8(c) (flows/splines would otherwise be identities), and
``write_onnx`` serialises it the way ``phoonnx_train/export_onnx.py:318-350`` does --
including the quirks the loader must undo (anonymous folded flow weights, ``-logs`` folded
into an ``Exp`` input, optional ``Identity`` de-duplication, metadata_props).
"""
from __future__ import annotations

import json
import math
from typing import Dict, List, Optional

import numpy as np

from . import onnx_reader as pb
from .weights import VitsArch

PRESETS = {
    # train.py:106-120, lightning.py:26-57
    "x_low": dict(hidden=96, inter=96, filter=384, dp_filter=96),
    "low": dict(sample_rate=16000),
    "medium": dict(),
    "high": dict(resblock="1", rb_kernels=(3, 7, 11), rb_dilations=((1, 3, 5),) * 3,
                 up_rates=(8, 8, 2, 2), up_init=512, up_kernels=(16, 16, 4, 4)),
    "tiny": dict(hidden=32, inter=32, filter=64, dp_filter=32, n_layers=2, up_init=32,
                 up_rates=(4, 4, 2), up_kernels=(8, 8, 4), rb_kernels=(3, 5),
                 rb_dilations=((1, 2), (2, 6))),
    "tiny_rb1": dict(hidden=32, inter=32, filter=64, dp_filter=32, n_layers=2, resblock="1",
                     up_init=32, up_rates=(4, 2, 2), up_kernels=(8, 4, 4), rb_kernels=(3, 7),
                     rb_dilations=((1, 3, 5), (1, 3, 5))),
}


def make_arch(preset: str = "medium", n_speakers: int = 1, n_vocab: int = 256,
              sample_rate: Optional[int] = None, use_sdp: bool = True) -> VitsArch:
    kw = dict(PRESETS[preset])
    a = VitsArch(**kw)
    a.n_vocab = n_vocab
    a.n_speakers = n_speakers
    a.gin = 0 if n_speakers <= 1 else (16 if preset.startswith("tiny") else 512)
    a.use_sdp = use_sdp
    if not use_sdp:
        a.dp_filter, a.cflows = 256, ()
    if preset == "x_low" and sample_rate is None:
        sample_rate = 16000
    if sample_rate:
        a.sample_rate = sample_rate
    return a


def synth_weights(a: VitsArch, seed: int = 1234, audio_gain: float = 16.0) -> Dict[str, np.ndarray]:
    rs = np.random.RandomState(seed)
    W: Dict[str, np.ndarray] = {}

    def normal(shape, std, mean=0.0):
        return (rs.standard_normal(shape) * std + mean).astype(np.float32)

    def uniform(shape, bound):
        return rs.uniform(-bound, bound, size=shape).astype(np.float32)

    def conv(name, cout, cin, k, groups=1, bias=True, std=None):
        fan_in = (cin // groups) * k
        b = 1.0 / math.sqrt(fan_in)
        W[name + ".weight"] = normal((cout, cin // groups, k), std) if std else uniform((cout, cin // groups, k), b)
        if bias:
            W[name + ".bias"] = uniform((cout,), b)

    def ln(name, c):
        W[name + ".gamma"] = normal((c,), 0.1, 1.0)
        W[name + ".beta"] = normal((c,), 0.1)

    H, C, Fc, dk = a.hidden, a.inter, a.filter, a.k_channels
    W["enc_p.emb.weight"] = normal((a.n_vocab, H), H ** -0.5)
    for i in range(a.n_layers):
        p = f"enc_p.encoder.attn_layers.{i}"
        W[p + ".emb_rel_k"] = normal((1, 2 * a.window + 1, dk), dk ** -0.5)
        W[p + ".emb_rel_v"] = normal((1, 2 * a.window + 1, dk), dk ** -0.5)
        xav = math.sqrt(6.0 / (H + H))
        for nm in ("conv_q", "conv_k", "conv_v"):
            W[f"{p}.{nm}.weight"] = uniform((H, H, 1), xav)
            W[f"{p}.{nm}.bias"] = uniform((H,), 1.0 / math.sqrt(H))
        conv(p + ".conv_o", H, H, 1)
        ln(f"enc_p.encoder.norm_layers_1.{i}", H)
        conv(f"enc_p.encoder.ffn_layers.{i}.conv_1", Fc, H, a.enc_kernel)
        conv(f"enc_p.encoder.ffn_layers.{i}.conv_2", H, Fc, a.enc_kernel)
        ln(f"enc_p.encoder.norm_layers_2.{i}", H)
    conv("enc_p.proj", 2 * C, H, 1)
    if a.n_speakers > 1:
        W["emb_g.weight"] = normal((a.n_speakers, a.gin), 1.0)

    def dds(prefix, ch):
        for i in range(a.dds_layers):
            conv(f"{prefix}.convs_sep.{i}", ch, ch, a.dp_kernel, groups=ch)
            conv(f"{prefix}.convs_1x1.{i}", ch, ch, 1)
            ln(f"{prefix}.norms_1.{i}", ch)
            ln(f"{prefix}.norms_2.{i}", ch)

    Fd = a.dp_filter
    if a.use_sdp:
        conv("dp.pre", Fd, H, 1)
        conv("dp.proj", Fd, Fd, 1)
        dds("dp.convs", Fd)
        if a.gin:
            conv("dp.cond", Fd, a.gin, 1)
        W["dp.flows.0.m"] = normal((2, 1), 0.3)
        W["dp.flows.0.logs"] = normal((2, 1), 0.3)
        W["dp.flows.0.m"][0] = -2.0          # realistic durations (mean ~3.5 frames/id)
        W["dp.flows.0.logs"][0] = 0.7
        for fi in a.cflows:
            conv(f"dp.flows.{fi}.pre", Fd, 1, 1)
            dds(f"dp.flows.{fi}.convs", Fd)
            nproj = 3 * a.num_bins - 1
            W[f"dp.flows.{fi}.proj.weight"] = normal((nproj, Fd, 1), 0.05)
            W[f"dp.flows.{fi}.proj.bias"] = normal((nproj,), 0.05)
    else:
        conv("dp.conv_1", Fd, H, a.dp_kernel)
        ln("dp.norm_1", Fd)
        conv("dp.conv_2", Fd, Fd, a.dp_kernel)
        ln("dp.norm_2", Fd)
        conv("dp.proj", 1, Fd, 1)
        W["dp.proj.bias"] = np.asarray([1.1], np.float32)   # exp(1.1) ~ 3 frames/id
        if a.gin:
            conv("dp.cond", H, a.gin, 1)
    for fi in a.flow_layers:
        p = f"flow.flows.{fi}"
        conv(p + ".pre", H, C // 2, 1)
        for i in range(a.wn_layers):
            conv(f"{p}.enc.in_layers.{i}", 2 * H, H, a.wn_kernel)
            conv(f"{p}.enc.res_skip_layers.{i}", 2 * H if i < a.wn_layers - 1 else H, H, 1)
        if a.gin:
            conv(p + ".enc.cond_layer", 2 * H * a.wn_layers, a.gin, 1)
        W[p + ".post.weight"] = normal((C // 2, H, 1), 0.05)
        W[p + ".post.bias"] = normal((C // 2,), 0.05)
    conv("dec.conv_pre", a.up_init, C, 7)
    ch = a.up_init
    nk = len(a.rb_kernels)
    for i, (u, k) in enumerate(zip(a.up_rates, a.up_kernels)):
        W[f"dec.ups.{i}.weight"] = normal((ch, ch // 2, k), 0.01)      # ConvTranspose layout [Cin, Cout, k]
        W[f"dec.ups.{i}.bias"] = uniform((ch // 2,), 1.0 / math.sqrt((ch // 2) * k))
        ch //= 2
        for j, (kk, dil) in enumerate(zip(a.rb_kernels, a.rb_dilations)):
            n = i * nk + j
            for c in range(len(dil)):
                if a.resblock == "1":
                    conv(f"dec.resblocks.{n}.convs1.{c}", ch, ch, kk, std=0.01)
                    conv(f"dec.resblocks.{n}.convs2.{c}", ch, ch, kk, std=0.01)
                else:
                    conv(f"dec.resblocks.{n}.convs.{c}", ch, ch, kk, std=0.01)
    conv("dec.conv_post", 1, ch, 7, bias=False)
    W["dec.conv_post.weight"] *= np.float32(audio_gain)   # random-init peak ~0.02 -> ~0.3
    if a.gin:
        conv("dec.cond", a.up_init, a.gin, 1)
    return W


def conv_attr_table(a: VitsArch) -> Dict[str, dict]:
    """dilations/pads/strides of every Conv the exporter would emit, keyed by weight name."""
    t: Dict[str, dict] = {}

    def c(name, k, d=1, group=1):
        pad = (k * d - d) // 2
        t[name + ".weight"] = {"dilations": [d], "group": group, "kernel_shape": [k], "pads": [pad, pad], "strides": [1]}

    for i in range(a.n_layers):
        p = f"enc_p.encoder.attn_layers.{i}"
        for nm in ("conv_q", "conv_k", "conv_v", "conv_o"):
            c(f"{p}.{nm}", 1)
        for nm in ("conv_1", "conv_2"):
            t[f"enc_p.encoder.ffn_layers.{i}.{nm}.weight"] = {"dilations": [1], "group": 1, "kernel_shape": [a.enc_kernel], "pads": [0, 0], "strides": [1]}
    c("enc_p.proj", 1)

    def dds(prefix, ch):
        for i in range(a.dds_layers):
            c(f"{prefix}.convs_sep.{i}", a.dp_kernel, a.dp_kernel ** i, ch)
            c(f"{prefix}.convs_1x1.{i}", 1)

    if a.use_sdp:
        c("dp.pre", 1), c("dp.proj", 1), dds("dp.convs", a.dp_filter)
        for fi in a.cflows:
            c(f"dp.flows.{fi}.pre", 1), dds(f"dp.flows.{fi}.convs", a.dp_filter), c(f"dp.flows.{fi}.proj", 1)
    else:
        c("dp.conv_1", a.dp_kernel), c("dp.conv_2", a.dp_kernel), c("dp.proj", 1)
    if a.gin:
        c("dp.cond", 1), c("dec.cond", 1)
    for fi in a.flow_layers:
        p = f"flow.flows.{fi}"
        c(p + ".pre", 1), c(p + ".post", 1)
        for i in range(a.wn_layers):
            c(f"{p}.enc.in_layers.{i}", a.wn_kernel, a.wn_dilation_rate ** i)
            c(f"{p}.enc.res_skip_layers.{i}", 1)
        if a.gin:
            c(p + ".enc.cond_layer", 1)
    c("dec.conv_pre", 7), c("dec.conv_post", 7)
    nk = len(a.rb_kernels)
    for i, (u, k) in enumerate(zip(a.up_rates, a.up_kernels)):
        t[f"dec.ups.{i}.weight"] = {"dilations": [1], "group": 1, "kernel_shape": [k], "pads": [(k - u) // 2] * 2, "strides": [u]}
        for j, (kk, dil) in enumerate(zip(a.rb_kernels, a.rb_dilations)):
            n = i * nk + j
            for ci, d in enumerate(dil):
                if a.resblock == "1":
                    c(f"dec.resblocks.{n}.convs1.{ci}", kk, d), c(f"dec.resblocks.{n}.convs2.{ci}", kk, 1)
                else:
                    c(f"dec.resblocks.{n}.convs.{ci}", kk, d)
    return t


def write_onnx(W: Dict[str, np.ndarray], a: VitsArch, path: str, dedup_identity: bool = True,
               phoneme_id_map: Optional[dict] = None, graph_inputs: Optional[List[str]] = None) -> None:
    """Serialise canonical tensors the way export_onnx.py's torch.onnx.export does."""
    attrs = conv_attr_table(a)
    inits: Dict[str, np.ndarray] = {}
    nodes: List[bytes] = []
    anon = [9000]

    def anon_name(op):
        anon[0] += 7
        return f"onnx::{op}_{anon[0]}"

    seen_bytes: Dict[bytes, str] = {}

    def add_named(name, arr):
        """Named initializer, or an Identity alias of an earlier byte-identical one (quirk ii)."""
        key = arr.tobytes() + str(arr.shape).encode()
        if dedup_identity and key in seen_bytes and arr.size > 1:
            nodes.append(pb.encode_node("Identity", f"Identity_{len(nodes)}", [seen_bytes[key]], [name]))
            return
        seen_bytes.setdefault(key, name)
        inits[name] = arr

    for name in sorted(W):
        arr = np.ascontiguousarray(W[name], dtype=np.float32)
        is_flow_wn = name.startswith("flow.flows.") and ".enc." in name and name.endswith(".weight")
        if is_flow_wn:
            continue                      # emitted below as anonymous (quirk i)
        if name == "dp.flows.0.logs":
            continue                      # emitted below as -logs feeding Exp (quirk iii)
        if name.endswith(".weight") and name in attrs:
            continue                      # emitted with its Conv node (keeps exporter order irrelevant)
        if name.endswith(".bias") and (name[:-5] + ".weight") in attrs:
            continue
        add_named(name, arr)
    for wname, at in attrs.items():
        base = wname[: -len(".weight")]
        w = np.ascontiguousarray(W[wname], dtype=np.float32)
        is_flow_wn = wname.startswith("flow.flows.") and ".enc." in wname
        if is_flow_wn:
            w_in = anon_name("Conv")
            inits[w_in] = w
        else:
            w_in = wname
            add_named(wname, w)
        ins = [f"/{base.replace('.', '/')}/in", w_in]
        if base + ".bias" in W:
            add_named(base + ".bias", np.ascontiguousarray(W[base + ".bias"], dtype=np.float32))
            ins.append(base + ".bias")
        op = "ConvTranspose" if base.startswith("dec.ups.") else "Conv"
        nodes.append(pb.encode_node(op, f"/{base.replace('.', '/')}/{op}", ins, [f"/{base.replace('.', '/')}/out"], at))
    if "dp.flows.0.logs" in W:
        nm = anon_name("Exp")
        inits[nm] = (-np.asarray(W["dp.flows.0.logs"], dtype=np.float32))
        nodes.append(pb.encode_node("Exp", "/dp/flows.0/Exp", [nm], ["/dp/flows.0/Exp_output_0"]))
    # `graph_inputs`: other exporters of the same model declare other input lists (no `scales`: voice.py:358; a `langid`: voice.py:369)
    inputs = list(graph_inputs) if graph_inputs else ["input", "input_lengths", "scales"] + (["sid"] if a.n_speakers > 1 else [])
    meta = {   # export_onnx.py:335-345
        "model_type": "vits", "n_speakers": a.n_speakers, "n_vocab": a.n_vocab,
        "sample_rate": a.sample_rate, "alphabet": "ipa", "phoneme_type": "raw",
        "phonemizer_model": "", "phoneme_id_map": json.dumps(phoneme_id_map or {}), "has_espeak": False,
    }
    data = pb.encode_model(inits, nodes, inputs, ["output"], meta)
    with open(path, "wb") as f:
        f.write(data)


def make_voice(path: str, preset: str = "medium", n_speakers: int = 1, seed: int = 1234,
               sample_rate: Optional[int] = None, use_sdp: bool = True, audio_gain: float = 16.0):
    """Write ``path`` (.onnx) + ``path.json`` (voice config the reference's VoiceConfig reads)."""
    a = make_arch(preset, n_speakers, sample_rate=sample_rate, use_sdp=use_sdp)
    W = synth_weights(a, seed, audio_gain)
    write_onnx(W, a, path)
    cfg = {
        "audio": {"sample_rate": a.sample_rate}, "lang_code": "en", "phoneme_type": "raw",
        "num_symbols": a.n_vocab, "num_speakers": a.n_speakers,
        "inference": {"noise_scale": 0.667, "length_scale": 1.0, "noise_w": 0.8},
        "phoneme_id_map": {"_": [0], "^": [1], "$": [2], " ": [3]},
    }
    with open(str(path) + ".json", "w") as f:
        json.dump(cfg, f)
    return W, a
