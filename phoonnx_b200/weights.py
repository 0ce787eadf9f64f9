"""Weight loader: exported model (or checkpoint state_dict) -> canonical fp32 tensors + arch.

Reads the initializers of a model written by ``phoonnx_train/export_onnx.py:318-350``
(SURVEY.md 8a-W), resolving the exporter's quirks:

  (i)   the flow's WaveNet convs keep weight-norm at export (export_onnx.py:244 only strips
        ``dec``), so their weights are constant-folded into anonymous ``onnx::Conv_N``
        initializers -> recovered through the Conv node whose *bias* input carries the
        state_dict name (``flow.flows.K.enc.in_layers.i.bias`` ...);
  (ii)  byte-identical initializers are de-duplicated and aliased through ``Identity`` nodes;
  (iii) ``dp.flows.0.logs`` only survives as the folded constant ``-logs`` feeding ``Exp``
        (modules.py:408);
  (iv)  no hyper-parameters are stored: the architecture is inferred from tensor shapes and
        from Conv node attributes (dilations), SURVEY.md Appendix D.

The canonical form is ``{state_dict key (weight-norm folded): float32 ndarray}`` plus a
``VitsArch``.  The same canonical form is produced from a Lightning checkpoint /
``state_dict`` (``model_g.`` prefix, ``weight_g``/``weight_v`` pairs, lightning.py:86).
"""
from __future__ import annotations

import re
from dataclasses import asdict, dataclass, field
from typing import Dict, List, Mapping, Optional, Tuple

import numpy as np

from .onnx_reader import OnnxModel, read_onnx


@dataclass
class VitsArch:
    n_vocab: int = 256
    hidden: int = 192            # H  (enc_p.emb.weight.shape[1], models.py:190)
    inter: int = 192             # inter_channels (enc_p.proj out / 2, models.py:196)
    filter: int = 768            # FFN filter channels (attentions.py:382)
    n_heads: int = 2
    n_layers: int = 6
    enc_kernel: int = 3
    window: int = 4              # relative attention window (attentions.py:21)
    n_speakers: int = 1
    gin: int = 0
    use_sdp: bool = True
    dp_filter: int = 192         # SDP filter == hidden (models.py:25); DP filter 256
    dp_kernel: int = 3
    dds_layers: int = 3
    cflows: Tuple[int, ...] = (7, 5, 3)      # ConvFlows used in reverse, in execution order (models.py:109-110)
    num_bins: int = 10
    flow_layers: Tuple[int, ...] = (6, 4, 2, 0)  # coupling layers in reverse execution order
    wn_layers: int = 4
    wn_kernel: int = 5
    wn_dilation_rate: int = 1
    resblock: str = "2"
    up_rates: Tuple[int, ...] = (8, 8, 4)
    up_kernels: Tuple[int, ...] = (16, 16, 8)
    up_init: int = 256
    rb_kernels: Tuple[int, ...] = (3, 5, 7)
    rb_dilations: Tuple[Tuple[int, ...], ...] = ((1, 2), (2, 6), (3, 12))
    sample_rate: int = 22050

    @property
    def hop(self) -> int:
        h = 1
        for u in self.up_rates:
            h *= u
        return h

    @property
    def k_channels(self) -> int:
        return self.hidden // self.n_heads

    def to_dict(self):
        return asdict(self)

    # exact closed-form MAC counts (SURVEY.md section 8 preset table) -----------------
    def dec_mac_per_frame(self) -> int:
        mac = 7 * self.inter * self.up_init  # conv_pre at frame rate
        ch, rate = self.up_init, 1
        for u, k in zip(self.up_rates, self.up_kernels):
            # ConvTranspose: each input sample contributes k taps to C_out outputs
            mac += rate * ch * (ch // 2) * k
            ch //= 2
            rate *= u
            for kk, dil in zip(self.rb_kernels, self.rb_dilations):
                n_convs = len(dil) * (2 if self.resblock == "1" else 1)
                mac += rate * n_convs * ch * ch * kk
        mac += rate * ch * 7
        return mac

    def flow_mac_per_frame(self) -> int:
        H, C = self.hidden, self.inter
        per = (C // 2) * H + H * (C // 2)
        for i in range(self.wn_layers):
            per += H * 2 * H * self.wn_kernel
            per += H * (2 * H if i < self.wn_layers - 1 else H)
        return per * len(self.flow_layers)

    def enc_mac_per_id(self) -> int:
        H, F, k = self.hidden, self.filter, self.enc_kernel
        return self.n_layers * (4 * H * H + 2 * H * F * k) + H * 2 * self.inter

    def dp_mac_per_id(self) -> int:
        if not self.use_sdp:
            Fd = self.dp_filter
            return self.hidden * Fd * 3 + Fd * Fd * 3 + Fd
        Fd = self.dp_filter
        dds = self.dds_layers * (Fd * self.dp_kernel + Fd * Fd)
        return self.hidden * Fd + dds + Fd * Fd + len(self.cflows) * (Fd + dds + Fd * (3 * self.num_bins - 1))


# ----------------------------------------------------------------------------- canonicalisation


def _fold_weight_norm(g: np.ndarray, v: np.ndarray) -> np.ndarray:
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v||_2 over dims (1,2)."""
    norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
    return (g.astype(np.float64) * v.astype(np.float64) / norm).astype(np.float32)


def canonical_from_state_dict(sd: Mapping[str, "np.ndarray"]) -> Dict[str, np.ndarray]:
    """Lightning checkpoint ``state_dict`` (``model_g.*``) or a plain SynthesizerTrn
    state_dict -> canonical tensors (weight-norm folded, training-only modules dropped)."""
    out: Dict[str, np.ndarray] = {}
    items = {}
    for k, v in sd.items():
        if k.startswith("model_d."):
            continue
        if k.startswith("model_g."):
            k = k[len("model_g."):]
        arr = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        items[k] = np.ascontiguousarray(arr, dtype=np.float32)
    for k, arr in items.items():
        if k.startswith("enc_q.") or k.startswith("dp.post_") or k.startswith("dp.flows.1."):
            continue  # unused in reverse (models.py:109-110)
        if k.endswith(".weight_g"):
            base = k[: -len("_g")]
            out[base] = _fold_weight_norm(arr, items[base + "_v"])
        elif k.endswith(".weight_v"):
            continue
        elif k.endswith("parametrizations.weight.original0"):
            base = k[: -len(".parametrizations.weight.original0")] + ".weight"
            out[base] = _fold_weight_norm(arr, items[k[:-1] + "1"])
        elif k.endswith("parametrizations.weight.original1"):
            continue
        else:
            out[k] = arr
    return out


def canonical_from_onnx(model: OnnxModel) -> Tuple[Dict[str, np.ndarray], Dict[str, dict]]:
    """Exported graph -> (canonical tensors, conv attributes keyed by canonical weight name)."""
    inits = model.initializers
    alias: Dict[str, str] = {}
    for n in model.nodes:
        if n.op_type == "Identity" and n.inputs and n.inputs[0] in inits:
            alias[n.outputs[0]] = n.inputs[0]

    def resolve(name: str) -> Optional[np.ndarray]:
        seen = 0
        while name in alias and seen < 8:
            name = alias[name]
            seen += 1
        return inits.get(name)

    out: Dict[str, np.ndarray] = {}
    for k, v in inits.items():
        if "::" in k or k.startswith("/") or v.dtype != np.float32:
            continue
        out[k] = v
    for k in alias:
        if "::" not in k and not k.startswith("/"):
            r = resolve(k)
            if r is not None and r.dtype == np.float32:
                out[k] = r
    conv_attrs: Dict[str, dict] = {}
    for n in model.nodes:
        if n.op_type in ("Conv", "ConvTranspose") and len(n.inputs) >= 2:
            wname = n.inputs[1]
            canon = None
            if "::" not in wname and not wname.startswith("/"):
                canon = wname
            elif len(n.inputs) >= 3 and n.inputs[2].endswith(".bias"):
                canon = n.inputs[2][: -len(".bias")] + ".weight"   # quirk (i)
            if canon is None:
                continue
            w = resolve(wname)
            if w is None:
                raise ValueError(f"Conv node {n.name}: weight {wname!r} is not an initializer")
            out[canon] = w
            if len(n.inputs) >= 3:
                b = resolve(n.inputs[2])
                if b is not None:
                    out[n.inputs[2]] = b
            conv_attrs[canon] = dict(n.attrs)
        elif n.op_type == "Exp" and "/dp/flows.0/" in n.name + "/" and n.inputs:
            r = resolve(n.inputs[0])                              # quirk (iii): -logs
            if r is not None and "dp.flows.0.logs" not in out:
                out["dp.flows.0.logs"] = (-r).astype(np.float32)
    if "dp.flows.0.m" in out and "dp.flows.0.logs" not in out:
        # fall back: some exports keep Neg as a node; then the named initializer exists
        # (handled above).  Otherwise the export is not one we understand.
        raise ValueError("cannot recover dp.flows.0.logs from the exported graph")
    return out, conv_attrs


# ----------------------------------------------------------------------------- arch inference


def _count(W: Mapping[str, np.ndarray], pattern: str) -> List[int]:
    rx = re.compile(pattern)
    idx = set()
    for k in W:
        m = rx.fullmatch(k)
        if m:
            idx.add(int(m.group(1)))
    return sorted(idx)


def infer_arch(W: Mapping[str, np.ndarray], conv_attrs: Optional[Mapping[str, dict]] = None,
               metadata: Optional[Mapping[str, str]] = None,
               sample_rate: Optional[int] = None) -> VitsArch:
    """SURVEY.md Appendix D rules."""
    conv_attrs = conv_attrs or {}
    a = VitsArch()
    try:
        a.n_vocab, a.hidden = (int(x) for x in W["enc_p.emb.weight"].shape)
        a.n_layers = len(_count(W, r"enc_p\.encoder\.attn_layers\.(\d+)\.conv_q\.weight"))
        rel = W["enc_p.encoder.attn_layers.0.emb_rel_k"]
        if rel.shape[0] != 1:
            raise ValueError("per-head relative embeddings (heads_share=False) are not supported")
        a.window = (rel.shape[1] - 1) // 2
        a.n_heads = a.hidden // rel.shape[2]
        f = W["enc_p.encoder.ffn_layers.0.conv_1.weight"]
        a.filter, a.enc_kernel = int(f.shape[0]), int(f.shape[2])
        a.inter = W["enc_p.proj.weight"].shape[0] // 2
    except KeyError as e:
        raise ValueError(f"not a phoonnx VITS export: missing tensor {e}") from None
    if "emb_g.weight" in W:
        a.n_speakers, a.gin = (int(x) for x in W["emb_g.weight"].shape)
    else:
        a.n_speakers, a.gin = 1, 0
    a.use_sdp = any(k.startswith("dp.flows.") for k in W)
    if a.use_sdp:
        a.dp_filter = int(W["dp.pre.weight"].shape[0])
        seps = _count(W, r"dp\.convs\.convs_sep\.(\d+)\.weight")
        a.dds_layers = len(seps)
        a.dp_kernel = int(W["dp.convs.convs_sep.0.weight"].shape[2])
        cf = _count(W, r"dp\.flows\.(\d+)\.proj\.weight")
        a.cflows = tuple(sorted(cf, reverse=True))
        a.num_bins = (int(W[f"dp.flows.{cf[0]}.proj.weight"].shape[0]) + 1) // 3
    else:
        a.dp_filter = int(W["dp.conv_1.weight"].shape[0])
        a.dp_kernel = int(W["dp.conv_1.weight"].shape[2])
        a.cflows = ()
    fl = _count(W, r"flow\.flows\.(\d+)\.pre\.weight")
    a.flow_layers = tuple(sorted(fl, reverse=True))
    a.wn_layers = len(_count(W, rf"flow\.flows\.{fl[0]}\.enc\.in_layers\.(\d+)\.bias"))
    wn_w = W[f"flow.flows.{fl[0]}.enc.in_layers.0.weight"]
    a.wn_kernel = int(wn_w.shape[2])
    a.wn_dilation_rate = 1
    if a.wn_layers > 1:
        at = conv_attrs.get(f"flow.flows.{fl[0]}.enc.in_layers.1.weight")
        if at and "dilations" in at:
            a.wn_dilation_rate = int(at["dilations"][0])
    a.up_init = int(W["dec.conv_pre.weight"].shape[0])
    ups = _count(W, r"dec\.ups\.(\d+)\.weight")
    kernels, rates = [], []
    for i in ups:
        w = W[f"dec.ups.{i}.weight"]
        k = int(w.shape[2])
        at = conv_attrs.get(f"dec.ups.{i}.weight")
        s = int(at["strides"][0]) if at and "strides" in at else k // 2
        if k != 2 * s:
            raise ValueError(f"dec.ups.{i}: kernel {k} / stride {s}: only k == 2*stride is supported")
        kernels.append(k)
        rates.append(s)
    a.up_kernels, a.up_rates = tuple(kernels), tuple(rates)
    rb1 = _count(W, r"dec\.resblocks\.(\d+)\.convs1\.0\.weight")
    rb2 = _count(W, r"dec\.resblocks\.(\d+)\.convs\.0\.weight")
    a.resblock = "1" if rb1 else "2"
    rbs = rb1 or rb2
    nk = len(rbs) // len(ups)
    key = "convs1" if rb1 else "convs"
    ks, ds = [], []
    for j in range(nk):
        ks.append(int(W[f"dec.resblocks.{j}.{key}.0.weight"].shape[2]))
        nconv = len(_count(W, rf"dec\.resblocks\.{j}\.{key}\.(\d+)\.weight"))
        dil = []
        for c in range(nconv):
            at = conv_attrs.get(f"dec.resblocks.{j}.{key}.{c}.weight")
            if at and "dilations" in at:
                dil.append(int(at["dilations"][0]))
        if len(dil) != nconv:
            # no graph (checkpoint source): the reference presets (train.py:106-120)
            if a.resblock == "1":
                dil = [1, 3, 5][:nconv]
            else:
                dil = {3: [1, 2], 5: [2, 6], 7: [3, 12]}.get(ks[-1], [1, 3])[:nconv]
        ds.append(tuple(dil))
    a.rb_kernels, a.rb_dilations = tuple(ks), tuple(ds)
    if metadata and "sample_rate" in metadata:
        a.sample_rate = int(metadata["sample_rate"])
    if sample_rate:
        a.sample_rate = int(sample_rate)
    return a


def load_model(path: str, sample_rate: Optional[int] = None) -> Tuple[Dict[str, np.ndarray], VitsArch, OnnxModel]:
    """Read ``*.onnx`` (or ``*.onnx.gz``) exported by export_onnx.py; returns canonical
    tensors, the inferred architecture and the parsed graph header (inputs/metadata)."""
    p = str(path)
    if p.endswith(".ckpt") or p.endswith(".pt") or p.endswith(".pth"):
        return load_checkpoint(p, sample_rate)
    model = read_onnx(p)
    W, attrs = canonical_from_onnx(model)
    arch = infer_arch(W, attrs, model.metadata, sample_rate)
    has_sid = "sid" in model.inputs
    if has_sid != (arch.n_speakers > 1):
        raise ValueError("graph inputs and emb_g disagree about multi-speaker support")
    if any(k.startswith("emb_l.") or k.startswith("emb_lang.") for k in W):
        # multi-lingual voices (a language embedding selected by `langid`, voice.py:369) condition the text encoder: not built
        raise NotImplementedError("this voice carries a language embedding (multi-lingual VITS): not supported by the B200 engine")
    return W, arch, model


def load_checkpoint(path: str, sample_rate: Optional[int] = None):
    """Lightning ``.ckpt`` (export_onnx.py:227) -> canonical tensors (next-row 8f-3)."""
    import pickle
    import torch

    class _Stub:  # pytorch_lightning.utilities.parsing.AttributeDict, callbacks, loggers, ...: inert placeholders
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, s):
            self.__dict__.update(s if isinstance(s, dict) else {})

        def __call__(self, *a, **k):
            return _Stub()

    # A voice checkpoint is a pickle: unpickling an untrusted one with the stock Unpickler executes whatever callable it names.
    # Only tensor reconstruction and plain containers resolve to real objects here; every other global becomes an inert stub, so a
    # hostile file can at worst fail to load.  (First choice is torch's own weights_only loader, which needs no pickle globals at all.)
    _ALLOWED = {
        ("collections", "OrderedDict"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set"),
        ("builtins", "int"), ("builtins", "float"), ("builtins", "str"), ("builtins", "bool"), ("builtins", "bytes"),
        ("builtins", "slice"), ("builtins", "complex"), ("builtins", "frozenset"),
        ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_parameter"), ("torch._utils", "_rebuild_tensor"),
        ("torch", "Size"), ("torch", "device"), ("torch", "dtype"),
        ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"), ("numpy", "dtype"),
        ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    }
    _STORAGES = {"FloatStorage", "DoubleStorage", "HalfStorage", "BFloat16Storage", "LongStorage", "IntStorage", "ShortStorage",
                 "CharStorage", "ByteStorage", "BoolStorage", "UntypedStorage"}

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if (module, name) in _ALLOWED or (module in ("torch", "torch.storage") and name in _STORAGES):
                return super().find_class(module, name)
            if name == "AttributeDict":
                return dict
            return _Stub

    class _P:
        Unpickler = _Unpickler
        __name__ = "pickle"
        load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())

    try:
        ck = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        ck = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_P)
    sd = ck.get("state_dict", ck)
    W = canonical_from_state_dict(sd)
    arch = infer_arch(W, None, None, sample_rate)
    inputs = ["input", "input_lengths", "scales"] + (["sid"] if arch.n_speakers > 1 else [])
    hdr = OnnxModel("checkpoint", 0, inputs, ["output"], {}, [], {})
    return W, arch, hdr
