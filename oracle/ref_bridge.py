"""TEST INFRASTRUCTURE ONLY -- bridge to the *real* reference (phoonnx) PyTorch model.

This module imports ``phoonnx_train.vits.models.SynthesizerTrn`` straight from the
read-only reference checkout (``$PHOONNX_REF``, default ``/root/reference``) and is used
ONLY to (a) pin ``oracle/vits_oracle.py`` against the reference's own arithmetic and
(b) mint the golden fixtures committed under ``tests/golden`` (see ``make_golden.py``).
Nothing in the product path (``phoonnx_b200``) imports it, and it is never available on
the GPU box (the reference checkout does not travel).

Recipes follow SURVEY.md Appendix C:
  * the Cython ``monotonic_align`` package is stubbed (training-only, models.py:646);
  * model construction uses the argument mapping of lightning.py:86-106 and the quality
    presets of train.py:106-120;
  * ``dec.remove_weight_norm()`` as export_onnx.py:242-245;
  * ONNX export re-states export_onnx.py:250-327 with the legacy TorchScript exporter and
    the ``onnx`` round-trip patched out (the ``onnx`` package is absent offline).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from typing import Dict, Optional

import numpy as np
import torch

REF_ROOT = os.environ.get("PHOONNX_REF", "/root/reference")

PRESETS = {
    # train.py:106-120 + lightning.py defaults
    "x_low": dict(hidden_channels=96, inter_channels=96, filter_channels=384),
    "medium": dict(),
    "high": dict(
        resblock="1",
        resblock_kernel_sizes=(3, 7, 11),
        resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)),
        upsample_rates=(8, 8, 2, 2),
        upsample_initial_channel=512,
        upsample_kernel_sizes=(16, 16, 4, 4),
    ),
    # not a reference preset: a miniature of the medium topology so that a *genuine*
    # exporter-format .onnx is small enough to commit as a fixture.
    "tiny": dict(
        hidden_channels=32, inter_channels=32, filter_channels=64, n_layers=2,
        upsample_initial_channel=32, upsample_rates=(4, 4, 2), upsample_kernel_sizes=(8, 8, 4),
        resblock_kernel_sizes=(3, 5), resblock_dilation_sizes=((1, 2), (2, 6)),
    ),
    # miniature of the high topology (ResBlock1)
    "tiny_rb1": dict(
        hidden_channels=32, inter_channels=32, filter_channels=64, n_layers=2,
        resblock="1", upsample_initial_channel=32, upsample_rates=(4, 2, 2),
        upsample_kernel_sizes=(8, 4, 4),
        resblock_kernel_sizes=(3, 7), resblock_dilation_sizes=((1, 3, 5), (1, 3, 5)),
    ),
}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "phoonnx_train", "vits"))


def import_reference():
    """Return the reference ``SynthesizerTrn`` class (models.py:522)."""
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    name = "phoonnx_train.vits.monotonic_align"
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
    from phoonnx_train.vits.models import SynthesizerTrn  # noqa: WPS433

    return SynthesizerTrn


def build_reference_model(preset: str = "medium", n_speakers: int = 1, n_vocab: int = 256,
                          seed: int = 1234, dezero_seed: Optional[int] = 4321,
                          use_sdp: bool = True):
    """Random-init reference model, eval mode, decoder weight-norm removed, de-zeroed."""
    SynthesizerTrn = import_reference()
    kw = dict(
        n_vocab=n_vocab, spec_channels=513, segment_size=32, inter_channels=192,
        hidden_channels=192, filter_channels=768, n_heads=2, n_layers=6, kernel_size=3,
        p_dropout=0.1, resblock="2", resblock_kernel_sizes=(3, 5, 7),
        resblock_dilation_sizes=((1, 2), (2, 6), (3, 12)), upsample_rates=(8, 8, 4),
        upsample_initial_channel=256, upsample_kernel_sizes=(16, 16, 8),
        n_speakers=n_speakers, gin_channels=(512 if n_speakers > 1 else 0), use_sdp=use_sdp,
    )
    kw.update(PRESETS[preset])
    if preset.startswith("tiny") and n_speakers > 1:
        kw["gin_channels"] = 16
    torch.manual_seed(seed)
    import warnings
    with warnings.catch_warnings(), contextlib.redirect_stdout(open(os.devnull, "w")):
        warnings.simplefilter("ignore")
        m = SynthesizerTrn(**kw).eval()
        m.dec.remove_weight_norm()  # export_onnx.py:242-245
    if dezero_seed is not None:
        dezero(m, dezero_seed)
    return m


@torch.no_grad()
def dezero(model, seed: int = 4321) -> None:
    """SURVEY.md 8(c) caveat 2: random-init VITS has zero-initialised layers
    (modules.py:398-399,444-445,493-494) that turn the flows into identities; give them
    small random values so every kernel is exercised, and pin the duration affine so the
    durations are realistic (mean ~3.5 frames/id)."""
    g = torch.Generator().manual_seed(seed)

    def rn(t, std, mean=0.0):
        t.copy_(torch.randn(t.shape, generator=g) * std + mean)

    sd = dict(model.named_parameters())
    for k in sorted(sd):
        v = sd[k]
        if k.startswith("flow.flows.") and (k.endswith(".post.weight") or k.endswith(".post.bias")):
            rn(v, 0.05)
        elif k.startswith("dp.flows.") and (k.endswith(".proj.weight") or k.endswith(".proj.bias")) \
                and ".convs." not in k:
            rn(v, 0.05)
        elif k.endswith(".gamma"):
            rn(v, 0.1, 1.0)
        elif k.endswith(".beta"):
            rn(v, 0.1)
        elif k in ("dp.flows.0.m", "dp.flows.0.logs"):
            rn(v, 0.3)
    if "dp.flows.0.m" in sd:
        sd["dp.flows.0.m"][0] = -2.0
        sd["dp.flows.0.logs"][0] = 0.7


@contextlib.contextmanager
def injected_noise(noise_dp: Optional[torch.Tensor], noise_z: Optional[torch.Tensor]):
    """Patch torch.randn / torch.randn_like (models.py:111,718) to return given tensors."""
    import phoonnx_train.vits.models as M

    real_randn, real_randn_like = torch.randn, torch.randn_like

    def fake_randn(*size, **kw):
        if noise_dp is None:
            return real_randn(*size, **kw)
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        assert tuple(noise_dp.shape) == tuple(int(s) for s in shape), (noise_dp.shape, shape)
        return noise_dp.clone()

    def fake_randn_like(t, **kw):
        if noise_z is None:
            return real_randn_like(t, **kw)
        assert noise_z.shape[0] == t.shape[0] and noise_z.shape[1] == t.shape[1]
        return noise_z[:, :, : t.shape[2]].clone()

    M.torch.randn, M.torch.randn_like = fake_randn, fake_randn_like
    try:
        yield
    finally:
        M.torch.randn, M.torch.randn_like = real_randn, real_randn_like


@torch.no_grad()
def reference_infer(model, ids: np.ndarray, scales=(0.667, 1.0, 0.8), sid: Optional[int] = None,
                    noise_dp: Optional[np.ndarray] = None, noise_z: Optional[np.ndarray] = None
                    ) -> Dict[str, np.ndarray]:
    """One utterance (B=1, the semantics TTSVoice uses, voice.py:350) through the
    reference's SynthesizerTrn.infer (models.py:681-722) with stage tensors captured."""
    x = torch.as_tensor(np.asarray(ids, dtype=np.int64))[None]
    lens = torch.tensor([x.shape[1]], dtype=torch.int64)
    sid_t = None if sid is None else torch.tensor([int(sid)], dtype=torch.int64)
    nd = None if noise_dp is None else torch.as_tensor(noise_dp, dtype=torch.float32)[None]
    nz = None if noise_z is None else torch.as_tensor(noise_z, dtype=torch.float32)[None]
    cap = {}
    hooks = []

    def grab(name):
        def fn(_m, _i, out):
            cap[name] = out
        return fn

    hooks.append(model.enc_p.register_forward_hook(grab("enc_p")))
    hooks.append(model.dp.register_forward_hook(grab("logw")))
    try:
        with injected_noise(nd, nz):
            o, attn, y_mask, (z, z_p, m_p, logs_p) = model.infer(
                x, lens, sid=sid_t, noise_scale=float(scales[0]), length_scale=float(scales[1]),
                noise_scale_w=float(scales[2]))
    finally:
        for h in hooks:
            h.remove()
    xe, m_t, logs_t, _ = cap["enc_p"]
    dur = attn[0, 0].sum(0).round().to(torch.int32)  # [T]
    return {
        "x": xe[0].T.contiguous().numpy(),            # [T, H]  (channel-last)
        "m_p": m_t[0].T.contiguous().numpy(),
        "logs_p": logs_t[0].T.contiguous().numpy(),
        "logw": cap["logw"][0, 0].numpy(),            # [T]
        "durations": dur.numpy(),
        "z_p": z_p[0].T.contiguous().numpy(),         # [Ty, C]
        "z": z[0].T.contiguous().numpy(),
        "audio": o[0, 0].numpy(),
    }


def export_onnx(model, path: str, n_speakers: int = 1, n_vocab: int = 256) -> None:
    """Re-statement of export_onnx.py:250-327 that works offline (no ``onnx`` package)."""
    import warnings
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    onnx_proto_utils._add_onnxscript_fn = lambda b, c: b  # the only use of ``onnx``

    model_g = model

    def infer_forward(text, text_lengths, scales, sid=None):
        audio = model_g.infer(text, text_lengths, noise_scale=scales[0], length_scale=scales[1],
                              noise_scale_w=scales[2], sid=sid)[0].unsqueeze(1)
        return audio

    orig_forward = model_g.forward
    model_g.forward = infer_forward
    torch.manual_seed(1234)
    sequences = torch.randint(low=0, high=n_vocab, size=(1, 50), dtype=torch.long)
    sequence_lengths = torch.LongTensor([sequences.size(1)])
    sid = None
    input_names = ["input", "input_lengths", "scales"]
    dyn = {"input": {0: "batch_size", 1: "phonemes"}, "input_lengths": {0: "batch_size"},
           "output": {0: "batch_size", 1: "time"}}
    if n_speakers > 1:
        sid = torch.LongTensor([0])
        input_names.append("sid")
        dyn["sid"] = {0: "batch_size"}
    scales = torch.FloatTensor([0.667, 1.0, 0.8])
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch.onnx.export(model=model_g, args=(sequences, sequence_lengths, scales, sid), f=str(path),
                              verbose=False, opset_version=15, input_names=input_names,
                              output_names=["output"], dynamic_axes=dyn, dynamo=False)
    finally:
        model_g.forward = orig_forward
