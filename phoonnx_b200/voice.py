"""Reference-side glue: make phoonnx's own ``TTSVoice`` run on the B200 engine.

``TTSVoice`` (phoonnx/voice.py:105-172) only ever touches its backend through
``session.get_inputs()`` (:347) and ``session.run(None, feed)`` (:374-377), and builds it in
``TTSVoice.load`` (:167-171) from the module-level name ``onnxruntime``.  So the drop-in is:

    from phoonnx_b200.voice import load_voice
    voice = load_voice("voice.onnx")            # a real phoonnx.voice.TTSVoice, B200 underneath
    voice.synthesize_wav("hello", wav_file)     # unchanged reference code path

or, without touching call sites, ``patch_phoonnx()`` which swaps the ``onnxruntime`` name inside
``phoonnx.voice`` so that ``TTSVoice.load(..., use_cuda=True)`` constructs a ``B200Session``.
``SynthesisConfig`` / ``VoiceConfig`` / ``phoneme_ids`` are used as they are.  phoonnx itself is
NOT a dependency of this package: these helpers import it lazily and fail with ImportError if
it is not installed.
"""
from __future__ import annotations

import json
import types
from typing import List, Optional, Sequence

import numpy as np

from .session import B200Session


def load_voice(model_path, config_path=None, device: int = 0, precision: str = "fp32", **voice_config_kwargs):
    """Same contract as ``TTSVoice.load`` (voice.py:125-172) with the session replaced."""
    from phoonnx.config import VoiceConfig          # reference code, unchanged
    from phoonnx.voice import TTSVoice
    if config_path is None:
        config_path = f"{model_path}.json"          # voice.py:143-145
    with open(config_path, "r", encoding="utf-8") as f:
        config_dict = json.load(f)
    cfg = VoiceConfig.from_dict(config_dict, **voice_config_kwargs)
    sess = B200Session(str(model_path), device=device, precision=precision, sample_rate=cfg.sample_rate)
    return TTSVoice(config=cfg, session=sess)


def patch_phoonnx(device: int = 0, precision: str = "fp32") -> None:
    """After this, ``phoonnx.voice.TTSVoice.load(path, use_cuda=True)`` builds a B200Session; with
    ``use_cuda=False`` the original onnxruntime CPU session is still used (if onnxruntime exists)."""
    import phoonnx.voice as pv
    real = pv.onnxruntime

    def make_session(path, sess_options=None, providers=None, **kw):
        names = [p if isinstance(p, str) else p[0] for p in (providers or [])]
        if "CUDAExecutionProvider" in names:
            return B200Session(path, sess_options, providers, device=device, precision=precision)
        return real.InferenceSession(path, sess_options=sess_options, providers=providers, **kw)

    pv.onnxruntime = types.SimpleNamespace(
        InferenceSession=make_session,
        SessionOptions=getattr(real, "SessionOptions", lambda: None),
    )


def synthesize_batch(voice, phoneme_id_lists: Sequence[Sequence[int]], syn_config=None) -> List[np.ndarray]:
    """Batched form of ``TTSVoice.phoneme_ids_to_audio`` (voice.py:328-379): all
    sentences go to the engine as one varlen batch instead of one ``run()`` per sentence
    (voice.py:265-269).  Returns one float32 array per sentence, identical to what the per-sentence
    call returns for the same noise."""
    sess = voice.session
    if not isinstance(sess, B200Session):
        raise TypeError("synthesize_batch needs a voice backed by B200Session")
    cfg = voice.config
    sc = syn_config
    g = lambda name, default: default if sc is None or getattr(sc, name, None) is None else getattr(sc, name)  # noqa: E731
    scales = np.array([g("noise_scale", cfg.noise_scale), g("length_scale", cfg.length_scale),
                       g("noise_w_scale", cfg.noise_w_scale)], dtype=np.float32)        # order: voice.py:364-367
    B = len(phoneme_id_lists)
    lens = np.array([len(p) for p in phoneme_id_lists], np.int64)
    x = np.zeros((B, int(lens.max())), np.int64)
    for b, p in enumerate(phoneme_id_lists):
        x[b, : len(p)] = np.asarray(p, np.int64)
    feed = {"input": x, "input_lengths": lens, "scales": scales}
    if sess.arch.n_speakers > 1:
        feed["sid"] = np.full((B,), g("speaker_id", 0) or 0, np.int64)
    audio, alen = sess.synthesize_packed(feed)
    out, off = [], 0
    for b in range(B):
        out.append(audio[off:off + int(alen[b])])
        off += int(alen[b])
    return out
