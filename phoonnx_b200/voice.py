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


def _syn_value(sc, cfg, name, default=None):
    v = None if sc is None else getattr(sc, name, None)
    if v is None:
        v = getattr(cfg, name, default)
    return default if v is None else v


def ids_to_audio_batch(voice, phoneme_id_lists: Sequence[Sequence[int]], syn_config=None) -> List[np.ndarray]:
    """Batched form of ``TTSVoice.phoneme_ids_to_audio`` (voice.py:328-379): all sentences go to the engine as ONE varlen
    batch instead of one ``run()`` per sentence (voice.py:265-269).  Returns one float32 array per sentence, identical to what
    the per-sentence call returns for the same noise (every utterance is synthesised with B=1 semantics)."""
    sess = voice.session
    if not isinstance(sess, B200Session):
        raise TypeError("synthesize_batch needs a voice backed by B200Session")
    if not len(phoneme_id_lists):
        return []
    cfg = voice.config
    scales = np.array([_syn_value(syn_config, cfg, "noise_scale", 0.667), _syn_value(syn_config, cfg, "length_scale", 1.0),
                       _syn_value(syn_config, cfg, "noise_w_scale", 0.8)], dtype=np.float32)        # order: voice.py:364-367
    B = len(phoneme_id_lists)
    lens = np.array([len(p) for p in phoneme_id_lists], np.int64)
    if lens.min() < 1:
        raise ValueError("empty phoneme-id sequence")
    x = np.zeros((B, int(lens.max())), np.int64)
    for b, p in enumerate(phoneme_id_lists):
        x[b, : len(p)] = np.asarray(p, np.int64)
    feed = {"input": x, "input_lengths": lens, "scales": scales}
    if sess.arch.n_speakers > 1:
        sid = _syn_value(syn_config, cfg, "speaker_id", 0) or 0                                     # voice.py:357, 370
        feed["sid"] = np.full((B,), int(sid), np.int64)
    audio, alen = sess.synthesize_packed(feed)
    out, off = [], 0
    for b in range(B):
        out.append(audio[off:off + int(alen[b])])
        off += int(alen[b])
    return out


def _postprocess(audio: np.ndarray, syn_config) -> np.ndarray:
    """The caller-side steps of TTSVoice.synthesize (voice.py:271-282), unchanged."""
    if syn_config is None or getattr(syn_config, "normalize_audio", True):
        max_val = np.max(np.abs(audio)) if audio.size else 0.0
        audio = np.zeros_like(audio) if max_val < 1e-8 else audio / max_val
    vol = 1.0 if syn_config is None else getattr(syn_config, "volume", 1.0)
    if vol != 1.0:
        audio = audio * vol
    return np.clip(audio, -1.0, 1.0).astype(np.float32)


def _text_to_id_lists(voice, text: str, syn_config) -> List[List[int]]:
    """Everything TTSVoice.synthesize does to a text before the session is called (voice.py:245-263): phonetic spellings,
    diacritics, phonemisation, id mapping -- the reference's own methods, on the CPU, untouched."""
    sc = syn_config
    spell = getattr(voice, "phonetic_spellings", None)
    if spell and (sc is None or getattr(sc, "enable_phonetic_spellings", True)):
        text = spell.apply(text)
    if sc is not None and getattr(sc, "add_diacritics", False):
        text = voice.phonemizer.add_diacritics(text, voice.config.lang_code)
    return [ids for ids in (voice.phonemes_to_ids(ph) for ph in voice.phonemize(text) if ph) if ids]


def synthesize_batch(voice, texts, syn_config=None, max_workers: Optional[int] = None):
    """Batched ``TTSVoice.synthesize`` (voice.py:236-290).

    ``texts``: a sequence of strings (or, for callers that phonemise themselves, of phoneme-id lists).  Texts are phonemised
    on a CPU thread pool with the voice's own phonemizer (the reference loops over them serially); ALL sentences of ALL texts
    then go to the engine as one varlen batch, and the per-sentence post-processing of ``synthesize`` (normalise, volume,
    clip) is applied to each result.  Returns one list per input: for a text, its ``AudioChunk``s in sentence order (the
    reference's class when phoonnx is importable); for an id list, a single float32 array (no post-processing, exactly
    ``phoneme_ids_to_audio``)."""
    texts = list(texts)
    if not texts:
        return []
    is_text = [isinstance(t, str) for t in texts]
    if any(is_text):
        from concurrent.futures import ThreadPoolExecutor
        idx = [i for i, f in enumerate(is_text) if f]
        with ThreadPoolExecutor(max_workers=max_workers or min(32, len(idx))) as pool:
            sent = list(pool.map(lambda i: _text_to_id_lists(voice, texts[i], syn_config), idx))
        per_text = dict(zip(idx, sent))
    else:
        per_text = {}
    flat: List[Sequence[int]] = []
    spans = []
    for i, t in enumerate(texts):
        lists = per_text[i] if is_text[i] else [list(t)]
        spans.append((len(flat), len(flat) + len(lists)))
        flat.extend(lists)
    audios = ids_to_audio_batch(voice, flat, syn_config) if flat else []
    chunk_cls = None
    if any(is_text):
        import sys
        chunk_cls = getattr(sys.modules.get(type(voice).__module__), "AudioChunk", None)
    out = []
    for i, (lo, hi) in enumerate(spans):
        if not is_text[i]:
            out.append(audios[lo])
            continue
        chunks = []
        for a in audios[lo:hi]:
            a = _postprocess(np.asarray(a), syn_config)
            if chunk_cls is not None:
                chunks.append(chunk_cls(sample_rate=voice.config.sample_rate, sample_width=2, sample_channels=1, audio_float_array=a))
            else:
                chunks.append(a)
        out.append(chunks)
    return out
