"""phoonnx_b200.voice on top of the UNMODIFIED reference caller (phoonnx.voice.TTSVoice), engine stubbed: `load_voice` builds a
real TTSVoice around our session (voice.py:125-172), `synthesize_batch` phonemises texts on CPU threads with the voice's own
methods and submits ALL sentences as one varlen batch (reference loop voice.py:259-269).  Needs the reference
checkout (build container only); the GPU-side twin with a real engine is tests/test_gpu_boundary.py."""
import json
import sys
import threading
import time
import types

import numpy as np
import pytest

from oracle import ref_bridge as rb

pytestmark = pytest.mark.skipif(not rb.reference_available(), reason="reference checkout not present")


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None


@pytest.fixture(scope="module")
def phoonnx_voice():
    saved = dict(sys.modules)
    for name in ("onnxruntime", "langcodes", "quebra_frases", "ovos_date_parser", "ovos_number_parser",
                 "ovos_number_parser.util", "unicode_rbnf"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    sys.modules["langcodes"].closest_match = lambda lang, langs: ("und", 1000)
    if rb.REF_ROOT not in sys.path:
        sys.path.insert(0, rb.REF_ROOT)
    import phoonnx.voice as pv
    yield pv
    for k in list(sys.modules):
        if k not in saved and (k.startswith("phoonnx.") or k == "phoonnx" or k in ("onnxruntime", "langcodes")):
            del sys.modules[k]


ID_MAP = {"_": [0], "^": [1], "$": [2], " ": [3], "a": [5], "b": [6], ".": [7]}
CFG = {"phoneme_type": "raw", "lang_code": "en", "phoneme_id_map": ID_MAP, "audio": {"sample_rate": 22050}, "num_speakers": 1}


class CharPhonemizer:
    """Stand-in for the reference's phonemizers (their text normalisation needs third-party data absent offline): sentences
    split on '.', one phoneme per character.  Records the threads it ran on."""

    def __init__(self):
        self.threads = set()

    def phonemize(self, text, lang):
        self.threads.add(threading.get_ident())
        time.sleep(0.05)
        return [list(s.strip()) + ["."] for s in text.split(".") if s.strip()]

    def add_diacritics(self, text, lang):
        return text


def _fake_session(n_speakers=1):
    from phoonnx_b200.modelgen import make_arch
    from phoonnx_b200.session import B200Session, NodeArg

    class Fake(B200Session):
        def synthesize_packed(self, feed, out="f32", volume=1.0, normalize=True, asynchronous=False):
            x, lens, scales, sid = self._unpack_feed(feed)          # the real session's own validation
            self.feeds.append({k: np.array(v) for k, v in feed.items()})
            alen = lens * 256
            audio = np.concatenate([np.linspace(-0.1, 0.2, int(n), dtype=np.float32) * (b + 1) for b, n in enumerate(alen)])
            return audio, alen

    s = Fake.__new__(Fake)
    s.arch = make_arch("tiny", n_speakers)
    s._inputs = [NodeArg("input", "tensor(int64)", None), NodeArg("input_lengths", "tensor(int64)", None),
                 NodeArg("scales", "tensor(float)", [3])] + ([NodeArg("sid", "tensor(int64)", None)] if n_speakers > 1 else [])
    s.feeds = []
    return s


def test_load_voice_builds_a_reference_ttsvoice(phoonnx_voice, monkeypatch, tmp_path):
    pv = phoonnx_voice
    import phoonnx_b200.voice as shim
    made = {}

    class FakeB200:
        def __init__(self, path, **kw):
            made["path"], made["kw"] = path, kw

        def get_inputs(self):
            return []

    monkeypatch.setattr(shim, "B200Session", FakeB200)
    model = tmp_path / "voice.onnx"
    model.write_bytes(b"")
    (tmp_path / "voice.onnx.json").write_text(json.dumps(CFG))            # default config path: model + ".json" (voice.py:143-145)
    voice = shim.load_voice(str(model), device=0, precision="bf16")
    assert isinstance(voice, pv.TTSVoice) and isinstance(voice.session, FakeB200)
    assert made["path"] == str(model) and made["kw"]["precision"] == "bf16" and made["kw"]["sample_rate"] == 22050
    assert voice.config.sample_rate == 22050 and voice.config.num_speakers == 1
    other = tmp_path / "cfg.json"
    other.write_text(json.dumps(dict(CFG, audio={"sample_rate": 16000})))
    assert shim.load_voice(str(model), str(other)).config.sample_rate == 16000 and made["kw"]["sample_rate"] == 16000


def test_synthesize_batch_texts_through_the_reference_voice(phoonnx_voice):
    pv = phoonnx_voice
    from phoonnx.config import SynthesisConfig, VoiceConfig
    from phoonnx_b200.voice import synthesize_batch
    ph = CharPhonemizer()
    sess = _fake_session()
    voice = pv.TTSVoice(session=sess, config=VoiceConfig.from_dict(dict(CFG)), phonemizer=ph)
    texts = ["ab ba. aab b.", "b.", "a ab.", "bb. a.", [1, 0, 5, 0, 2]]
    sc = SynthesisConfig(noise_scale=0.3, length_scale=1.2, noise_w_scale=0.5, volume=0.5)
    out = synthesize_batch(voice, texts, sc, max_workers=4)
    # ONE engine call for every sentence of every text (the reference issues one run() per sentence, voice.py:265-269)
    assert len(sess.feeds) == 1
    f = sess.feeds[0]
    want = []
    for t in texts:
        want += [voice.phonemes_to_ids(p) for p in voice.phonemize(t)] if isinstance(t, str) else [list(t)]
    assert f["input"].shape[0] == len(want) and f["input_lengths"].tolist() == [len(w) for w in want]
    for b, w in enumerate(want):
        assert f["input"][b, :len(w)].tolist() == w and not f["input"][b, len(w):].any()
    assert f["scales"].dtype == np.float32 and np.allclose(f["scales"], [0.3, 1.2, 0.5])          # order: voice.py:364-367
    assert len(ph.threads) > 1                                                                   # phonemised on a thread pool
    # structure: per text its AudioChunks in sentence order; per id list one raw array
    assert len(out) == len(texts)
    b = 0
    for t, o in zip(texts, out):
        if isinstance(t, str):
            n = len(voice.phonemize(t))
            assert len(o) == n and all(isinstance(c, pv.AudioChunk) for c in o)
            for c in o:
                a = c.audio_float_array
                assert a.dtype == np.float32 and a.shape == (256 * len(want[b]),)
                assert abs(float(np.abs(a).max()) - 0.5) < 1e-6                                   # normalised, then volume 0.5
                assert c.sample_rate == 22050 and c.audio_int16_array.dtype == np.int16
                b += 1
        else:
            assert isinstance(o, np.ndarray) and o.shape == (256 * len(t),)                       # exactly phoneme_ids_to_audio
            assert np.allclose(o, np.linspace(-0.1, 0.2, 256 * len(t), dtype=np.float32) * (b + 1))
            b += 1
    assert synthesize_batch(voice, []) == []


def test_synthesize_batch_needs_our_session(phoonnx_voice):
    pv = phoonnx_voice
    from phoonnx.config import VoiceConfig
    from phoonnx_b200.voice import synthesize_batch

    class NotOurs:
        def get_inputs(self):
            return []

    voice = pv.TTSVoice(session=NotOurs(), config=VoiceConfig.from_dict(dict(CFG)), phonemizer=CharPhonemizer())
    with pytest.raises(TypeError):
        synthesize_batch(voice, ["a."])


def test_synthesize_many_keeps_the_finished_result_when_a_feed_raises():
    """a later feed that raises must not drop the result already in flight; blocking calls from inside the consumer
    loop are untouched by the generator (asynchrony is per call, waited for by ticket)."""
    from phoonnx_b200.session import B200Session

    class Eng:
        hop = 256

        def __init__(self):
            self.last_ticket, self.waited, self.n = 0, [], 0

        def wait_ticket(self, t):
            self.waited.append(t)

    class S(B200Session):
        def synthesize_packed(self, feed, out="f32", volume=1.0, normalize=True, asynchronous=False):
            if feed == "bad":
                raise ValueError("boom")
            self.engine.n += 1
            if asynchronous:
                self.engine.last_ticket = self.engine.n
            return np.full((4,), float(self.engine.n), np.float32), np.array([4])

    s = S.__new__(S)
    s.engine = Eng()
    got = []
    with pytest.raises(ValueError):
        for audio, alen in s.synthesize_many([1, 2, "bad", 3]):
            got.append(float(audio[0]))
            if len(got) == 1:
                s.synthesize_packed(99)                     # a blocking call from inside the consumer loop
    assert got == [1.0, 2.0] or got == [1.0, 3.0]          # both finished results were yielded before the error surfaced
    assert len(got) == 2 and s.engine.waited == sorted(s.engine.waited) and len(s.engine.waited) == 2
