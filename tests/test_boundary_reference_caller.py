"""Boundary conformance: the UNMODIFIED reference caller (phoonnx.voice.TTSVoice) on top of a
session object shaped like ours.  Needs the reference checkout (build container only); the text-side
third-party modules that are absent offline are stubbed."""
import os
import sys
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


class RecordingSession:
    """What B200Session looks like to TTSVoice, minus the GPU: validates feeds with B200Session's own checker."""

    def __init__(self, n_speakers):
        from phoonnx_b200.session import B200Session, NodeArg
        from phoonnx_b200.modelgen import make_arch
        self._v = B200Session.__new__(B200Session)
        self._v.arch = make_arch("tiny", n_speakers)
        self._v._inputs = [NodeArg("input", "tensor(int64)", None), NodeArg("input_lengths", "tensor(int64)", None),
                           NodeArg("scales", "tensor(float)", [3])] + ([NodeArg("sid", "tensor(int64)", None)] if n_speakers > 1 else [])
        self.feeds = []

    def get_inputs(self):
        return self._v.get_inputs()

    def run(self, names, feed):
        assert names is None
        x, lens, scales, sid = self._v._unpack_feed(feed)     # raises on anything our session would reject
        self.feeds.append({k: np.array(v) for k, v in feed.items()})
        return [np.linspace(-0.5, 0.25, 1280, dtype=np.float32).reshape(1, 1, 1, 1280)]


@pytest.mark.parametrize("n_speakers", [1, 8])
def test_reference_ttsvoice_feeds_our_session(phoonnx_voice, n_speakers):
    pv = phoonnx_voice
    from phoonnx.config import SynthesisConfig, VoiceConfig
    cfg = VoiceConfig.from_dict({"phoneme_type": "raw", "lang_code": "en", "phoneme_id_map": {"_": [0], "^": [1], "$": [2]},
                                 "audio": {"sample_rate": 22050}, "num_speakers": n_speakers})
    sess = RecordingSession(n_speakers)
    voice = pv.TTSVoice(session=sess, config=cfg)
    audio = voice.phoneme_ids_to_audio([1, 0, 20, 0, 2], SynthesisConfig(speaker_id=3))
    assert audio.shape == (1280,) and audio.dtype == np.float32           # .squeeze() of [1,1,1,N] (voice.py:374-377)
    f = sess.feeds[-1]
    assert f["input"].dtype == np.int64 and f["input"].tolist() == [[1, 0, 20, 0, 2]]
    assert f["input_lengths"].tolist() == [5]
    assert f["scales"].dtype == np.float32 and np.allclose(f["scales"], [0.667, 1.0, 0.8])   # config.py:9-11 defaults
    assert ("sid" in f) == (n_speakers > 1) and "langid" not in f                            # filtered by get_inputs names
    if n_speakers > 1:
        assert f["sid"].tolist() == [3]
    voice.phoneme_ids_to_audio([1, 2], SynthesisConfig(noise_scale=0.1, length_scale=1.5, noise_w_scale=0.2))
    assert np.allclose(sess.feeds[-1]["scales"], [0.1, 1.5, 0.2])                            # order noise, length, noise_w


def test_patch_routes_use_cuda_to_b200(phoonnx_voice, monkeypatch, tmp_path):
    pv = phoonnx_voice
    import phoonnx_b200.voice as shim
    made = {}

    class FakeB200:
        def __init__(self, path, sess_options=None, providers=None, **kw):
            made["path"], made["kw"] = path, kw

    monkeypatch.setattr(shim, "B200Session", FakeB200)
    monkeypatch.setattr(pv, "onnxruntime", pv.onnxruntime)
    shim.patch_phoonnx(device=0, precision="bf16")
    s = pv.onnxruntime.InferenceSession("m.onnx", sess_options=None,
                                        providers=[("CUDAExecutionProvider", {"cudnn_conv_algo_search": "HEURISTIC"})])
    assert isinstance(s, FakeB200) and made["path"] == "m.onnx" and made["kw"]["precision"] == "bf16"
