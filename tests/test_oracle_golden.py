"""The oracle (oracle/vits_oracle.py) against the golden fixtures minted from the REAL reference
(oracle/make_golden.py ran phoonnx_train's SynthesizerTrn.infer, models.py:681-722, and the
reference's exporter logic, export_onnx.py:250-327).  Runs anywhere, no reference needed."""
import os

import numpy as np
import pytest

from oracle.vits_oracle import VitsOracle, postprocess_int16
from phoonnx_b200.weights import load_model

VOICES = ["tiny_spk1", "tiny_spk3", "tiny_rb1_spk1"]
N_UTT = 6
STAGE_TOL = 2e-5      # fp32 re-association between the reference's formulation and the restatement


@pytest.fixture(scope="module", params=VOICES)
def voice(request, golden_dir):
    name = request.param
    W, arch, hdr = load_model(os.path.join(golden_dir, name + ".onnx.gz"))
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    return name, VitsOracle(W, arch), arch, gold, hdr


def test_genuine_export_is_understood(voice):
    name, _, arch, _, hdr = voice
    assert hdr.producer == "pytorch" and hdr.opset == 15
    assert hdr.outputs == ["output"]
    assert hdr.inputs[:3] == ["input", "input_lengths", "scales"]
    assert ("sid" in hdr.inputs) == (arch.n_speakers > 1)
    assert arch.hidden == 32 and arch.inter == 32 and arch.filter == 64 and arch.n_layers == 2
    assert arch.window == 4 and arch.n_heads == 2 and arch.cflows == (7, 5, 3) and arch.flow_layers == (6, 4, 2, 0)
    if "rb1" in name:
        assert arch.resblock == "1" and arch.rb_dilations == ((1, 3, 5), (1, 3, 5)) and arch.up_rates == (4, 2, 2)
    else:
        assert arch.resblock == "2" and arch.rb_dilations == ((1, 2), (2, 6)) and arch.up_rates == (4, 4, 2)
    assert arch.n_speakers == (3 if "spk3" in name else 1)


@pytest.mark.parametrize("u", range(N_UTT))
@pytest.mark.parametrize("tag", ["n", "z"])
def test_oracle_matches_reference_outputs(voice, u, tag):
    _, orc, arch, g, _ = voice
    k = f"u{u}{tag}_"
    sid = int(g[k + "sid"])
    nd = g[k + "noise_dp"] if (k + "noise_dp") in g else None
    nz = g[k + "noise_z"] if (k + "noise_z") in g else None
    r = orc.infer(g[k + "ids"], tuple(g[k + "scales"]), None if sid < 0 else sid, nd, nz)
    # integer path: bit-exact
    assert np.array_equal(r["durations"], g[k + "durations"])
    assert r["audio"].shape == g[k + "audio"].shape == (int(g[k + "durations"].sum()) * arch.hop,) or g[k + "durations"].sum() == 0
    for s in ("x", "m_p", "logs_p", "logw", "z_p", "z"):
        assert r[s].shape == g[k + s].shape, s
        assert np.abs(r[s] - g[k + s]).max() <= STAGE_TOL * max(1.0, np.abs(g[k + s]).max()), s
    assert np.abs(r["audio"] - g[k + "audio"]).max() <= 1e-5


def test_frame_index_is_generate_path(voice):
    """commons.generate_path + attn@m_p == gather by searchsorted index, incl. zero durations."""
    import torch
    dur = torch.tensor([2, 0, 3, 1, 0, 0, 4])
    idx = VitsOracle.frame_index(dur, int(dur.sum()))
    assert idx.tolist() == [0, 0, 2, 2, 2, 3, 6, 6, 6, 6]
    # all-zero durations: y_len clamps to 1 and the single frame selects nothing
    z = torch.zeros(4, dtype=torch.int64)
    assert VitsOracle.frame_index(z, 1).tolist() == [-1]


def test_durations_rule():
    import torch
    logw = torch.tensor([0.0, np.log(2.0) + 1e-3, -50.0, 1.0])
    d, y = VitsOracle.durations_from_logw(logw, 1.0)
    assert d.tolist() == [1, 3, 1, 3] and y == 8
    d, y = VitsOracle.durations_from_logw(torch.tensor([-200.0]), 1.0)   # exp underflows to 0 -> ceil 0 -> clamp 1
    assert d.tolist() == [0] and y == 1


def test_phoneme_id_layout_example(golden_dir):
    """ID layout the engine consumes (phoneme_ids.py:184-187,242-308): bos 1, blank 0 interleaved, space 3, eos 2."""
    ids = np.load(os.path.join(golden_dir, "phoneme_ids_example.npy")).tolist()
    assert ids == [1, 0, 20, 0, 59, 0, 24, 0, 120, 0, 27, 0, 100, 0, 3, 0, 35, 0, 120, 0, 62, 0, 122, 0, 24, 0, 17, 0, 2]
    assert ids[0] == 1 and ids[-1] == 2 and all(v == 0 for v in ids[1::2])


def test_postprocess_int16_matches_voice_py():
    """voice.py:271-282 + 88-91."""
    a = np.array([0.0, 0.25, -0.5, 0.1], np.float32)
    out = postprocess_int16(a)
    assert out.dtype == np.int16 and out.tolist() == [0, 16383, -32767, 6553]
    assert postprocess_int16(np.zeros(4, np.float32)).tolist() == [0, 0, 0, 0]
    assert postprocess_int16(a, volume=4.0, normalize=False).tolist() == [0, 32767, -32767, 13106]
