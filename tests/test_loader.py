"""Weight loader: exporter-format files -> canonical tensors + inferred architecture."""
import gzip
import os

import numpy as np
import pytest

from phoonnx_b200 import modelgen, onnx_reader
from phoonnx_b200.weights import canonical_from_state_dict, infer_arch, load_model


@pytest.mark.parametrize("preset,ns,sdp", [("tiny", 1, True), ("tiny", 3, True), ("tiny_rb1", 1, False),
                                           ("x_low", 1, True), ("high", 1, True), ("medium", 8, True)])
def test_synthetic_voice_roundtrip(tmp_path, preset, ns, sdp):
    p = str(tmp_path / "v.onnx")
    W, a = modelgen.make_voice(p, preset, ns, use_sdp=sdp)
    W2, a2, hdr = load_model(p)
    assert a2 == a
    assert set(W2) == set(W)
    for k in W:
        assert np.array_equal(W[k], W2[k]), k
    assert hdr.metadata["n_speakers"] == str(ns) and hdr.metadata["sample_rate"] == str(a.sample_rate)
    assert ("sid" in hdr.inputs) == (ns > 1)
    # the quirks the loader must undo are really present in the file
    raw = onnx_reader.read_onnx(p)
    anon = [k for k in raw.initializers if k.startswith("onnx::Conv_")]
    assert len(anon) == len(a.flow_layers) * (2 * a.wn_layers + (1 if ns > 1 else 0))
    assert not any(k.startswith("flow.flows.") and ".enc." in k and k.endswith(".weight") for k in raw.initializers)
    if sdp:
        assert "dp.flows.0.logs" not in raw.initializers and any(k.startswith("onnx::Exp_") for k in raw.initializers)


def test_identity_dedup_aliases_are_followed(tmp_path):
    a = modelgen.make_arch("tiny")
    W = modelgen.synth_weights(a, 1)
    # stock random init: every LayerNorm gamma/beta is identical -> exporter de-duplicates them (SURVEY 8a-W ii)
    for k in W:
        if k.endswith(".gamma"):
            W[k] = np.ones_like(W[k])
        if k.endswith(".beta"):
            W[k] = np.zeros_like(W[k])
    p = str(tmp_path / "v.onnx")
    modelgen.write_onnx(W, a, p, dedup_identity=True)
    raw = onnx_reader.read_onnx(p)
    n_ident = sum(1 for n in raw.nodes if n.op_type == "Identity")
    assert n_ident >= 10
    W2, a2, _ = load_model(p)
    assert a2 == a and set(W2) == set(W)
    for k in W:
        assert np.array_equal(W[k], W2[k]), k


def test_state_dict_source_folds_weight_norm():
    a = modelgen.make_arch("tiny")
    W = modelgen.synth_weights(a, 2)
    rs = np.random.RandomState(0)
    sd = {}
    for k, v in W.items():
        if k.startswith("flow.flows.") and ".enc." in k and k.endswith(".weight"):
            vv = rs.randn(*v.shape).astype(np.float32)
            norm = np.sqrt((vv.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
            g = rs.rand(v.shape[0], 1, 1).astype(np.float32) + 0.5
            sd["model_g." + k[:-len("weight")] + "weight_g"] = g
            sd["model_g." + k[:-len("weight")] + "weight_v"] = vv
            W[k] = (g * vv / norm).astype(np.float32)
        else:
            sd["model_g." + k] = v
    sd["model_g.enc_q.pre.weight"] = np.zeros((4, 4, 1), np.float32)      # training-only modules are dropped
    sd["model_d.whatever"] = np.zeros((1,), np.float32)
    W2 = canonical_from_state_dict(sd)
    assert set(W2) == set(W)
    for k in W:
        assert np.allclose(W[k], W2[k], atol=1e-6), k
    assert infer_arch(W2) == a


def test_rejects_non_vits_files(tmp_path):
    p = str(tmp_path / "x.onnx")
    with open(p, "wb") as f:
        f.write(onnx_reader.encode_model({"foo": np.zeros((2, 2), np.float32)}, [], ["input"], ["output"], {}))
    with pytest.raises(ValueError):
        load_model(p)
    with open(p, "wb") as f:
        f.write(b"\x00\x01garbage")
    with pytest.raises(ValueError):
        load_model(p)


def test_gzip_and_plain_are_equivalent(tmp_path, golden_dir):
    gz = os.path.join(golden_dir, "tiny_spk1.onnx.gz")
    plain = str(tmp_path / "t.onnx")
    with gzip.open(gz, "rb") as f, open(plain, "wb") as g:
        g.write(f.read())
    W1, a1, _ = load_model(gz)
    W2, a2, _ = load_model(plain)
    assert a1 == a2 and all(np.array_equal(W1[k], W2[k]) for k in W1)


def test_checkpoint_loader_never_runs_pickled_callables(tmp_path):
    """a voice checkpoint is a pickle; only tensor reconstruction and plain containers may resolve to real objects."""
    import builtins
    import torch
    from phoonnx_b200.weights import load_checkpoint
    a = modelgen.make_arch("tiny")
    W = modelgen.synth_weights(a, 3)

    class Evil:
        def __reduce__(self):
            return (exec, ("import builtins; builtins.PWNED_BY_CHECKPOINT = 1",))

    p = str(tmp_path / "evil.ckpt")
    torch.save({"state_dict": {"model_g." + k: torch.from_numpy(v.copy()) for k, v in W.items()}, "callbacks": Evil(),
                "hyper_parameters": {"x": Evil()}}, p)
    W2, a2, hdr = load_checkpoint(p)
    assert not hasattr(builtins, "PWNED_BY_CHECKPOINT")
    assert a2 == a and set(W2) == set(W) and all(np.array_equal(W[k], W2[k]) for k in W)
    assert hdr.inputs[:3] == ["input", "input_lengths", "scales"]


def test_multilingual_voices_are_refused(tmp_path):
    a = modelgen.make_arch("tiny")
    W = modelgen.synth_weights(a, 1)
    W["emb_l.weight"] = np.zeros((2, 4), np.float32)
    p = str(tmp_path / "ml.onnx")
    modelgen.write_onnx(W, a, p, graph_inputs=["input", "input_lengths", "scales", "langid"])
    with pytest.raises(NotImplementedError):
        load_model(p)


def test_session_feed_follows_the_graph_inputs():
    """voice.py:347-373: the caller filters its feed by session.get_inputs(); a voice exported without `scales` gets none and runs
    at the session's default scales, a single-language voice that declares `langid` accepts (and ignores) it."""
    from phoonnx_b200.session import B200Session, NodeArg
    s = B200Session.__new__(B200Session)
    s.arch = modelgen.make_arch("tiny", 1)
    s._inputs = [NodeArg("input", "tensor(int64)", None), NodeArg("input_lengths", "tensor(int64)", None), NodeArg("langid", "tensor(int64)", None)]
    s._has_scales = False
    s.default_scales = np.asarray((0.5, 1.1, 0.7), np.float32)
    ids = np.array([[1, 0, 5, 0, 2]], np.int64)
    x, lens, scales, sid = s._unpack_feed({"input": ids, "input_lengths": np.array([5], np.int64)})
    assert np.array_equal(scales, s.default_scales) and sid is None
    s._unpack_feed({"input": ids, "input_lengths": np.array([5], np.int64), "langid": np.array([0], np.int64)})
    with pytest.raises(ValueError):
        s._unpack_feed({"input": ids, "input_lengths": np.array([5], np.int64), "langid": np.array([1], np.int64)})
    with pytest.raises(ValueError):
        s._unpack_feed({"input": ids, "input_lengths": np.array([5], np.int64), "scales": np.zeros(3, np.float32)})   # not an input of this graph
