"""Pin the oracle against the REAL reference modules at full preset sizes.  Only runs where the
reference checkout exists (the build container); on the GPU box the golden fixtures stand in."""
import os

import numpy as np
import pytest

from oracle import ref_bridge as rb

pytestmark = pytest.mark.skipif(not rb.reference_available(), reason="reference checkout not present")


@pytest.mark.parametrize("preset,ns", [("medium", 1), ("x_low", 1), ("high", 1), ("medium", 8)])
def test_oracle_equals_reference_full_size(preset, ns, tmp_path):
    from oracle.vits_oracle import VitsOracle
    from phoonnx_b200.weights import canonical_from_state_dict, infer_arch, load_model
    m = rb.build_reference_model(preset, n_speakers=ns)
    W2 = canonical_from_state_dict(m.state_dict())
    if preset == "medium" and ns == 1:
        # through the genuine exporter-format file (anonymous flow weights, folded -logs, ...)
        p = str(tmp_path / "m.onnx")
        rb.export_onnx(m, p, n_speakers=ns)
        W, arch, hdr = load_model(p)
        assert set(W) == set(W2)
        for k in W:
            assert np.allclose(W[k], W2[k], atol=1e-6), k
        assert arch == infer_arch(W2)
        assert (arch.dec_mac_per_frame(), arch.flow_mac_per_frame(), arch.enc_mac_per_id(), arch.dp_mac_per_id()) == \
            (22683648, 7077888, 6266880, 540288)          # SURVEY.md section 8 preset table
    else:
        W, arch = W2, infer_arch(W2)
    orc = VitsOracle(W, arch)
    rs = np.random.RandomState(0)
    T = 33
    ids = rs.randint(0, 256, (T,))
    nd = rs.randn(2, T).astype(np.float32)
    nz = rs.randn(arch.inter, 1500).astype(np.float32)
    sid = 5 if ns > 1 else None
    r = rb.reference_infer(m, ids, (0.667, 1.0, 0.8), sid, nd, nz)
    o = orc.infer(ids, (0.667, 1.0, 0.8), sid, nd, nz)
    assert np.array_equal(r["durations"], o["durations"])
    for k in ("x", "m_p", "logs_p", "logw", "z_p", "z"):
        assert np.abs(r[k] - o[k]).max() <= 3e-5 * max(1.0, np.abs(r[k]).max()), k
    assert np.abs(r["audio"] - o["audio"]).max() <= 1e-5


def test_mac_closed_forms_high_and_xlow():
    from phoonnx_b200.modelgen import make_arch
    assert make_arch("high").dec_mac_per_frame() == 307453952
    assert make_arch("x_low").dec_mac_per_frame() == 22511616
    assert make_arch("x_low").flow_mac_per_frame() == 1769472
    assert make_arch("x_low").enc_mac_per_id() == 1566720
    assert make_arch("x_low").dp_mac_per_id() == 141120


def test_lightning_checkpoint_file_loads_like_the_exported_file(tmp_path):
    """a Lightning ``.ckpt`` of the reference's generator (``model_g.*`` keys, weight-norm pairs still in the coupling
    flow, ``hyper_parameters`` alongside; export_onnx.py:227-245 reads the same file) gives the same canonical tensors and
    architecture as the ONNX file exported from that model."""
    import torch
    from phoonnx_b200.weights import load_model
    m = rb.build_reference_model("x_low", n_speakers=1)
    sd = {"model_g." + k: v.detach().clone() for k, v in m.state_dict().items()}
    sd["model_d.discriminators.0.convs.0.bias"] = torch.zeros(4)            # discriminator weights ride along in real checkpoints
    assert any(k.endswith(".weight_g") for k in sd), "the reference keeps weight-norm in the flow at checkpoint time"
    ck = str(tmp_path / "voice.ckpt")
    torch.save({"state_dict": sd, "hyper_parameters": {"sample_rate": 16000, "quality": "x_low"}, "epoch": 3}, ck)
    onnx_path = str(tmp_path / "voice.onnx")
    rb.export_onnx(m, onnx_path, n_speakers=1)
    Wc, ac, hc = load_model(ck)
    Wo, ao, ho = load_model(onnx_path)
    assert set(Wc) == set(Wo)
    for k in Wo:
        assert np.allclose(Wc[k], Wo[k], atol=1e-6), k
    assert ac == ao and hc.inputs == ho.inputs
