"""Kernel-layout packing (flip folding, interleaved gate, polyphase transposed conv, speaker tables)
verified on CPU: a numpy emulation of the kernels' indexing over the PACKED blobs vs the oracle."""
import numpy as np
import pytest
import torch

import emulate
from oracle.vits_oracle import VitsOracle
from phoonnx_b200 import modelgen, packing


@pytest.mark.parametrize("preset,ns", [("tiny", 1), ("tiny", 3), ("tiny_rb1", 1), ("x_low", 1)])
def test_flow_and_decoder_packing(preset, ns):
    a = modelgen.make_arch(preset, ns)
    W = modelgen.synth_weights(a, 7)
    blobs, opts = packing.pack_model(W, a)
    orc = VitsOracle(W, a)
    rs = np.random.RandomState(0)
    zp = rs.randn(29, a.inter).astype(np.float32)
    sid = 1 if ns > 1 else None
    g = None if sid is None else orc.W["emb_g.weight"][sid][None, :, None]
    z_ref = orc.flow_reverse(torch.from_numpy(zp.T.copy())[None], g)[0].T.numpy()
    assert np.abs(z_ref - emulate.flow_reverse(zp, blobs, a, sid)).max() < 5e-6
    o_ref = orc.decoder(torch.from_numpy(z_ref.T.copy())[None], g)[0, 0].numpy()
    assert np.abs(o_ref - emulate.decoder(z_ref, blobs, a, sid)).max() < 5e-6
    assert opts["dp.ea_m"] == float(W["dp.flows.0.m"][0, 0])


def test_tc_weight_layout_and_bf16_rounding():
    rs = np.random.RandomState(0)
    w = rs.randn(3, 32, 48).astype(np.float32)          # [taps, cin, n]
    out = {}
    packing.pack_conv(out, "c", w, rs.randn(48).astype(np.float32), tc=True)
    wt = out["c.wtc"]
    assert wt.dtype == np.uint16 and wt.shape == (3, 4, 48, 8)
    back = (wt.astype(np.uint32) << 16).view(np.float32)
    # [tap][ci/8][n][8] -> element (tap, ci, n)
    assert np.array_equal(back[1, 2, 5, 3], packing.bf16_round(w)[1, 2 * 8 + 3, 5])
    # round-to-nearest-even against torch
    x = rs.randn(1000).astype(np.float32)
    assert np.array_equal(packing.bf16_round(x), torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy())
    # N not a multiple of 16 / cin not a multiple of 16 -> no tensor-core blob
    out2 = {}
    packing.pack_conv(out2, "d", rs.randn(1, 32, 29).astype(np.float32), None, tc=True)
    assert "d.wtc" not in out2 and out2["d.w"].shape == (1, 32, 32)


def test_bf16x3_operands_reproduce_fp32_products():
    """The text side runs x*w as xh*wh + xh*wl + xl*wh on the tensor cores (csrc/conv_tc.cuh split3); the packed
    operand "<n>.wtc3.<j>" must hold [wh | wl | wh] per K slice, and the three-term sum must be fp32-faithful."""
    from phoonnx_b200 import packing
    rs = np.random.RandomState(0)
    taps, cin, n = 3, 384, 32
    w = (rs.randn(taps, cin, n) / np.sqrt(cin * taps)).astype(np.float32)
    blobs = {}
    packing.pack_conv(blobs, "c", w, None, tc3=True)
    sl = packing.split3_slice(cin)
    nsl = cin // sl
    assert sl == packing.SPLIT3_MAX_CIN and cin % sl == 0 and f"c.wtc3.{nsl - 1}" in blobs and f"c.wtc3.{nsl}" not in blobs
    x = rs.randn(50, cin).astype(np.float32)
    xh = packing.bf16_round(x)
    xl = packing.bf16_round(x - xh)
    acc = np.zeros((50, n), np.float64)
    for j in range(cin // sl):
        blob = blobs[f"c.wtc3.{j}"]                                   # [tap][3*sl/8][n16][8] bf16 bits
        wk = (blob.astype(np.uint32) << 16).view(np.float32).transpose(0, 1, 3, 2).reshape(taps, 3 * sl, -1)[:, :, :n]
        xs = np.concatenate([xh[:, j * sl:(j + 1) * sl], xh[:, j * sl:(j + 1) * sl], xl[:, j * sl:(j + 1) * sl]], axis=1)
        acc += xs.astype(np.float64) @ wk[1].astype(np.float64)       # centre tap only: a plain GEMM
    want = x.astype(np.float64) @ w[1].astype(np.float64)
    assert np.abs(acc - want).max() < 2e-5 * np.abs(want).max() + 1e-6
    plain = xh.astype(np.float64) @ packing.bf16_round(w[1]).astype(np.float64)
    assert np.abs(plain - want).max() > 50 * np.abs(acc - want).max()   # the split really buys ~2^8


def test_flow_tail_composition_matches_skip_then_post():
    """packing.py "mskip": m = post(sum_l skip_l(acts_l)) folded into one GEMM over the concatenated gate outputs,
    res_skip reduced to its residual half (modules.py:196-209, 447-466).  Checked against the unfused blobs."""
    import tempfile, os
    from phoonnx_b200 import modelgen
    from phoonnx_b200.weights import load_model
    from phoonnx_b200.packing import pack_model
    import emulate as E
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "v.onnx")
        modelgen.make_voice(path, "x_low", n_speakers=1, seed=3)
        W, a, _ = load_model(path)
    blobs, _ = pack_model(W, a)
    H, half, L = a.hidden, a.inter // 2, a.wn_layers
    rs = np.random.RandomState(0)
    acts = [rs.randn(40, H).astype(np.float32) for _ in range(L)]
    for s in range(len(a.flow_layers)):
        skip = np.zeros((40, H), np.float64)
        for i in range(L):
            w, b = blobs[f"flow.{s}.rs.{i}.w"], blobs[f"flow.{s}.rs.{i}.b"]
            full = E.conv_cl(acts[i], w, b, [0], w.shape[2])
            if i < L - 1:
                skip += full[:, H:2 * H]
                res = E.conv_cl(acts[i], blobs[f"flow.{s}.rsr.{i}.w"], blobs[f"flow.{s}.rsr.{i}.b"], [0], H)
                assert np.array_equal(res, full[:, :H])
            else:
                skip += full[:, :H]
        want = E.conv_cl(skip.astype(np.float32), blobs[f"flow.{s}.post.w"], blobs[f"flow.{s}.post.b"], [0], half)
        got = sum(acts[i].astype(np.float64) @ blobs[f"flow.{s}.mskip.{i}.w"][0, :, :half].astype(np.float64) for i in range(L))
        got = got + blobs[f"flow.{s}.mskip.b"][:half]
        assert np.abs(got - want).max() < 2e-5 * max(1.0, np.abs(want).max()), np.abs(got - want).max()


def test_summed_second_convs_algebra():
    """The unfused ResBlock2 stage runs `sum_r [ x1_r + conv_{k_r,d_r2}(lrelu x1_r) ] / n_r` as ONE launch over the rows
    [lrelu(x1_0) | lrelu(x1_1) | lrelu(x1_2)] (K slices with per-slice taps, residual recovered by inverting the leaky-relu;
    csrc/conv_tc.cuh, engine.cu `stage_sum2`).  Same numbers as modules.py:355-364 evaluated resblock by resblock."""
    import emulate as E
    rs = np.random.RandomState(3)
    L, C, slope = 300, 16, 0.1
    ks, dils = (3, 5, 7), ((1, 2), (2, 6), (3, 12))
    x = rs.randn(L, C).astype(np.float32)
    w1 = [rs.randn(k, C, C).astype(np.float32) / np.sqrt(k * C) for k in ks]
    w2 = [rs.randn(k, C, C).astype(np.float32) / np.sqrt(k * C) for k in ks]
    b1 = [rs.randn(C).astype(np.float32) * 0.1 for _ in ks]
    b2 = [rs.randn(C).astype(np.float32) * 0.1 for _ in ks]
    # reference order: resblock by resblock
    ref = np.zeros((L, C), np.float64)
    x1s = []
    for r, k in enumerate(ks):
        x1 = x + E.conv_cl(E.lrelu(x, slope), w1[r], b1[r], E.sym_taps(k, dils[r][0]), C)
        x1s.append(x1)
        ref += x1 + E.conv_cl(E.lrelu(x1, slope), w2[r], b2[r], E.sym_taps(k, dils[r][1]), C)
    ref /= len(ks)
    # summed launch: operand rows side by side, one accumulator, flattened tap list
    rows = np.concatenate([E.lrelu(x1, slope) for x1 in x1s], axis=1)                 # [L, 3C]
    acc = np.zeros((L, C), np.float64)
    for r, k in enumerate(ks):
        acc += E.conv_cl(rows[:, r * C:(r + 1) * C], w2[r], None, E.sym_taps(k, dils[r][1]), C)
    res = np.zeros((L, C), np.float64)
    for r in range(len(ks)):
        v = rows[:, r * C:(r + 1) * C].astype(np.float64)
        res += np.minimum(v, v / slope)                                               # lrelu^-1: x = min(v, v / slope)
    out = (acc + res + np.sum(b2, axis=0)) / len(ks)
    assert np.abs(out - ref).max() < 1e-5
