"""Parity tests proper (run with -m gpu on a B200): the CUDA path, called through the C ABI
(phoonnx_b200.engine -> libvits_b200.so), against (a) the golden fixtures minted from the real
reference and (b) the CPU oracle on the same seeded inputs and injected noise.

Bars (BASELINE.json north_star): integer durations / cumsum / alignment path bit-exact given
identical logw; fp32 mode waveform max-abs <= 1e-3; bf16 tensor-core mode SNR >= 40 dB."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(ROOT, "tools"))

FP32_TOL = 1e-3          # stated tolerance (north_star); observed ~1e-6
STAGE_TOL = 5e-5         # our own tighter bar on intermediate tensors in fp32 mode
BF16_SNR_DB = 40.0
SCALES = np.array([0.667, 1.0, 0.8], np.float32)


def snr_db(ref, got):
    return 10 * np.log10(float((ref.astype(np.float64) ** 2).sum()) / max(float(((got - ref).astype(np.float64) ** 2).sum()), 1e-30))


@pytest.fixture(scope="module")
def lib(built_lib):
    return built_lib


# ------------------------------------------------------------------------------------------ conv kernels
@pytest.mark.parametrize("use_tc", [0, 1, 2])
def test_conv_kernels_vs_numpy(lib, use_tc):
    """use_tc 0: fp32 CUDA-core kernel; 1: tcgen05 bf16 (vs bf16-rounded operands); 2: tcgen05 bf16x3 (hi/lo split,
    fp32-faithful: compared against the UNROUNDED fp32 reference)."""
    import gpu_diag as gd
    eng = gd.mini_engine()
    rs = np.random.RandomState(0)
    for (L, cin, n, taps, kw) in gd.CASES:
        kw = dict(kw)
        if use_tc and (cin % 16 or n % 16):
            continue
        if use_tc == 2 and (kw.get("epi", 0) != 0 or cin > 192):
            continue
        x = rs.randn(L, cin).astype(np.float32)
        w = (rs.randn(len(taps), cin, n) / np.sqrt(cin * len(taps))).astype(np.float32)
        b = rs.randn(n).astype(np.float32)
        res = rs.randn(L, n).astype(np.float32) if kw.pop("with_res", False) else None
        oi = rs.randn(L, n // 2 if kw.get("epi") == 1 else n).astype(np.float32) if kw.get("accumulate") else None
        got = gd.run_conv(eng, use_tc, x, w, b, taps, res=res, out_init=oi, **kw)
        want = gd.ref_conv(x, w, b, taps, res=res, out_init=oi, bf16=(use_tc == 1), **kw)
        assert np.abs(got - want).max() < (2e-4 if use_tc == 2 else 5e-5), (L, cin, n, taps, kw)


# ------------------------------------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize("name", ["tiny_spk1", "tiny_spk3", "tiny_rb1_spk1"])
def test_golden_reference_outputs(lib, golden_dir, name):
    """Genuine exporter-format file -> loader -> engine, vs the reference's own infer() outputs."""
    from phoonnx_b200.session import B200Session
    sess = B200Session(os.path.join(golden_dir, name + ".onnx.gz"), precision="fp32")
    sess.engine.set_option("debug_keep_zp", 1)
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    arch = sess.arch
    for u in range(6):
        for tag in ("n", "z"):
            k = f"u{u}{tag}_"
            ids = g[k + "ids"][None].astype(np.int64)
            T = ids.shape[1]
            feed = {"input": ids, "input_lengths": np.array([T], np.int64), "scales": g[k + "scales"].astype(np.float32)}
            if arch.n_speakers > 1:
                feed["sid"] = np.array([int(g[k + "sid"])], np.int64)
            if tag == "n":
                feed["noise_dp"] = g[k + "noise_dp"][None]
                feed["noise_z"] = g[k + "noise_z"][None]
            out = sess.run(None, feed)[0]
            assert out.dtype == np.float32 and out.ndim == 4 and out.shape[:3] == (1, 1, 1)
            eng = sess.engine
            assert np.array_equal(eng.fetch("durations"), g[k + "durations"]), (name, k)
            assert np.abs(eng.fetch("logw") - g[k + "logw"]).max() < STAGE_TOL
            assert np.abs(eng.fetch("x").reshape(T, -1) - g[k + "x"]).max() < STAGE_TOL
            st = eng.fetch("stats").reshape(T, -1)
            assert np.abs(st[:, :arch.inter] - g[k + "m_p"]).max() < STAGE_TOL
            assert np.abs(st[:, arch.inter:] - g[k + "logs_p"]).max() < STAGE_TOL
            assert np.abs(eng.fetch("z_p").reshape(-1, arch.inter) - g[k + "z_p"]).max() < STAGE_TOL
            assert np.abs(eng.fetch("z").reshape(-1, arch.inter) - g[k + "z"]).max() < STAGE_TOL
            audio = out[0, 0, 0]
            assert audio.shape == g[k + "audio"].shape
            assert np.abs(audio - g[k + "audio"]).max() <= FP32_TOL
            assert np.abs(audio - g[k + "audio"]).max() <= 1e-5     # observed bar


# ------------------------------------------------------------------------------------------ presets vs oracle
def _voice(tmp_path_factory, preset, ns, sdp=True, seed=5):
    from phoonnx_b200 import modelgen
    from phoonnx_b200.weights import load_model
    from oracle.vits_oracle import VitsOracle
    p = str(tmp_path_factory.mktemp("voice") / f"{preset}_{ns}.onnx")
    modelgen.make_voice(p, preset, ns, seed=seed, use_sdp=sdp)
    W, arch, _ = load_model(p)
    return p, arch, VitsOracle(W, arch)


def _compare(sess, orc, arch, lens, ns, precision, rs):
    B, T = len(lens), int(max(lens))
    lens = np.asarray(lens, np.int64)
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    nd = rs.randn(B, 2, T).astype(np.float32)
    nz = rs.randn(B, arch.inter, 16 * T + 64).astype(np.float32)
    feed = {"input": ids, "input_lengths": lens, "scales": SCALES, "noise_dp": nd, "noise_z": nz}
    sid = None
    if ns > 1:
        sid = (np.arange(B) % ns).astype(np.int64)
        feed["sid"] = sid
    audio, alen = sess.synthesize_packed(feed)
    dur = sess.engine.fetch("durations")
    logw = sess.engine.fetch("logw")
    off = aoff = 0
    worst_err, worst_snr = 0.0, 1e9
    for b in range(B):
        L = int(lens[b])
        r = orc.infer(ids[b, :L], SCALES, None if sid is None else int(sid[b]), nd[b][:, :L], nz[b], stages=True)
        assert np.abs(logw[off:off + L] - r["logw"]).max() < 1e-4
        if not np.array_equal(dur[off:off + L], r["durations"]):
            # a ceil() tie: only legal where exp(logw)*length_scale is within a few ulp of an integer
            w = np.exp(r["logw"].astype(np.float64))
            bad = np.nonzero(dur[off:off + L] != r["durations"])[0]
            assert all(abs(w[i] - round(w[i])) < 1e-4 for i in bad), "duration mismatch that is not a ceil tie"
            r = orc.infer(ids[b, :L], SCALES, None if sid is None else int(sid[b]), nd[b][:, :L], nz[b],
                          logw_override=logw[off:off + L])
            assert np.array_equal(dur[off:off + L], r["durations"])
        a_g = audio[aoff:aoff + int(alen[b])]
        assert a_g.shape == r["audio"].shape
        worst_err = max(worst_err, float(np.abs(a_g - r["audio"]).max()))
        worst_snr = min(worst_snr, snr_db(r["audio"], a_g))
        off += L
        aoff += int(alen[b])
    return worst_err, worst_snr


@pytest.mark.parametrize("preset,ns,sdp,lens", [
    ("tiny", 1, True, [37, 5, 64, 1, 23, 2, 3, 6]),
    ("tiny", 1, False, [9, 30]),
    ("x_low", 1, True, [64, 17, 120]),
    ("medium", 1, True, [128, 40]),
    ("medium", 8, True, [50, 77, 64, 12]),
    ("high", 1, True, [96, 20]),
])
def test_fp32_mode_matches_oracle(lib, tmp_path_factory, preset, ns, sdp, lens):
    from phoonnx_b200.session import B200Session
    p, arch, orc = _voice(tmp_path_factory, preset, ns, sdp)
    sess = B200Session(p, precision="fp32")
    err, snr = _compare(sess, orc, arch, lens, ns, "fp32", np.random.RandomState(3))
    assert err <= FP32_TOL, err
    assert err <= 2e-5, err            # observed bar, so regressions show
    assert snr >= 100.0


@pytest.mark.parametrize("preset,ns,lens", [
    ("x_low", 1, [64, 17, 120]),
    ("medium", 1, [128, 40]),
    ("medium", 8, [50, 77, 64, 12]),
    ("high", 1, [96, 20]),
])
def test_bf16_tensor_core_mode_snr(lib, tmp_path_factory, preset, ns, lens):
    from phoonnx_b200.session import B200Session
    p, arch, orc = _voice(tmp_path_factory, preset, ns)
    sess = B200Session(p, precision="bf16")
    n0 = sess.engine.launch_count()
    err, snr = _compare(sess, orc, arch, lens, ns, "bf16", np.random.RandomState(4))
    assert snr >= BF16_SNR_DB, snr
    assert sess.engine.launch_count() > n0


# ------------------------------------------------------------------------------------------ integer path
def test_integer_path_bit_exact_given_logw(lib, tmp_path_factory):
    """durations, cumsum and frame->id index are exact given identical logw, incl. zero durations,
    all-zero utterances (y_len clamps to 1), ragged lengths, long utterances (multi-chunk scan)."""
    import torch
    from phoonnx_b200.session import B200Session
    from oracle.vits_oracle import VitsOracle
    p, arch, orc = _voice(tmp_path_factory, "tiny", 1)
    sess = B200Session(p, precision="fp32")
    rs = np.random.RandomState(7)
    lens = np.array([1, 2, 300, 17, 256, 257, 700, 5], np.int64)
    B, T = len(lens), int(lens.max())
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    for length_scale in (1.0, 0.37, 2.5):
        logw = (rs.randn(int(lens.sum())) * 1.2 + 0.3).astype(np.float32)
        logw[rs.rand(logw.size) < 0.1] = -120.0           # exp -> 0 -> duration 0
        o5 = int(lens[:7].sum())
        logw[o5:o5 + 5] = -120.0                          # utterance 7: every duration 0
        feed = {"input": ids, "input_lengths": lens, "scales": np.array([0.0, length_scale, 0.0], np.float32), "logw": logw}
        audio, alen = sess.synthesize_packed(feed)
        dur = sess.engine.fetch("durations")
        cum = sess.engine.fetch("cum")
        off = 0
        for b in range(B):
            L = int(lens[b])
            d, y = VitsOracle.durations_from_logw(torch.from_numpy(logw[off:off + L]), length_scale)
            assert np.array_equal(dur[off:off + L], d.numpy()), (b, length_scale)
            assert np.array_equal(cum[off:off + L], np.cumsum(d.numpy())), b
            assert int(alen[b]) == y * arch.hop
            off += L
        assert int(alen[7]) == arch.hop                   # clamp_min(sum, 1) (models.py:704)
    # frame index of the last chunk == searchsorted (single-utterance call so the chunk is the utterance)
    L = 300
    logw = (rs.randn(L) * 1.0 + 0.5).astype(np.float32)
    logw[::7] = -120.0
    feed = {"input": ids[2:3, :L], "input_lengths": np.array([L], np.int64), "scales": np.array([0.0, 1.0, 0.0], np.float32), "logw": logw}
    sess.synthesize_packed(feed)
    d, y = VitsOracle.durations_from_logw(torch.from_numpy(logw), 1.0)
    assert np.array_equal(sess.engine.fetch("frame_index"), VitsOracle.frame_index(d, y).numpy())


def test_batch_equals_single_and_chunking_is_invisible(lib, tmp_path_factory):
    """Per-utterance (B=1) semantics: an utterance's audio does not depend on its batch mates nor on
    how the frame side is chunked."""
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, "x_low", 1)
    sess = B200Session(p, precision="fp32")
    rs = np.random.RandomState(11)
    lens = np.array([40, 9, 77, 23], np.int64)
    B, T = 4, 77
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    nd = rs.randn(B, 2, T).astype(np.float32)
    nz = rs.randn(B, arch.inter, 1400).astype(np.float32)
    feed = {"input": ids, "input_lengths": lens, "scales": SCALES, "noise_dp": nd, "noise_z": nz}
    full = sess.run(None, feed)[0]
    assert full.shape[0] == B and full.shape[1:3] == (1, 1)
    alen = sess.last_lengths
    for b in range(B):
        L = int(lens[b])
        one = sess.run(None, {"input": ids[b:b + 1, :L], "input_lengths": lens[b:b + 1], "scales": SCALES,
                              "noise_dp": nd[b:b + 1, :, :L], "noise_z": nz[b:b + 1]})[0]
        assert one.shape[-1] == alen[b]
        assert np.array_equal(one[0, 0, 0], full[b, 0, 0, :alen[b]])
        assert not full[b, 0, 0, alen[b]:].any()          # zero beyond the utterance's own length
    sess.engine.set_option("max_chunk_frames", 64)        # forces one chunk per utterance
    again = sess.run(None, feed)[0]
    assert np.array_equal(again, full)
    for prec in ("bf16",):
        s2 = B200Session(p, precision=prec)
        a = s2.run(None, feed)[0]
        s2.engine.set_option("max_chunk_frames", 64)
        b_ = s2.run(None, feed)[0]
        assert np.array_equal(a, b_)


def test_int16_postprocessing_exact(lib, tmp_path_factory):
    """Device-side normalise/volume/clip/x32767 (voice.py:271-282, 88-91) == numpy, bit for bit."""
    from phoonnx_b200.session import B200Session
    from oracle.vits_oracle import postprocess_int16
    p, arch, _ = _voice(tmp_path_factory, "tiny", 1)
    sess = B200Session(p, precision="fp32", seed=3)
    rs = np.random.RandomState(5)
    lens = np.array([30, 11], np.int64)
    ids = rs.randint(0, arch.n_vocab, (2, 30)).astype(np.int64)
    nd = rs.randn(2, 2, 30).astype(np.float32)
    nz = rs.randn(2, arch.inter, 600).astype(np.float32)
    feed = {"input": ids, "input_lengths": lens, "scales": SCALES, "noise_dp": nd, "noise_z": nz}
    f32, alen = sess.synthesize_packed(feed)
    for vol, norm in ((1.0, True), (0.5, True), (3.0, False)):
        i16, _ = sess.synthesize_packed(feed, out="i16", volume=vol, normalize=norm)
        off = 0
        for b in range(2):
            n = int(alen[b])
            assert np.array_equal(i16[off:off + n], postprocess_int16(f32[off:off + n], vol, norm)), (vol, norm, b)
            off += n


def test_error_behaviour(lib, tmp_path_factory):
    """Bad feeds raise ValueError like ORT's InvalidArgument would; nothing is silently clamped."""
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, "tiny", 3)
    sess = B200Session(p)
    assert [i.name for i in sess.get_inputs()] == ["input", "input_lengths", "scales", "sid"]
    assert [o.name for o in sess.get_outputs()] == ["output"]
    good = {"input": np.array([[1, 0, 20, 0, 2]], np.int64), "input_lengths": np.array([5], np.int64),
            "scales": SCALES, "sid": np.array([2], np.int64)}
    out = sess.run(None, good)[0]
    assert out.ndim == 4 and out.shape[-1] % arch.hop == 0 and out.squeeze().ndim == 1
    for key, val in (("input", np.array([[1, 0, 999, 0, 2]], np.int64)), ("sid", np.array([3], np.int64)),
                     ("input", np.array([[1, 0, 20, 0, 2]], np.int32)), ("input_lengths", np.array([9], np.int64)),
                     ("scales", np.array([0.6, 1.0], np.float32)), ("input_lengths", np.array([0], np.int64))):
        bad = dict(good)
        bad[key] = val
        with pytest.raises(ValueError):
            sess.run(None, bad)
    with pytest.raises(ValueError):
        sess.run(None, {k: v for k, v in good.items() if k != "sid"})
    with pytest.raises(ValueError):
        sess.run(None, dict(good, langid=np.array([0], np.int64)))
    sess.run(None, good)       # the handle is still usable after errors


def test_device_noise_is_seeded_and_scaled(lib, tmp_path_factory):
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, "tiny", 1)
    feed = {"input": np.arange(40, dtype=np.int64)[None] % 200, "input_lengths": np.array([40], np.int64), "scales": SCALES}
    a = B200Session(p, seed=1).run(None, feed)[0]
    b = B200Session(p, seed=1).run(None, feed)[0]
    c = B200Session(p, seed=2).run(None, feed)[0]
    assert np.array_equal(a, b)
    assert a.shape != c.shape or not np.array_equal(a, c)
    z = dict(feed, scales=np.array([0.0, 1.0, 0.0], np.float32))
    s = B200Session(p, seed=9)
    assert np.array_equal(s.run(None, z)[0], s.run(None, z)[0])    # zero noise: deterministic


def test_full_size_properties(lib, tmp_path_factory):
    """BASELINE config 3 shape (medium, 8 speakers, batch 64, 64-256 ids) through size-independent
    properties: audio length == hop * sum(durations), finite, tanh-bounded, batch == singles on a probe."""
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, "medium", 8)
    sess = B200Session(p, precision="bf16", seed=5)
    rs = np.random.RandomState(1)
    lens = rs.randint(64, 257, size=(64,)).astype(np.int64)
    ids = rs.randint(0, arch.n_vocab, (64, int(lens.max()))).astype(np.int64)
    feed = {"input": ids, "input_lengths": lens, "scales": SCALES, "sid": (np.arange(64) % 8).astype(np.int64)}
    audio, alen = sess.synthesize_packed(feed)
    dur = sess.engine.fetch("durations")
    off = 0
    for b in range(64):
        assert int(alen[b]) == arch.hop * max(int(dur[off:off + lens[b]].sum()), 1)
        off += int(lens[b])
    assert audio.shape[0] == int(alen.sum()) and np.isfinite(audio).all() and np.abs(audio).max() <= 1.0
    assert 2.0 < dur.mean() < 6.0


@pytest.mark.parametrize("preset", ["x_low", "medium"])
def test_fused_mrf_stage_equals_unfused(lib, tmp_path_factory, preset):
    """mrf3_tc.cuh (a whole multi-receptive-field stage in ONE kernel, the last one with its ConvTranspose1d and conv_post) vs the
    conv-by-conv tcgen05 path on fp32 rows (`no_fused_mrf` + `no_stage_bf16`): the fused kernel keeps inter-stage rows as bf16 lrelu
    operands and recovers the residual from them (one extra bf16 rounding per stage), so the two agree to bf16 noise; tile height,
    issue order and where the ConvTranspose / conv_post run do not move a single rounding point.  Ragged tile edges included."""
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, preset, 1)
    rs = np.random.RandomState(21)
    lens = np.array([97, 3, 160, 41, 1], np.int64)
    B, T = len(lens), int(lens.max())
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    nd = rs.randn(B, 2, T).astype(np.float32)
    nz = rs.randn(B, arch.inter, 2600).astype(np.float32)
    feed = {"input": ids, "input_lengths": lens, "scales": SCALES, "noise_dp": nd, "noise_z": nz}
    plain = B200Session(p, precision="bf16")
    plain.engine.set_option("no_fused_mrf", 1)
    plain.engine.set_option("no_stage_bf16", 1)   # fp32 rows between the per-conv launches: the reference for the fused kernel
    n0 = plain.engine.launch_count()
    b, blen = plain.synthesize_packed(feed)
    b = np.array(b)
    n_plain = plain.engine.launch_count() - n0

    fused = B200Session(p, precision="bf16")
    n0 = fused.engine.launch_count()
    d, dlen = fused.synthesize_packed(feed)
    d = np.array(d)
    n_fused = fused.engine.launch_count() - n0
    assert np.array_equal(dlen, blen)
    assert n_fused < n_plain                      # the fused path really ran
    assert snr_db(b, d) > 45.0, snr_db(b, d)
    for opts in ({"mrf_nb": 1}, {"mrf_nb": 2}, {"mrf_interleave": 1}, {"no_fused_ups": 1}, {"no_fused_post": 1},
                 {"no_fused_ups": 1, "no_fused_post": 1}):
        alt = B200Session(p, precision="bf16")
        for k, v in opts.items():
            alt.engine.set_option(k, v)
        c, clen = alt.synthesize_packed(feed)
        c = np.array(c)
        assert np.array_equal(dlen, clen)
        assert snr_db(b, c) > 45.0, (opts, snr_db(b, c))
        if "no_fused_post" not in opts:
            assert np.abs(d - c).max() < 1e-4, (opts, np.abs(d - c).max())
    # the conv-by-conv path on bf16 operand rows (what a stage that does not qualify for the fused kernel runs)
    rows = B200Session(p, precision="bf16")
    rows.engine.set_option("no_fused_mrf", 1)
    e, elen = rows.synthesize_packed(feed)
    assert np.array_equal(elen, blen) and snr_db(b, np.array(e)) > 45.0


def test_synthesize_many_equals_serial_calls(lib, tmp_path_factory):
    """The pipelined batch call (page-locked results, device->host transfer of batch k under the kernels of batch
    k+1, alternating device audio buffers) returns exactly what the serial per-batch calls return."""
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, "x_low", 1)
    rs = np.random.RandomState(5)
    feeds = []
    for B in (3, 1, 4, 2, 5):
        lens = rs.randint(5, 60, size=(B,)).astype(np.int64)
        ids = rs.randint(0, arch.n_vocab, (B, int(lens.max()))).astype(np.int64)
        feeds.append({"input": ids, "input_lengths": lens, "scales": SCALES})
    for prec in ("fp32", "bf16"):
        a = B200Session(p, precision=prec, seed=77)
        serial = [a.synthesize_packed(f) for f in feeds]
        serial = [(x.copy(), n.copy()) for x, n in serial]
        b = B200Session(p, precision=prec, seed=77)
        b.engine.set_option("max_chunk_frames", 96)          # several chunks per batch: per-chunk transfers
        many = list(b.synthesize_many(feeds))
        assert len(many) == len(serial)
        for (x0, n0), (x1, n1) in zip(serial, many):
            assert np.array_equal(n0, n1)
            assert x0.shape == x1.shape and np.array_equal(x0, x1)
        # the session is back in blocking mode afterwards and page-locked blocks are recycled
        again, alen = b.synthesize_packed(feeds[0])              # (a new call number => new noise, new durations)
        assert again.dtype == np.float32 and again.shape == (int(alen.sum()),) and np.isfinite(again).all()


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["x_low", "medium"])
def test_tensor_core_attention_matches_fp32_kernel(lib, tmp_path_factory, preset):
    """attention.cuh: the mma.sync bf16x3 kernel (bf16 mode) vs the fp32 CUDA-core kernel on the same qkv --
    every key-tile count (T = 1 .. 256), ragged q tiles, utterances shorter than the relative window
    (attentions.py:295-305); logw agrees to fp32 noise, so durations are identical (or exact ceil ties)."""
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, preset, 1)
    rs = np.random.RandomState(33)
    lens = np.array([256, 193, 129, 128, 65, 64, 63, 6, 5, 3, 1], np.int64)
    B, T = len(lens), int(lens.max())
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    nd = rs.randn(B, 2, T).astype(np.float32)
    feed = {"input": ids, "input_lengths": lens, "scales": SCALES, "noise_dp": nd}
    res = {}
    for name, opt in (("mma", 0), ("fp32", 1)):
        sess = B200Session(p, precision="bf16")
        sess.engine.set_option("attention_fp32", opt)
        n0 = sess.engine.launch_count()
        sess.synthesize_packed(feed, out="none")
        assert sess.engine.launch_count() > n0
        res[name] = (sess.engine.fetch("logw").copy(), sess.engine.fetch("durations").copy())
    d = np.abs(res["mma"][0] - res["fp32"][0]).max()
    assert d < 3e-5, d
    bad = np.nonzero(res["mma"][1] != res["fp32"][1])[0]
    w = np.exp(res["fp32"][0].astype(np.float64))
    assert all(abs(w[i] - round(w[i])) < 1e-4 for i in bad), "duration mismatch that is not a ceil tie"


@pytest.mark.parametrize("preset", ["x_low", "medium"])
def test_summed_second_convs_equal_per_conv_launches(lib, tmp_path_factory, preset):
    """Unfused ResBlock2 stage (the 128-channel first stage): one launch that sums the second convs of all resblocks in TMEM
    (K slices with per-slice taps, residual = sum of the bf16 x1 rows) vs one launch per conv with fp32 accumulation of the
    stage output -- same operand rounding, so they agree to fp32 re-association noise; ragged utterance edges included."""
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, preset, 1)
    rs = np.random.RandomState(5)
    lens = np.array([97, 3, 160, 41, 1], np.int64)
    B, T = len(lens), int(lens.max())
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    nd = rs.randn(B, 2, T).astype(np.float32)
    nz = rs.randn(B, arch.inter, 2600).astype(np.float32)
    feed = {"input": ids, "input_lengths": lens, "scales": SCALES, "noise_dp": nd, "noise_z": nz}
    outs = []
    for opt in (0, 1):
        sess = B200Session(p, precision="bf16")
        sess.engine.set_option("no_stage_sum2", opt)
        n0 = sess.engine.launch_count()
        a, alen = sess.synthesize_packed(feed)
        outs.append((a, alen, sess.engine.launch_count() - n0))
    assert np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] < outs[1][2]               # fewer launches: the summed path really ran
    assert snr_db(outs[1][0], outs[0][0]) > 55.0, snr_db(outs[1][0], outs[0][0])


@pytest.mark.parametrize("preset,B,tlo,thi,sr", [("x_low", 32, 64, 257, 16000), ("high", 16, 512, 513, 22050)])
def test_baseline_configs_2_and_4_full_size(lib, tmp_path_factory, preset, B, tlo, thi, sr):
    """BASELINE configs[1] (x_low, batch of 32 utterances of 64-256 ids) and configs[3] (high = ResBlock1, kernels 3/7/11, batch
    16 x 512 ids: 8 attention key tiles, multi-chunk decode) at full size through size-independent properties: audio length ==
    hop * sum(durations), finite, tanh-bounded, and one utterance of the batch re-run alone (B = 1 semantics, voice.py:350)."""
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, preset, 1)
    sess = B200Session(p, precision="bf16", seed=9, max_chunk_frames=8192)
    rs = np.random.RandomState(2)
    lens = rs.randint(tlo, thi, size=(B,)).astype(np.int64)
    T = int(lens.max())
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    nd = rs.randn(B, 2, T).astype(np.float32)
    feed = {"input": ids, "input_lengths": lens, "scales": SCALES, "noise_dp": nd}
    audio, alen = sess.synthesize_packed(feed)
    dur = sess.engine.fetch("durations").copy()
    off = 0
    for b in range(B):
        assert int(alen[b]) == arch.hop * max(int(dur[off:off + lens[b]].sum()), 1)
        off += int(lens[b])
    assert audio.shape[0] == int(alen.sum()) and np.isfinite(audio).all() and np.abs(audio).max() <= 1.0
    # the same utterance alone: identical durations (the text side is per-utterance and deterministic given noise_dp)
    b = B // 2
    one = B200Session(p, precision="bf16", seed=9)
    L = int(lens[b])
    _, alen1 = one.synthesize_packed({"input": ids[b:b + 1, :L], "input_lengths": lens[b:b + 1], "scales": SCALES, "noise_dp": nd[b:b + 1, :, :L]})
    d1 = one.engine.fetch("durations")
    o = int(lens[:b].sum())
    assert np.array_equal(d1[:L], dur[o:o + L]) and int(alen1[0]) == int(alen[b])


def test_cluster_weight_multicast_is_bit_identical(lib, tmp_path_factory):
    """Optional 2-CTA cluster mode of the conv kernel (`conv_cluster`): CTA pairs share the weight ring through multicast bulk
    copies and multicast tcgen05.commit.  Same MMAs in the same order -> the audio is bit-identical to the default launches; odd
    tile counts exercise the round where only one CTA of a pair has a tile."""
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, "medium", 1)
    rs = np.random.RandomState(8)
    lens = np.array([97, 3, 160, 41, 1, 77, 130], np.int64)
    B, T = len(lens), int(lens.max())
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    nd = rs.randn(B, 2, T).astype(np.float32)
    nz = rs.randn(B, arch.inter, 2600).astype(np.float32)
    feed = {"input": ids, "input_lengths": lens, "scales": SCALES, "noise_dp": nd, "noise_z": nz}
    outs = []
    for opt in (0, 1):
        sess = B200Session(p, precision="bf16")
        sess.engine.set_option("conv_cluster", opt)
        a, alen = sess.synthesize_packed(feed)
        outs.append((a.copy(), alen.copy()))
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.isfinite(outs[1][0]).all() and np.array_equal(outs[0][0], outs[1][0])


@pytest.mark.parametrize("preset,option", [("medium", "mrf_tma"), ("x_low", "mrf_tma"), ("medium", "conv_tma"), ("high", "conv_tma"), ("medium", "pdl")])
def test_tma_fed_input_tiles_are_bit_identical(lib, tmp_path_factory, preset, option):
    """The fused stage kernels (`mrf_tma`) and every conv_tc launch on bf16 operand rows (`conv_tma`: coupling flow, ConvTranspose,
    unfused stages) take their activation tiles by TMA (cp.async.bulk.tensor boxes, default) or through cp.async loader warps
    (option = 0).  Same operand bytes in shared memory -> bit-identical audio.  `pdl`: every kernel launched with the programmatic-
    dependent-launch attribute (default) or plainly (0): only launch latency may overlap, never data.  The lengths put utterance starts and ends inside
    tiles (rows of the neighbouring utterance must be zeroed after the TMA delivered them), the first utterance at the start of the
    array and the last at its end (rows outside the array: zeros from the TMA unit), and a one-id utterance shorter than any halo."""
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, preset, 1)
    rs = np.random.RandomState(12)
    lens = np.array([1, 97, 3, 160, 41, 2, 77, 130, 5], np.int64)
    B, T = len(lens), int(lens.max())
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    nd = rs.randn(B, 2, T).astype(np.float32)
    nz = rs.randn(B, arch.inter, 2600).astype(np.float32)
    feed = {"input": ids, "input_lengths": lens, "scales": SCALES, "noise_dp": nd, "noise_z": nz}
    outs = []
    for opt in (0, 1):
        for chunk in (None, 900):                      # one chunk, and several (the tensor map follows the chunk's row count)
            sess = B200Session(p, precision="bf16")
            sess.engine.set_option(option, opt)
            if option == "conv_tma":
                sess.engine.set_option("conv_tma_max_cin", 4096)     # every launch on bf16 rows, the coupling flow's 192 channels too
            if chunk:
                sess.engine.set_option("max_chunk_frames", chunk)
            a, alen = sess.synthesize_packed(feed)
            outs.append((a.copy(), alen.copy()))
    sess.engine.set_option(option, 1)                 # conv_tma* are process-wide: leave the defaults behind
    sess.engine.set_option("conv_tma_max_cin", 128)
    for o in outs[1:]:
        assert np.array_equal(outs[0][1], o[1])
        assert np.isfinite(o[0]).all() and np.array_equal(outs[0][0], o[0])


def test_resblock1_stages_on_bf16_rows_match_fp32_rows(lib, tmp_path_factory):
    """ResBlock1 (`high`, modules.py:301-314) conv by conv: bf16 operand rows between the convs (default) vs fp32 rows with a
    conversion pass per launch (`no_stage_bf16`): same MMAs, the residual stream rounds to bf16 once per conv pair."""
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, "high", 1)
    rs = np.random.RandomState(12)
    lens = np.array([97, 3, 160, 41, 1], np.int64)
    B, T = len(lens), int(lens.max())
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    nd = rs.randn(B, 2, T).astype(np.float32)
    nz = rs.randn(B, arch.inter, 2600).astype(np.float32)
    feed = {"input": ids, "input_lengths": lens, "scales": SCALES, "noise_dp": nd, "noise_z": nz}
    outs = []
    for opt in (0, 1):
        sess = B200Session(p, precision="bf16")
        sess.engine.set_option("no_stage_bf16", opt)
        a, alen = sess.synthesize_packed(feed)
        outs.append((np.array(a), np.array(alen)))
    assert np.array_equal(outs[0][1], outs[1][1])
    # measured 44.9 dB between the two variants on these (short, ragged) utterances; against the fp32 oracle the bf16-row path reads
    # 61.6 dB on the `high` preset (profiles/r02_snr_report.txt) and is gated at 40 dB per utterance by test_gpu_baseline_configs C4
    assert snr_db(outs[1][0], outs[0][0]) > 40.0, snr_db(outs[1][0], outs[0][0])


def test_fused_resblock1_pairs_equal_conv_by_conv(lib, tmp_path_factory):
    """`high` preset, 128- / 64- / 32-channel stages: each (conv_{k,d} -> conv_{k,1}) pair of ResBlock1 (modules.py:301-314) as ONE fused
    launch (default) vs two conv launches with the intermediate as bf16 rows in HBM (`no_fused_rb1`).  With `no_rb1_rows_out` the fused
    path has the same rounding points as the conv-by-conv one (intermediate and pair output are bf16 lrelu rows, per-resblock results
    accumulate in fp32) and only the accumulation order inside a convolution differs; by default the per-resblock results of a stage
    that feeds another ConvTranspose also travel as bf16 rows and are combined by the last pair -- one more bf16 rounding per
    resblock (61.7 dB against the fp32 oracle either way, tools/snr_report.py)."""
    from phoonnx_b200.session import B200Session
    p, arch, _ = _voice(tmp_path_factory, "high", 1)
    rs = np.random.RandomState(13)
    lens = np.array([97, 3, 160, 41, 1, 2, 130], np.int64)
    B, T = len(lens), int(lens.max())
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    nd = rs.randn(B, 2, T).astype(np.float32)
    nz = rs.randn(B, arch.inter, 2600).astype(np.float32)
    feed = {"input": ids, "input_lengths": lens, "scales": SCALES, "noise_dp": nd, "noise_z": nz}
    outs = {}
    for name, opts in (("conv_by_conv", {"no_fused_rb1": 1}), ("fused_fp32_sum", {"no_rb1_rows_out": 1}), ("fused", {}),
                       ("fused_chunked", {"max_chunk_frames": 700}), ("fused_post", {"rb1_fused_post": 1})):
        sess = B200Session(p, precision="bf16")
        for k, v in opts.items():
            sess.engine.set_option(k, v)
        a, alen = sess.synthesize_packed(feed)
        outs[name] = (np.array(a), np.array(alen))
    for name in outs:
        assert np.array_equal(outs[name][1], outs["conv_by_conv"][1]) and np.isfinite(outs[name][0]).all(), name
    ref = outs["conv_by_conv"][0]
    assert snr_db(outs["fused_fp32_sum"][0], ref) > 55.0, snr_db(outs["fused_fp32_sum"][0], ref)
    assert snr_db(outs["fused"][0], ref) > 45.0, snr_db(outs["fused"][0], ref)
    # opt-in (`rb1_fused_post`): the last pair of the last stage also runs lrelu -> conv_post -> tanh on its bf16 operand tile (as the
    # medium voice's fused last stage does) instead of a separate fp32 conv_post pass: 40.3 dB against the conv-by-conv path on these
    # short, ragged utterances, 56.5 dB against the fp32 oracle on a full utterance (tools/snr_report.py) -- 7 dB of the preset's
    # margin for 6.5 % of its speed, which is why it is not the default
    assert snr_db(outs["fused_post"][0], ref) > 38.0, snr_db(outs["fused_post"][0], ref)
    assert np.array_equal(outs["fused"][0], outs["fused_chunked"][0])      # chunking is invisible to the fused path too
