"""GPU diagnostics: conv kernels vs numpy, then every stage of the path vs the oracle.
Usage: python tools/gpu_diag.py [conv_f32|conv_tc|stages:<preset>:<nspk>:<prec>] ...  (each section in a subprocess)"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def mini_engine(lib_path=None):
    from phoonnx_b200 import engine, modelgen, packing
    a = modelgen.make_arch("tiny")
    W = modelgen.synth_weights(a, 3)
    blobs, opts = packing.pack_model(W, a)
    if lib_path:
        engine._lib = engine.load_library(lib_path)
    return engine.Engine(a, blobs, opts)


def run_conv(eng, use_tc, x, w_tcn, bias, taps, in_slope=None, epi=0, res=None, accumulate=False, out_div=1.0,
             out_act=0, out_init=None):
    from phoonnx_b200 import packing
    blobs = {}
    packing.pack_conv(blobs, "c", w_tcn, bias, tc=True, tc3=(use_tc == 2))
    L, cin = x.shape
    n = w_tcn.shape[2]
    out_cols = n // 2 if epi == 1 else n
    out = np.zeros((L, out_cols), np.float32) if out_init is None else out_init.copy()
    taps_a = np.asarray(taps, np.int32)
    lib = eng.lib
    lib.vits_test_conv.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int,
                                   C.c_float, C.c_int, C.c_void_p, C.c_int]
    lib.vits_test_conv.restype = C.c_int
    p = lambda a_: None if a_ is None else a_.ctypes.data_as(C.c_void_p)
    wtc = blobs.get("c.wtc")
    if use_tc == 2:
        if "c.wtc3.0" not in blobs:
            raise ValueError("shape has no bf16x3 operands")
        parts, j = [], 0
        while f"c.wtc3.{j}" in blobs:
            parts.append(blobs[f"c.wtc3.{j}"].reshape(-1)); j += 1
        wtc = np.ascontiguousarray(np.concatenate(parts))
    xb = np.ascontiguousarray(x, np.float32)
    rc = lib.vits_test_conv(eng._h, int(use_tc), p(xb), L, cin, p(taps_a), len(taps), p(blobs["c.w"]), p(wtc),
                            p(blobs.get("c.b")), n, 0 if in_slope is None else 1, float(in_slope or 1.0), epi,
                            p(None if res is None else np.ascontiguousarray(res, np.float32)), int(accumulate),
                            float(out_div), out_act, p(out), out_cols)
    if rc != 0:
        raise RuntimeError(f"vits_test_conv rc={rc}: {lib.vits_last_error(eng._h).decode()}")
    return out


def ref_conv(x, w_tcn, bias, taps, in_slope=None, epi=0, res=None, accumulate=False, out_div=1.0, out_act=0,
             out_init=None, bf16=False):
    import emulate
    from phoonnx_b200.packing import bf16_round
    xx = x if in_slope is None else emulate.lrelu(x, in_slope)
    ww = w_tcn
    if bf16:
        xx, ww = bf16_round(xx), bf16_round(w_tcn)
    n = w_tcn.shape[2]
    v = emulate.conv_cl(xx, ww, bias, taps, n).astype(np.float64)
    if epi == 1:
        return (np.tanh(v[:, 0::2]) / (1.0 + np.exp(-v[:, 1::2]))).astype(np.float32)
    if epi == 3:
        return (res - v).astype(np.float32)
    if res is not None:
        v = v + res
    if out_act == 1:
        v = np.maximum(v, 0)
    if accumulate:
        v = v + out_init
    v = v / out_div
    if out_act == 2:
        v = np.tanh(v)
    return v.astype(np.float32)


CASES = [
    # (L, cin, n, taps, kwargs)
    (300, 64, 64, [-3, 0, 2], {}),
    (130, 32, 32, [-9, -6, -3, 0, 3, 6, 9], dict(in_slope=0.1, with_res=True)),
    (1000, 32, 32, [-36, -24, -12, 0, 12, 24, 36], dict(in_slope=0.1, with_res=True, accumulate=True, out_div=3.0)),
    (257, 128, 128, [-12, -6, 0, 6, 12], dict(in_slope=0.1, with_res=True)),
    (90, 256, 512, [-1, 0], dict(in_slope=0.1)),
    (77, 192, 256, [-3, -2, -1, 0, 1, 2, 3], {}),
    (200, 192, 384, [-2, -1, 0, 1, 2], dict(epi=1)),
    (200, 192, 384, [0], dict(epi=2, accumulate=True)),
    (200, 192, 192, [0], dict(accumulate=True)),
    (64, 64, 128, [0, 1], dict(in_slope=0.1)),
    (5, 16, 16, [-1, 0, 1], {}),
    (129, 48, 96, [-1, 0, 1], dict(out_act=1)),
]


def sec_conv(use_tc, lib_path=None):
    eng = mini_engine(lib_path)
    rs = np.random.RandomState(0)
    worst = 0.0
    for (L, cin, n, taps, kw) in CASES:
        kw = dict(kw)
        if use_tc and (cin % 16 or n % 16):
            continue
        x = rs.randn(L, cin).astype(np.float32)
        w = (rs.randn(len(taps), cin, n) / np.sqrt(cin * len(taps))).astype(np.float32)
        b = rs.randn(n).astype(np.float32)
        res = rs.randn(L, n).astype(np.float32) if kw.pop("with_res", False) else None
        out_init = rs.randn(L, n // 2 if kw.get("epi") == 1 else n).astype(np.float32) if kw.get("accumulate") else None
        t = time.time()
        try:
            got = run_conv(eng, use_tc, x, w, b, taps, res=res, out_init=out_init, **kw)
        except Exception as e:  # noqa
            print(f"  case L={L} cin={cin} n={n} taps={taps} {kw}: FAILED {e}")
            raise
        want = ref_conv(x, w, b, taps, res=res, out_init=out_init, bf16=bool(use_tc), **kw)
        err = float(np.abs(got - want).max())
        worst = max(worst, err)
        bad = np.argwhere(np.abs(got - want) > 1e-2)
        print(f"  case L={L} cin={cin} n={n} taps={taps} {kw}: max err {err:.3e} (|ref| max {np.abs(want).max():.2f}) "
              f"{'' if not len(bad) else 'first bad ' + str(bad[:4].tolist()) + ' n_bad=' + str(len(bad))} [{time.time()-t:.2f}s]")
    print(f"conv {'tc' if use_tc else 'f32'} worst err {worst:.3e}")


def sec_stages(preset, nspk, prec, use_sdp=True):
    import tempfile
    from phoonnx_b200 import modelgen
    from phoonnx_b200.session import B200Session
    from phoonnx_b200.weights import load_model
    from oracle.vits_oracle import VitsOracle
    td = tempfile.mkdtemp()
    path = os.path.join(td, "v.onnx")
    modelgen.make_voice(path, preset, nspk, seed=5, use_sdp=use_sdp)
    W, arch, _ = load_model(path)
    orc = VitsOracle(W, arch)
    sess = B200Session(path, precision=prec)
    sess.engine.set_option("debug_keep_zp", 1)
    rs = np.random.RandomState(1)
    lens = np.array([37, 5, 64, 1, 23], np.int64) if preset.startswith("tiny") else np.array([50, 17], np.int64)
    B, T = len(lens), int(lens.max())
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    nd = rs.randn(B, 2, T).astype(np.float32)
    sid = (np.arange(B) % nspk).astype(np.int64) if nspk > 1 else None
    scales = np.array([0.667, 1.0, 0.8], np.float32)
    nz = rs.randn(B, arch.inter, 4000).astype(np.float32)
    feed = {"input": ids, "input_lengths": lens, "scales": scales, "noise_dp": nd, "noise_z": nz}
    if sid is not None:
        feed["sid"] = sid
    o1 = [orc.infer(ids[b, :lens[b]], scales, None if sid is None else int(sid[b]), nd[b][:, :lens[b]], nz[b])
          for b in range(B)]
    t = time.time()
    audio, alen = sess.synthesize_packed(feed)
    dt = time.time() - t
    eng = sess.engine
    x = eng.fetch("x").reshape(-1, arch.hidden)
    stats = eng.fetch("stats").reshape(-1, 2 * arch.inter)
    logw = eng.fetch("logw")
    dur = eng.fetch("durations")
    print(f"[{preset} nspk={nspk} {prec} sdp={use_sdp}] synth {dt*1e3:.1f} ms, launches {eng.launch_count()}")
    off = 0
    aoff = 0
    zall = eng.fetch("z").reshape(-1, arch.inter)
    zpall = eng.fetch("z_p").reshape(-1, arch.inter)
    foff = 0
    for b in range(B):
        r = o1[b]
        L = int(lens[b])
        e = lambda u, v: float(np.abs(u - v).max())
        dd = dur[off:off + L]
        same = np.array_equal(dd, r["durations"])
        msg = (f"  utt{b} T={L}: x {e(x[off:off+L], r['x']):.2e} m_p {e(stats[off:off+L,:arch.inter], r['m_p']):.2e} "
               f"logs_p {e(stats[off:off+L,arch.inter:], r['logs_p']):.2e} logw {e(logw[off:off+L], r['logw']):.2e} dur_equal {same}")
        ny = int(alen[b]) // arch.hop
        if same:
            a_g = audio[aoff:aoff + int(alen[b])]
            zp_g = zpall[foff:foff + ny]; z_g = zall[foff:foff + ny]
            snr = 10 * np.log10((r["audio"] ** 2).sum() / max(((a_g - r["audio"]) ** 2).sum(), 1e-30))
            msg += (f" z_p {e(zp_g, r['z_p']):.2e} z {e(z_g, r['z']):.2e} audio {e(a_g, r['audio']):.2e} "
                    f"(peak {np.abs(r['audio']).max():.2f}) SNR {snr:.1f} dB")
        print(msg)
        off += L; aoff += int(alen[b]); foff += ny


def main():
    args = sys.argv[1:]
    if not args:
        args = ["conv_f32", "conv_tc", "stages:tiny:1:fp32", "stages:tiny:3:fp32", "stages:tiny_rb1:1:fp32",
                "stages:tiny:1:fp32:nosdp", "stages:x_low:1:fp32", "stages:medium:1:fp32", "stages:medium:8:fp32",
                "stages:x_low:1:bf16", "stages:medium:1:bf16", "stages:high:1:bf16"]
    if len(args) == 1 and args[0].startswith("@"):
        sec = args[0][1:]
        if sec == "conv_f32":
            sec_conv(0)
        elif sec.startswith("conv_tc"):
            parts = sec.split("=")
            sec_conv(1, parts[1] if len(parts) > 1 else None)
        elif sec.startswith("stages:"):
            p = sec.split(":")
            sec_stages(p[1], int(p[2]), p[3], use_sdp=not (len(p) > 4 and p[4] == "nosdp"))
        return
    for sec in args:
        print(f"===== {sec}", flush=True)
        t = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "@" + sec], timeout=300)
            print(f"===== {sec}: exit {r.returncode} in {time.time()-t:.1f}s", flush=True)
        except subprocess.TimeoutExpired:
            print(f"===== {sec}: TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
