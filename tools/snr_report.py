"""bf16-mode SNR vs the fp32 oracle for the presets (the numbers DESIGN.md section 4 quotes)."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phoonnx_b200 import modelgen
from phoonnx_b200.session import B200Session
from phoonnx_b200.weights import load_model
from oracle.vits_oracle import VitsOracle

SC = np.asarray((0.667, 1.0, 0.8), np.float32)
for preset, ns, lens in (("x_low", 1, [64, 120]), ("medium", 1, [128, 40]), ("medium", 8, [77, 64]), ("high", 1, [96])):
    td = tempfile.mkdtemp(); p = os.path.join(td, "v.onnx")
    modelgen.make_voice(p, preset, n_speakers=ns, seed=11)
    W, arch, _ = load_model(p)
    orc = VitsOracle(W, arch)
    rs = np.random.RandomState(4)
    B, T = len(lens), max(lens)
    ids = rs.randint(0, arch.n_vocab, (B, T)).astype(np.int64)
    nd = rs.randn(B, 2, T).astype(np.float32); nz = rs.randn(B, arch.inter, 16 * T + 64).astype(np.float32)
    feed = {"input": ids, "input_lengths": np.asarray(lens, np.int64), "scales": SC, "noise_dp": nd, "noise_z": nz}
    if ns > 1: feed["sid"] = (np.arange(B) % ns).astype(np.int64)
    for prec in ("fp32", "bf16"):
        sess = B200Session(p, precision=prec)
        audio, alen = sess.synthesize_packed(feed)
        off = 0; out = []
        for b in range(B):
            L = lens[b]
            r = orc.infer(ids[b, :L], SC, None if ns == 1 else int(feed["sid"][b]), nd[b][:, :L], nz[b], stages=False)
            a = audio[off:off + int(alen[b])]; off += int(alen[b])
            if a.shape == r["audio"].shape:
                err = float(np.abs(a - r["audio"]).max())
                snr = 10 * np.log10(float((r["audio"] ** 2).sum()) / max(float(((a - r["audio"]) ** 2).sum()), 1e-30))
                out.append(f"max-abs {err:.2e} SNR {snr:.1f} dB (peak {np.abs(r['audio']).max():.3f})")
            else:
                out.append(f"length {a.shape} vs {r['audio'].shape}")
        print(preset, ns, prec, "|", "; ".join(out), flush=True)
