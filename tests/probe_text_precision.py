"""Why the text side runs as bf16x3: duration flips of the text-side contractions under four operand
precisions, emulated on the CPU oracle (TEST INFRASTRUCTURE: run from tests/, never from the product path).

The text encoder + duration predictor end in ceil(exp(logw) * length_scale) (models.py:702-704): an error of 1e-3 in logw flips
some durations by one frame, and a flip changes the LENGTH of the audio.  The probe rounds the operands of every text-side
convolution / 1x1 GEMM (conv_q/k/v/o, FFN, proj, duration predictor) the way a tensor-core path would and counts durations
that differ from the fp32 result:

    fp32      reference
    bf16      x, w rounded to bf16, fp32 accumulate                      (what the decoder / flow use)
    tf32      x, w rounded to 10 explicit mantissa bits
    bf16x3    x*w ~= xh*wh + xh*wl + xl*wh  (xh = bf16(x), xl = bf16(x - xh); same for w)      (what the engine runs)

usage: python tests/probe_text_precision.py [n_utts]     -> markdown on stdout (profiles/r02_text_precision_probe.md)"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.vits_oracle import VitsOracle  # noqa: E402
from phoonnx_b200 import modelgen  # noqa: E402
from phoonnx_b200.weights import load_model  # noqa: E402


def _bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _tf32(t):
    u = t.contiguous().view(torch.int32)
    r = (u + 0x0FFF + ((u >> 13) & 1)) & ~0x1FFF            # round to nearest even at 13 dropped bits
    return r.view(torch.float32)


class ProbeOracle(VitsOracle):
    mode = "fp32"

    def _conv(self, x, name, dilation=1, padding=0, groups=1):
        text_side = name.startswith("enc_p.") or name.startswith("dp.")
        if self.mode == "fp32" or not text_side or groups != 1:
            return super()._conv(x, name, dilation=dilation, padding=padding, groups=groups)
        w, b = self.W[name + ".weight"], self.W.get(name + ".bias")
        conv = lambda xx, ww: F.conv1d(xx.double(), ww.double(), None, dilation=dilation, padding=padding).float()   # exact products, fp32 result
        if self.mode == "bf16":
            y = conv(_bf16(x), _bf16(w))
        elif self.mode == "tf32":
            y = conv(_tf32(x), _tf32(w))
        else:
            xh, wh = _bf16(x), _bf16(w)
            xl, wl = _bf16(x - xh), _bf16(w - wh)
            y = conv(xh, wh) + conv(xh, wl) + conv(xl, wh)
        return y if b is None else y + b[None, :, None]


def run(n_utts=24, seed=0, preset="medium"):
    tmp = tempfile.mkdtemp()
    p = os.path.join(tmp, "v.onnx")
    modelgen.make_voice(p, preset, 1, seed=1234)
    W, arch, _ = load_model(p)
    orc = ProbeOracle(W, arch)
    rs = np.random.RandomState(seed)
    lens = rs.randint(64, 257, size=(n_utts,))
    utts = [rs.randint(0, arch.n_vocab, size=(int(L),)).astype(np.int64) for L in lens]
    noise = [rs.randn(2, int(L)).astype(np.float32) for L in lens]
    out = {}
    for mode in ("fp32", "bf16", "tf32", "bf16x3"):
        orc.mode = mode
        lw, du = [], []
        for u, nd in zip(utts, noise):
            x, _, _ = orc.text_encoder(u)
            logw = orc.sdp_reverse(x, nd, 0.8, None)
            d, _ = orc.durations_from_logw(logw, 1.0)
            lw.append(logw.numpy().copy()); du.append(d.numpy().copy())
        out[mode] = (np.concatenate(lw), np.concatenate(du))
    ref_lw, ref_d = out["fp32"]
    rows = []
    for mode in ("bf16", "tf32", "bf16x3"):
        lw, d = out[mode]
        rows.append((mode, float(np.abs(lw - ref_lw).max()), int((d != ref_d).sum())))
    return int(ref_d.size), rows


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    ids, rows = run(n)
    print(f"# Text-side operand precision vs duration flips ({ids} phoneme ids, {n} utterances, medium voice, noise_w 0.8, length_scale 1)\n")
    print("Emulated on the CPU oracle by `tests/probe_text_precision.py` (operands of every text-side convolution / 1x1 GEMM rounded,")
    print("products exact, everything else fp32).  A flipped duration changes the audio length by one frame (256 samples).\n")
    print("| operand precision | max abs(logw - logw_fp32) | durations that differ from fp32 |")
    print("|---|---|---|")
    for mode, e, f in rows:
        print(f"| {mode} | {e:.2e} | {f} of {ids} |")
