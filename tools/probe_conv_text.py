"""GPU diagnostics: phase timeline of k_conv_tc (CTA (0,0)) for the text-side (bf16x3) launches of one prepare call."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phoonnx_b200 import modelgen  # noqa: E402
from phoonnx_b200.session import B200Session  # noqa: E402
from probe_conv import NAMES  # noqa: E402


def main():
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "m.onnx")
    _, arch = modelgen.make_voice(path, "medium", n_speakers=1, seed=1234)
    sess = B200Session(path, precision="bf16")
    eng = sess.engine
    rs = np.random.RandomState(0)
    B = 800
    lens = rs.randint(64, 257, size=(B,)).astype(np.int64)
    x = np.zeros((B, int(lens.max())), np.int64)
    for b in range(B):
        x[b, :lens[b]] = rs.randint(0, arch.n_vocab, size=(int(lens[b]),))
    feed = {"input": x, "input_lengths": lens, "scales": np.asarray((0.667, 1.0, 0.8), np.float32)}
    sess.synthesize_packed(feed, out="none")
    print("ids", int(lens.sum()))
    for idx, what in ((1, "enc qkv 1x1 192->576 (bf16x3, 2 K slices, 3 N tiles)"), (2, "enc o 1x1 192->192 + fp32 residual"),
                      (3, "enc ffn1 k3 192->768 relu (3 N tiles)"), (4, "enc ffn2 k3 768->192 + residual (8 K slices)")):
        eng.set_option("conv_text_dbg", idx)
        sess.synthesize_packed(feed, out="none")
        buf = np.zeros((16 * 16 * 2,), np.float32)
        n = eng.lib.vits_fetch(eng._h, b"conv_dbg", buf.ctypes.data_as(C.c_void_p), buf.size)
        st = buf.view(np.uint64).reshape(16, 16).astype(np.int64)
        t0 = int(st[0, 15])
        print(f"=== text conv launch {idx}: {what} (fetched {n})")
        for it in range(5):
            ev = sorted((int(st[it, s]) - t0, NAMES.get(s, str(s))) for s in range(13) if st[it, s])
            if not ev:
                break
            print(f" tile {it}: " + "  ".join(f"{nm}@{t}" for t, nm in ev) + f"  | epi phase1 {int(st[it, 13]) & 0xFFFFF} (prefetch +{(int(st[it, 13]) >> 20) & 0xFFFFF}, tmem +{(int(st[it, 13]) >> 40) & 0xFFFFF}) phase2 {int(st[it, 14])}")
    eng.set_option("conv_text_dbg", 0)


if __name__ == "__main__":
    main()
