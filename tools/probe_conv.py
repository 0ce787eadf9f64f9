"""GPU diagnostics: phase timeline of k_conv_tc (CTA (0,0)) for selected launches of one decode call."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phoonnx_b200 import modelgen  # noqa: E402
from phoonnx_b200.session import B200Session  # noqa: E402

NAMES = {0: "E top", 1: "E geom", 2: "E acc full", 3: "E done", 4: "L top", 5: "L buf free", 6: "L staged", 7: "W top", 8: "W all issued",
         9: "M top", 10: "M A full", 11: "M acc free", 12: "M issued", 15: "kernel start"}


def main():
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "m.onnx")
    _, arch = modelgen.make_voice(path, "medium", n_speakers=1, seed=1234)
    sess = B200Session(path, precision="bf16", max_chunk_frames=int(sys.argv[1]) if len(sys.argv) > 1 else 131072)
    eng = sess.engine
    if len(sys.argv) > 2:
        eng.set_option("num_sms", int(sys.argv[2]))
    rs = np.random.RandomState(0)
    B = 96
    lens = rs.randint(150, 257, size=(B,)).astype(np.int64)
    x = np.zeros((B, int(lens.max())), np.int64)
    for b in range(B):
        x[b, :lens[b]] = rs.randint(0, arch.n_vocab, size=(int(lens[b]),))
    feed = {"input": x, "input_lengths": lens, "scales": np.asarray((0.667, 1.0, 0.8), np.float32)}
    sess.synthesize_packed(feed, out="none")
    print("frames", int(sess.last_lengths.sum()) // 256)
    for idx, what in ((2, "flow in-layer k5 192->384 gate -> bf16 acts"), (3, "flow rsr 1x1 192->192 bf16 in, fh accumulate"),
                      (9, "flow mskip K=4x192 -> 96 subfrom"), (37, "dec conv_pre k7 192->256"), (38, "ups0 A -> bf16"),
                      (40, "stage1 rb0 conv1 k3 d1 bf16 in/res -> bf16 x1 rows"), (42, "stage1 rb2 conv1 k7 d3 -> bf16 x1 rows"),
                      (43, "stage1 summed second convs: 3 K slices, 15 taps, 3 bf16 residuals -> bf16 rows"),
                      (44, "ups1 A 128->4*64, bf16 rows in -> bf16")):
        eng.set_option("conv_dbg", idx)
        sess.synthesize_packed(feed, out="none")
        buf = np.zeros((16 * 16 * 2,), np.float32)
        n = eng.lib.vits_fetch(eng._h, b"conv_dbg", buf.ctypes.data_as(C.c_void_p), buf.size)
        st = buf.view(np.uint64).reshape(16, 16).astype(np.int64)
        t0 = int(st[0, 15])
        print(f"=== conv launch {idx}: {what} (fetched {n})")
        for it in range(6):
            ev = sorted((int(st[it, s]) - t0, NAMES.get(s, str(s))) for s in range(13) if st[it, s])
            if not ev:
                break
            print(f" tile {it}: " + "  ".join(f"{nm}@{t}" for t, nm in ev) + f"  | epi phase1 {int(st[it, 13]) & 0xFFFFF} (prefetch issued +{(int(st[it, 13]) >> 20) & 0xFFFFF}, tmem loaded +{(int(st[it, 13]) >> 40) & 0xFFFFF}, cumulative over the warp's groups) phase2 {int(st[it, 14])}")
    eng.set_option("conv_dbg", 0)


if __name__ == "__main__":
    main()
