"""GPU diagnostics: (1) tcgen05.mma issue-rate probe vs N; (2) phase timeline of the fused MRF kernel (CTA 0)."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phoonnx_b200 import modelgen  # noqa: E402
from phoonnx_b200.session import B200Session  # noqa: E402


def probe(sess):
    lib, h = sess.engine.lib, sess.engine._h
    lib.vits_test_mma_probe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_double), C.POINTER(C.c_double)]
    print("N  nd na rows ctas | issue cyc/MMA  total cyc/MMA  (floor N/2)")
    for nctas in (1, 148):
        for n, nd, na, rows in ((32, 1, 1, 545), (32, 4, 7, 545), (32, 8, 7, 545), (64, 1, 1, 545), (64, 2, 7, 300), (64, 4, 7, 545),
                                (128, 1, 1, 300), (128, 2, 5, 300), (192, 2, 5, 300), (256, 2, 5, 300), (16, 4, 7, 545)):
            a, b = C.c_double(), C.c_double()
            rc = lib.vits_test_mma_probe(h, n, 4096, nd, na, rows, nctas, C.byref(a), C.byref(b))
            print(f"{n:3d} {nd:2d} {na:2d} {rows:4d} {nctas:4d} | {a.value:8.1f} {b.value:8.1f}   ({n / 2:.0f})  rc={rc}")
    # operand placement: does the small-N cost follow the smem layout (swizzle) or the A operand's trip through shared memory at all?
    lib.vits_test_mma_probe_mode.argtypes = [C.c_void_p, C.c_int] + lib.vits_test_mma_probe.argtypes[1:]
    names = {0: "A,B smem no-swizzle", 1: "A,B smem SWIZZLE_128B", 2: "A tmem, B smem no-swizzle", 3: "A tmem, B smem SWIZZLE_128B",
             4: "cp A smem->tmem + MMA A tmem", 5: "cp A smem->tmem alone",
             6: "cta_group::2, A,B smem", 7: "cta_group::2, A tmem"}
    print("mode N  nd na | total cyc/MMA at 1 CTA, at 148 CTAs  (floor N/2)")
    for mode in (0, 1, 2, 3, 4, 5, 6, 7):
        for n, nd, na in ((16, 4, 7), (32, 4, 7), (64, 4, 7), (96, 4, 7), (128, 2, 5), (192, 2, 5), (256, 1, 5)):
            if mode >= 6 and n % 32:
                continue
            if mode >= 2 and mode != 6 and nd * n > 384:
                nd = 384 // n
            tot = []
            for nctas in ((2, 148) if mode >= 6 else (1, 148)):
                a, b = C.c_double(), C.c_double()
                rc = lib.vits_test_mma_probe_mode(h, mode, n, 4096, nd, na, 545, nctas, C.byref(a), C.byref(b))
                tot.append(b.value if rc == 0 else float("nan"))
            print(f"{mode} [{names[mode]:28s}] {n:3d} {nd:2d} {na:2d} | {tot[0]:8.1f} {tot[1]:8.1f}   ({n / 2:.0f})")


def timeline(sess, stage, arch):
    eng = sess.engine
    eng.set_option("mrf_dbg", stage)
    rs = np.random.RandomState(0)
    B = 24
    lens = rs.randint(150, 257, size=(B,)).astype(np.int64)
    x = np.zeros((B, int(lens.max())), np.int64)
    for b in range(B):
        x[b, :lens[b]] = rs.randint(0, arch.n_vocab, size=(int(lens[b]),))
    feed = {"input": x, "input_lengths": lens, "scales": np.asarray((0.667, 1.0, 0.8), np.float32)}
    for _ in range(2):
        sess.synthesize_packed(feed, out="none")
    buf = np.zeros((24 * 48 * 2,), np.float32)
    n = eng.lib.vits_fetch(eng._h, b"mrf_dbg", buf.ctypes.data_as(C.c_void_p), buf.size)
    st = buf.view(np.uint64).reshape(24, 48).astype(np.int64)
    t0 = st[0, 0]
    names = {0: "E top", 1: "E ups acc ready", 2: "E X staged (E0)", 3: "E post(prev) done", 16: "E c2 final done", 17: "E final epi done",
             20: "M top", 21: "M ups issued", 22: "M X ready", 40: "M post rdy", 41: "M post issued", 44: "L top", 45: "L in free", 46: "L landed"}
    names.update({36: "M before wait x1(0)", 37: "M before wait x1(1)", 38: "M before wait x1(2)", 18: "M after wait x1(0)",
                  23: "E(warp 15) x1(0) staged", 42: "E(warp 15) x1(1) staged", 43: "E(warp 15) x1(2) staged"})
    for r in range(3):
        names[4 + 4 * r] = f"E c1({r}) done"; names[5 + 4 * r] = f"E E1({r}) math done"
        names[6 + 4 * r] = f"E c2({r - 1}) done"; names[7 + 4 * r] = f"E x1({r}) staged"
    for cv in range(2):
        for r in range(3):
            names[24 + 2 * (cv * 3 + r)] = f"M start C{cv + 1}({r})"; names[25 + 2 * (cv * 3 + r)] = f"M issued C{cv + 1}({r})"
    print(f"--- stage {stage} timeline (cycles rel. to tile 0 start), fetched {n}")
    for it in (1, 2, 5, 6):
        ev = sorted((int(st[it, s]) - int(t0), names.get(s, str(s))) for s in range(48) if st[it, s])
        print(f"tile {it}: span {ev[-1][0] - ev[0][0]} cycles")
        base = ev[0][0]
        for t, nm in ev:
            print(f"   {t - base:8d}  {nm}")
    # one merged, absolute listing of three consecutive tiles: the hand-over between tiles is where the single-tile views are blind
    ev = sorted((int(st[it, s]) - int(t0), f"[{it}] " + names.get(s, str(s))) for it in (3, 4, 5) for s in range(48) if st[it, s])
    print("merged tiles 3-5 (absolute cycles since tile 0's first stamp):")
    for t, nm in ev:
        if nm[4] in "ME":
            print(f"   {t:8d}  {nm}")
    starts = st[:, 0]
    print("tile period (E top to E top):", np.diff(starts[starts > 0])[:20])
    eng.set_option("mrf_dbg", 0)


def main():
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "m.onnx")
    _, arch = modelgen.make_voice(path, "medium", n_speakers=1, seed=1234)
    sess = B200Session(path, precision="bf16")
    if '--probe' in sys.argv:
        probe(sess)
    if '--probe-only' in sys.argv:
        return
    timeline(sess, 3, arch)
    timeline(sess, 2, arch)


if __name__ == "__main__":
    main()
