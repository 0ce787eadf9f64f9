"""Mint the golden fixtures under tests/golden/ from the REAL reference (run in the build
container, where /root/reference exists; the fixtures travel, the reference does not).

For each miniature voice (same topology as the reference presets, small widths so the genuine
exporter-format file is committable):
  <name>.onnx.gz   -- written by the reference's exporter logic (export_onnx.py:250-327)
  <name>.npz       -- inputs (ids, injected noise, scales, sid) and the reference's own
                      SynthesizerTrn.infer outputs + stage tensors for several utterances
Also records the phoneme-id layout example of SURVEY.md 8(c) produced by the reference's
phoneme_ids.phonemes_to_ids.

Usage: python -m oracle.make_golden
"""
import gzip
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_bridge as rb  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

VOICES = [("tiny", 1), ("tiny", 3), ("tiny_rb1", 1)]
LENGTHS = [40, 7, 1, 3, 5, 6]          # incl. the T <= window+1 branches of attentions.py:295-305


def main():
    os.makedirs(GOLD, exist_ok=True)
    for preset, ns in VOICES:
        name = f"{preset}_spk{ns}"
        m = rb.build_reference_model(preset, n_speakers=ns)
        with tempfile.TemporaryDirectory() as td:
            p = os.path.join(td, "m.onnx")
            rb.export_onnx(m, p, n_speakers=ns)
            with open(p, "rb") as f, gzip.GzipFile(os.path.join(GOLD, name + ".onnx.gz"), "wb", mtime=0) as g:
                g.write(f.read())
        rs = np.random.RandomState(1234)
        out = {}
        for u, T in enumerate(LENGTHS):
            ids = rs.randint(0, 256, (T,)).astype(np.int64)
            nd = rs.randn(2, T).astype(np.float32)
            nz_full = rs.randn(32, 64 * T + 8).astype(np.float32)
            sid = (u % ns) if ns > 1 else None
            for tag, scales, a_nd, a_nz in (("n", (0.667, 1.0, 0.8), nd, nz_full),
                                            ("z", (0.0, 1.3, 0.0), None, None)):
                r = rb.reference_infer(m, ids, scales, sid, a_nd, a_nz)
                ty = r["z"].shape[0]
                k = f"u{u}{tag}_"
                out[k + "ids"] = ids
                out[k + "scales"] = np.asarray(scales, np.float32)
                out[k + "sid"] = np.asarray(-1 if sid is None else sid, np.int64)
                if a_nd is not None:
                    out[k + "noise_dp"] = a_nd
                    out[k + "noise_z"] = np.ascontiguousarray(a_nz[:, :ty])
                for s in ("x", "m_p", "logs_p", "logw", "durations", "z_p", "z", "audio"):
                    out[k + s] = r[s]
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(name, "utterances", len(LENGTHS), "files written")
    # phoneme-id layout example from the reference's own phonemes_to_ids (phoneme_ids.py:209-310)
    try:
        sys.path.insert(0, rb.REF_ROOT)
        from phoonnx.phoneme_ids import phonemes_to_ids
        ids = phonemes_to_ids(list("həlˈoʊ wˈɜːld"))
        np.save(os.path.join(GOLD, "phoneme_ids_example.npy"), np.asarray(ids, np.int64))
        print("phoneme id example:", ids)
    except Exception as e:  # pragma: no cover
        print("phoneme id example skipped:", e)


if __name__ == "__main__":
    main()
