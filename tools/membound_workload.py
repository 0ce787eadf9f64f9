"""Workload behind profiles/*_membound.txt: one 2048-utterance medium device batch (the benched shape), int16 output so that the
post-processing kernels run too.  Prints the sizes the algorithmic-byte formulas of tools/ncu_membound.py need."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phoonnx_b200 import modelgen, scheduler  # noqa: E402
from phoonnx_b200.session import B200Session  # noqa: E402


def main():
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "m.onnx")
    _, arch = modelgen.make_voice(path, "medium", n_speakers=1, seed=1234)
    rs = np.random.RandomState(2)
    lengths = rs.randint(64, 257, size=(2048,)).astype(np.int64)
    utts = [rs.randint(0, arch.n_vocab, size=(int(L),)).astype(np.int64) for L in lengths]
    order = scheduler.plan(lengths, 1, 0, max_ids=1 << 20, max_utts=2048)[0]
    x, lens = scheduler.pad_batch([utts[i] for i in order])
    sess = B200Session(path, precision="bf16", max_chunk_frames=262144, seed=3)
    feed = {"input": x, "input_lengths": lens, "scales": np.asarray((0.667, 1.0, 0.8), np.float32)}
    frames = 0
    for _ in range(int(os.environ.get("PASSES", "2"))):
        audio, alen = sess.synthesize_packed(feed, out="i16")
        frames = int(alen.sum()) // arch.hop
    print(json.dumps({"ids": int(lens.sum()), "frames": frames, "utterances": int(lens.size), "hidden": arch.hidden, "inter": arch.inter,
                      "dp_filter": arch.dp_filter, "hop": arch.hop, "chunk_frames": 262144}))


if __name__ == "__main__":
    main()
