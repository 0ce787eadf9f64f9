"""One-utterance calls of B200Session.run (bench C1's workload) for a launch list: python tools/c1_one_call.py [ncalls]"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phoonnx_b200 import modelgen
from phoonnx_b200.session import B200Session

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    td = tempfile.mkdtemp()
    path = os.path.join(td, "m.onnx")
    _, arch = modelgen.make_voice(path, "medium", n_speakers=1, seed=1234)
    sess = B200Session(path, precision="bf16")
    rs = np.random.RandomState(0)
    ids = rs.randint(0, arch.n_vocab, (1, 128)).astype(np.int64)
    feed = {"input": ids, "input_lengths": np.asarray([128], np.int64), "scales": np.asarray((0.667, 1.0, 0.8), np.float32)}
    for _ in range(n):
        out = sess.run(None, feed)[0]
    print(out.shape)

if __name__ == "__main__":
    main()
