"""Build libvits_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to the GPU box)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "engine.cu")
SRC_MAS = os.path.join(HERE, "csrc", "mas.cu")          # monotonic alignment search (include/mas_b200.h), same library
OUT = os.path.join(HERE, "libvits_b200.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("engine.cu", "common.cuh", "kernels_f32.cuh", "attention.cuh", "conv_tc.cuh", "mrf_tiles.cuh", "mrf3_tc.cuh", "probe_tc.cuh", "voice_file.h", "mas.cu")] + \
       [os.path.join(os.path.dirname(HERE), "include", f) for f in ("vits_b200.h", "vits_b200_test.h", "mas_b200.h")]


def nvcc_path() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.exists(p) or p == "nvcc"):
            return p
    return "nvcc"


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = OUT) -> str:
    if not force and os.path.exists(out):
        newest = max(os.path.getmtime(d) for d in DEPS)
        if os.path.getmtime(out) >= newest:
            return out
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", out, SRC, SRC_MAS, "-lz", *extra_flags]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libvits_b200.so")
    if verbose:
        print(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
