"""The C++ voice loader inside libvits_b200.so (csrc/voice_file.h: what vits_open runs) against the Python loader that stays its
reference implementation (onnx_reader.py + weights.py + packing.py): same architecture, same blob set, same bytes -- on
exporter-format files of every preset and on the genuine torch.onnx.export fixtures.  No GPU needed: the hooks stop before CUDA."""
import ctypes as C
import os

import numpy as np
import pytest

from phoonnx_b200 import modelgen
from phoonnx_b200.engine import CArch, to_c_arch
from phoonnx_b200.packing import pack_model
from phoonnx_b200.weights import load_model


@pytest.fixture(scope="module")
def lib(built_lib):
    lib = C.CDLL(built_lib)
    lib.vits_test_file_arch.argtypes = [C.c_char_p, C.POINTER(CArch), C.c_char_p, C.c_size_t]
    lib.vits_test_file_arch.restype = C.c_int
    lib.vits_test_file_blob.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int)]
    lib.vits_test_file_blob.restype = C.c_int64
    return lib


def _native_blobs(lib, path):
    p = path.encode()
    n = lib.vits_test_file_blob(p, None, None, 0, None)
    assert n > 0, n
    out = {}
    for i in range(n):
        nb = C.create_string_buffer(256)
        ln = lib.vits_test_file_blob(p, f"#{i}".encode(), nb, 256, None)
        assert 0 < ln < 256
        name = nb.value.decode()
        dt = C.c_int(-1)
        size = lib.vits_test_file_blob(p, name.encode(), None, 0, C.byref(dt))
        assert size > 0, name
        buf = np.empty((size,), np.uint8)
        assert lib.vits_test_file_blob(p, name.encode(), buf.ctypes.data_as(C.c_void_p), size, C.byref(dt)) == size
        out[name] = (buf, dt.value)
    return out


def _check(lib, path):
    W, arch, _ = load_model(path)
    want_arch = to_c_arch(arch)
    got_arch = CArch()
    err = C.create_string_buffer(512)
    assert lib.vits_test_file_arch(path.encode(), C.byref(got_arch), err, 512) == 0, err.value
    assert bytes(got_arch) == bytes(want_arch), "architecture inferred by the C++ loader differs from the Python loader's"
    blobs, opts = pack_model(W, arch)
    native = _native_blobs(lib, path)
    assert set(native) == set(blobs) | {"opt:" + k for k in opts}
    derived = 0
    for name, arr in blobs.items():
        buf, dt = native[name]
        assert dt == {np.dtype(np.float32): 0, np.dtype(np.uint16): 1}[arr.dtype], name
        want = np.ascontiguousarray(arr)
        assert buf.size == want.nbytes, (name, buf.size, want.nbytes)
        got = buf.view(want.dtype).reshape(want.shape)
        if np.array_equal(got, want):
            continue
        # tables / composed GEMM weights accumulated in double: numpy's BLAS sums in another order -> at most the last fp32 / bf16 bit
        assert ".cond_tab" in name or ".mskip" in name, f"{name}: bytes differ"
        derived += 1
        if want.dtype == np.float32:
            assert np.allclose(got, want, rtol=3e-7, atol=1e-9), name
        else:
            g32 = (got.astype(np.uint32) << 16).view(np.float32); w32 = (want.astype(np.uint32) << 16).view(np.float32)
            assert np.abs(got.astype(np.int64) - want.astype(np.int64)).max() <= 1 and np.allclose(g32, w32, rtol=8e-3, atol=1e-9), name
    for k, v in opts.items():
        buf, dt = native["opt:" + k]
        assert dt == 3 and abs(float(buf.view(np.float64)[0]) - float(v)) < 1e-12, k
    return len(blobs), derived


@pytest.mark.parametrize("preset,ns,sdp", [("tiny", 1, True), ("tiny", 3, True), ("tiny_rb1", 1, False), ("x_low", 1, True),
                                           ("medium", 8, True), ("high", 1, True)])
def test_native_loader_equals_python_loader(lib, tmp_path, preset, ns, sdp):
    p = str(tmp_path / "v.onnx")
    modelgen.make_voice(p, preset, ns, use_sdp=sdp, seed=3)
    n, derived = _check(lib, p)
    assert n > 50


@pytest.mark.parametrize("name", ["tiny_spk1", "tiny_spk3", "tiny_rb1_spk1"])
def test_native_loader_on_genuine_exporter_files(lib, golden_dir, name):
    """torch.onnx.export output of the reference's own model (gzip-compressed fixtures: read through zlib)."""
    _check(lib, os.path.join(golden_dir, name + ".onnx.gz"))


def test_native_loader_rejects_what_the_python_loader_rejects(lib, tmp_path):
    from phoonnx_b200 import onnx_reader
    err = C.create_string_buffer(512)
    a = CArch()
    p = str(tmp_path / "x.onnx")
    with open(p, "wb") as f:
        f.write(onnx_reader.encode_model({"foo": np.zeros((2, 2), np.float32)}, [], ["input"], ["output"], {}))
    assert lib.vits_test_file_arch(p.encode(), C.byref(a), err, 512) == -1 and b"missing tensor" in err.value
    with open(p, "wb") as f:
        f.write(b"\x00\x01garbage")
    assert lib.vits_test_file_arch(p.encode(), C.byref(a), err, 512) == -1
    assert lib.vits_test_file_arch(str(tmp_path / "nope.onnx").encode(), C.byref(a), err, 512) == -1
    arch = modelgen.make_arch("tiny")
    W = modelgen.synth_weights(arch, 1)
    W["emb_l.weight"] = np.zeros((2, 4), np.float32)
    modelgen.write_onnx(W, arch, p, graph_inputs=["input", "input_lengths", "scales", "langid"])
    assert lib.vits_test_file_arch(p.encode(), C.byref(a), err, 512) == -1 and b"multi-lingual" in err.value
