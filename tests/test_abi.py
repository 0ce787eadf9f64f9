"""The C-ABI shared library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute calls here (no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT, has_gpu


def declared_symbols(header="vits_b200.h"):
    hdr = open(os.path.join(ROOT, "include", header)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vits_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported(built_lib):
    names = declared_symbols()
    assert {"vits_create", "vits_upload", "vits_finalize", "vits_prepare", "vits_decode", "vits_destroy",
            "vits_last_error", "vits_fetch"} <= set(names)
    assert {"vits_describe", "vits_max_output_samples", "vits_set_stream", "vits_output_ticket", "vits_wait_ticket"} <= set(names)
    lib = ctypes.CDLL(built_lib)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vits_b200.h but not exported"
    lib.vits_abi_version.restype = ctypes.c_int
    assert lib.vits_abi_version() == 2


def test_test_hooks_live_in_their_own_header(built_lib):
    """The product header declares no vits_test_* entry point; the test header's hooks are exported too."""
    assert not [n for n in declared_symbols() if n.startswith("vits_test_")]
    hooks = declared_symbols("vits_b200_test.h")
    assert hooks and all(n.startswith("vits_test_") for n in hooks)
    lib = ctypes.CDLL(built_lib)
    for n in hooks:
        assert hasattr(lib, n), n


def test_mas_header_symbols_are_exported(built_lib):
    """include/mas_b200.h (monotonic alignment search) lives in the same library."""
    hdr = open(os.path.join(ROOT, "include", "mas_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(mas_[a-z0-9_]+)\s*\(", hdr)))
    from phoonnx_b200 import monotonic_align
    assert set(names) == set(monotonic_align.MAS_SYMBOLS)
    lib = ctypes.CDLL(built_lib)
    for n in names:
        assert hasattr(lib, n), n


def test_mas_flag_values_match_the_header():
    from phoonnx_b200 import monotonic_align as ma
    hdr = open(os.path.join(ROOT, "include", "mas_b200.h")).read()
    flags = dict(re.findall(r"#define\s+(MAS_[A-Z_0-9]+)\s+(0x[0-9a-fA-F]+)", hdr))
    assert int(flags["MAS_DEVICE_PTRS"], 16) == ma.MAS_DEVICE_PTRS and int(flags["MAS_PATH_F32"], 16) == ma.MAS_PATH_F32
    assert int(flags["MAS_TIMED"], 16) == ma.MAS_TIMED and int(flags["MAS_ROW_KERNEL"], 16) == ma.MAS_ROW_KERNEL


def test_mas_refuses_cpu_tensors(built_lib):
    import torch
    from phoonnx_b200 import monotonic_align
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        monotonic_align.maximum_path(torch.zeros(1, 4, 2), torch.ones(1, 4, 2))


def test_python_binding_covers_header(built_lib):
    from phoonnx_b200 import engine
    assert set(engine.EXPORTED_SYMBOLS) == set(declared_symbols())
    assert set(engine.TEST_SYMBOLS) == set(declared_symbols("vits_b200_test.h"))
    engine.load_library()


def test_blackwell_instructions_present(built_lib):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP (B200_PROFILING.md)."""
    exe = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UBLKCP" in sass
    assert "UTMALDG" in sass                      # cp.async.bulk.tensor: the fused stage kernels' input tiles
    assert "sm_100a" in subprocess.run([exe, "-lelf", built_lib], capture_output=True, text=True).stdout


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback(tmp_path, built_lib):
    """The product path must fail loudly without a usable sm_100 device."""
    from phoonnx_b200 import modelgen
    from phoonnx_b200.session import B200Session
    p = str(tmp_path / "v.onnx")
    modelgen.make_voice(p, "tiny")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200Session(p)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "phoonnx_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn
