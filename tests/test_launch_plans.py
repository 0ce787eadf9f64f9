"""Host-side launch plans of the tensor-core kernels (mrf3_plan, conv_tc_plan), through the test hooks of include/vits_b200_test.h.
Plain host code: runs without a GPU.  What is checked is what the kernels silently rely on: the shared-memory budget, the tensor-memory
column count, the TMA box geometry (whole boxes of a multiple of 8 rows, at most 256), and that all ResBlock1 pairs of one stage step
through the utterances with ONE tile table."""
import ctypes as C

import numpy as np
import pytest

SMEM_LIMIT = 227 * 1024
MRF_KEYS = ("nb", "span", "hmax", "h1max", "t_out", "t_step", "rx", "rx1", "smem_bytes", "tmem_cols", "nstages", "resident", "tma", "nboxes",
            "box_rows", "u_rows")
CONV_KEYS = ("ntile", "rows_a", "rows_need", "a_bytes", "slot_bytes", "nstages", "resident", "nabuf", "smem_bytes", "tmem_cols", "tma",
             "nboxes", "box_rows", "nepi", "nload", "naccbuf")


@pytest.fixture(scope="module")
def lib(built_lib):
    lib = C.CDLL(built_lib)
    lib.vits_test_mrf3_plan.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p]
    lib.vits_test_conv_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return lib


def mrf_plan(lib, Cc, ks, d1, d2, rb1=0, up_cin=0, nb_pref=4, fuse_post=0, use_tma=1, min_hmax=0):
    k, a, b = (np.asarray(v, np.int32) for v in (ks, d1, d2))
    out = np.zeros(16, np.int32)
    ok = lib.vits_test_mrf3_plan(Cc, len(ks), k.ctypes.data, a.ctypes.data, b.ctypes.data, rb1, up_cin, nb_pref, fuse_post, use_tma, min_hmax,
                                 out.ctypes.data)
    return dict(zip(MRF_KEYS, out.tolist())) if ok else None


def conv_plan(lib, cin, n, toff, xb=0, xb_rows=0, split3=0, ntiles=10000, num_sms=148):
    t = np.asarray(toff, np.int32)
    out = np.zeros(16, np.int32)
    ok = lib.vits_test_conv_plan(cin, n, len(toff), t.ctypes.data, xb, xb_rows, split3, ntiles, num_sms, out.ctypes.data)
    return dict(zip(CONV_KEYS, out.tolist())) if ok else None


def check_mrf(p, Cc, nrb):
    assert p["smem_bytes"] <= SMEM_LIMIT and p["tmem_cols"] <= 512 and (nrb + 1) * p["nb"] * Cc <= p["tmem_cols"]
    assert p["span"] == 128 * p["nb"] and p["t_out"] == p["span"] - 2 * p["hmax"] and 32 <= p["t_step"] <= p["t_out"]
    assert p["rx1"] >= p["span"] + 2 * p["hmax"]
    assert p["nstages"] >= 3 or p["resident"]
    if p["tma"]:
        rows = p["u_rows"] if p["u_rows"] else p["rx"]
        assert rows == p["nboxes"] * p["box_rows"] and p["box_rows"] % 8 == 0 and 8 <= p["box_rows"] <= 256


def test_medium_voice_stages(lib):
    """The ResBlock2 stages of the exported voices (kernels 3/5/7, dilations (1,2) (2,6) (3,12)): 64 channels, and 32 channels with the fused
    ConvTranspose (from 64 channels) and conv_post."""
    ks, d1, d2 = (3, 5, 7), (1, 2, 3), (2, 6, 12)
    p64 = mrf_plan(lib, 64, ks, d1, d2, nb_pref=2)
    check_mrf(p64, 64, 3)
    assert p64["nb"] == 2 and p64["hmax"] == 36 and p64["h1max"] == 9 and p64["rx"] >= p64["span"] + 18
    p32 = mrf_plan(lib, 32, ks, d1, d2, up_cin=64, nb_pref=4, fuse_post=1)
    check_mrf(p32, 32, 3)
    assert p32["nb"] == 4 and p32["t_step"] == p32["t_out"] - 6 and p32["resident"] == 1 and p32["u_rows"] > 0
    # the cp.async loader's tiles have an odd number of rows (bank spread); the TMA-fed ones whole boxes
    q64 = mrf_plan(lib, 64, ks, d1, d2, nb_pref=2, use_tma=0)
    assert q64["tma"] == 0 and q64["rx"] % 2 == 1 and q64["t_step"] == p64["t_step"]


@pytest.mark.parametrize("Cc,nb", [(32, 4), (64, 2), (128, 1)])
def test_resblock1_pairs_of_a_stage_share_one_tile_table(lib, Cc, nb):
    """`high` preset: kernels 3/7/11 x dilations 1/3/5, second conv undilated; every pair is planned with the halo of the widest second
    conv so that the nine kernels of a stage use the same t_step."""
    hmax = (11 - 1) // 2
    steps = set()
    for k in (3, 7, 11):
        for d in (1, 3, 5):
            p = mrf_plan(lib, Cc, (k,), (d,), (1,), rb1=1, nb_pref=nb, min_hmax=hmax)
            assert p is not None, (Cc, k, d)
            check_mrf(p, Cc, 1)
            assert p["nb"] == nb and p["hmax"] == hmax and p["h1max"] == d * (k - 1) // 2
            steps.add(p["t_step"])
    assert steps == {128 * nb - 2 * hmax}
    # a ResBlock2-style stage of 128 channels is not a shape of the fused kernel (only ResBlock1 pairs are)
    assert mrf_plan(lib, 128, (3, 5, 7), (1, 2, 3), (2, 6, 12), nb_pref=2) is None
    assert mrf_plan(lib, 128, (3,), (1,), (1,), rb1=0, nb_pref=1) is None


def test_unsupported_shapes_are_refused(lib):
    assert mrf_plan(lib, 48, (3,), (1,), (1,)) is None                       # channel count
    assert mrf_plan(lib, 32, (4,), (1,), (1,)) is None                       # even kernel
    assert mrf_plan(lib, 32, (3, 5), (1, 2), (2, 6), rb1=1) is None          # a pair is one resblock
    assert mrf_plan(lib, 32, (3,), (1,), (1,), up_cin=64) is None            # fused ConvTranspose needs the conv1 buffers of >= 2 resblocks


def test_conv_plans(lib):
    taps3, taps5 = (-1, 0, 1), (-2, -1, 0, 1, 2)
    for cin, n, toff, split3 in ((192, 576, (0,), 1), (192, 768, taps3, 1), (768, 192, taps3, 1), (192, 384, taps5, 0), (128, 128, (-36, -24, -12, 0, 12, 24, 36), 0),
                                 (192, 32, (0,), 1), (256, 512, (-1, 0), 0)):
        cin_k = 96 if split3 else cin                       # bf16x3 convs are planned per 96-channel K slice
        p = conv_plan(lib, cin_k, n, toff, split3=split3)
        assert p is not None, (cin, n)
        assert p["smem_bytes"] <= SMEM_LIMIT and p["tmem_cols"] <= 512 and p["naccbuf"] * p["ntile"] <= p["tmem_cols"]
        assert n % p["ntile"] == 0 or p["ntile"] == ((n + 15) // 16) * 16
        assert p["rows_need"] == 128 + (max(toff) - min(toff)) and p["rows_a"] >= p["rows_need"]
        assert p["nstages"] >= 2 or p["resident"]
        assert p["nepi"] + p["nload"] == 14
    # latency mode: a launch of a few tiles spreads its output columns over narrower N tiles (same K order: bit-identical results)
    wide = conv_plan(lib, 96, 768, taps3, split3=1, ntiles=5000)
    narrow = conv_plan(lib, 96, 768, taps3, split3=1, ntiles=1)
    assert narrow["ntile"] < wide["ntile"] and narrow["ntile"] >= 32 and 768 % narrow["ntile"] == 0
