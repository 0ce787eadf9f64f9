"""ctypes binding of libvits_b200.so (C ABI: include/vits_b200.h).

This is the whole Python<->native boundary: plain pointers and sizes.  There is no CPU
fallback -- if the shared library is missing or no sm_100 GPU is usable, constructing an
``Engine`` raises ``RuntimeError`` (north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Dict, Optional, Sequence

import numpy as np

from .weights import VitsArch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvits_b200.so")

VITS_MAX_UPS, VITS_MAX_RBK, VITS_MAX_DIL, VITS_MAX_FLOWS = 8, 8, 4, 8


class CArch(C.Structure):
    _fields_ = [
        ("n_vocab", C.c_int32), ("hidden", C.c_int32), ("inter", C.c_int32), ("filter", C.c_int32),
        ("n_heads", C.c_int32), ("n_layers", C.c_int32), ("enc_kernel", C.c_int32), ("window", C.c_int32),
        ("n_speakers", C.c_int32), ("gin", C.c_int32),
        ("use_sdp", C.c_int32), ("dp_filter", C.c_int32), ("dp_kernel", C.c_int32), ("dds_layers", C.c_int32),
        ("n_cflows", C.c_int32), ("cflows", C.c_int32 * VITS_MAX_FLOWS), ("num_bins", C.c_int32),
        ("n_flow", C.c_int32), ("flow_layers", C.c_int32 * VITS_MAX_FLOWS), ("wn_layers", C.c_int32),
        ("wn_kernel", C.c_int32), ("wn_dilation_rate", C.c_int32),
        ("resblock_type", C.c_int32),
        ("n_ups", C.c_int32), ("up_rates", C.c_int32 * VITS_MAX_UPS), ("up_kernels", C.c_int32 * VITS_MAX_UPS),
        ("up_init", C.c_int32),
        ("n_rbk", C.c_int32), ("rb_kernels", C.c_int32 * VITS_MAX_RBK), ("rb_ndil", C.c_int32 * VITS_MAX_RBK),
        ("rb_dilations", (C.c_int32 * VITS_MAX_DIL) * VITS_MAX_RBK),
        ("sample_rate", C.c_int32),
    ]


class CInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_vocab", "n_speakers", "has_sid", "hidden", "inter", "sample_rate", "hop",
                                         "resblock_type", "use_sdp", "precision", "device", "num_sms", "finalized",
                                         "has_scales", "has_langid")] + \
               [("reserved", C.c_int32 * 6)]


def to_c_arch(a: VitsArch) -> CArch:
    c = CArch()
    for f in ("n_vocab", "hidden", "inter", "filter", "n_heads", "n_layers", "enc_kernel", "window", "n_speakers",
              "gin", "dp_filter", "dp_kernel", "dds_layers", "num_bins", "wn_layers", "wn_kernel",
              "wn_dilation_rate", "up_init", "sample_rate"):
        setattr(c, f, int(getattr(a, f)))
    c.use_sdp = 1 if a.use_sdp else 0
    if len(a.cflows) > VITS_MAX_FLOWS or len(a.flow_layers) > VITS_MAX_FLOWS or len(a.up_rates) > VITS_MAX_UPS \
            or len(a.rb_kernels) > VITS_MAX_RBK:
        raise ValueError("architecture exceeds the ABI's static limits")
    c.n_cflows = len(a.cflows)
    for i, v in enumerate(a.cflows):
        c.cflows[i] = v
    c.n_flow = len(a.flow_layers)
    for i, v in enumerate(a.flow_layers):
        c.flow_layers[i] = v
    c.resblock_type = 1 if a.resblock == "1" else 2
    c.n_ups = len(a.up_rates)
    for i, (u, k) in enumerate(zip(a.up_rates, a.up_kernels)):
        c.up_rates[i], c.up_kernels[i] = u, k
    c.n_rbk = len(a.rb_kernels)
    for j, (k, dil) in enumerate(zip(a.rb_kernels, a.rb_dilations)):
        if len(dil) > VITS_MAX_DIL:
            raise ValueError("too many dilations per resblock")
        c.rb_kernels[j], c.rb_ndil[j] = k, len(dil)
        for m, d in enumerate(dil):
            c.rb_dilations[j][m] = d
    return c


_lib = None


def load_library(path: Optional[str] = None):
    """dlopen the C-ABI library and declare every symbol of include/vits_b200.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("VITS_B200_LIB") or LIB_PATH      # the env override exists for A/B runs of two builds on one box
    if not os.path.exists(p):
        raise RuntimeError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(p)
    H = C.c_void_p
    lib.vits_abi_version.restype = C.c_int
    lib.vits_create.argtypes = [C.POINTER(CArch), C.c_int, C.POINTER(H)]
    lib.vits_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(H), C.c_char_p, C.c_size_t]
    lib.vits_open.restype = C.c_int
    lib.vits_set_output_offsets.argtypes = [H, C.c_void_p, C.c_int32]
    lib.vits_host_register.argtypes = [C.c_void_p, C.c_size_t]
    lib.vits_host_unregister.argtypes = [C.c_void_p]
    for fn in ("vits_set_output_offsets", "vits_host_register", "vits_host_unregister"):
        getattr(lib, fn).restype = C.c_int
    lib.vits_upload.argtypes = [H, C.c_char_p, C.c_void_p, C.c_size_t, C.c_int]
    lib.vits_finalize.argtypes = [H]
    lib.vits_set_option.argtypes = [H, C.c_char_p, C.c_double]
    lib.vits_prepare.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int64, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_int64)]
    lib.vits_decode.argtypes = [H, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_float, C.c_int32]
    lib.vits_fetch.argtypes = [H, C.c_char_p, C.c_void_p, C.c_int64]
    lib.vits_fetch.restype = C.c_int64
    lib.vits_timer_start.argtypes = [H]
    lib.vits_timer_stop.argtypes = [H, C.POINTER(C.c_float)]
    lib.vits_stage_ms.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.vits_kernel_ms.argtypes = [H, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    lib.vits_launch_count.argtypes = [H]
    lib.vits_launch_count.restype = C.c_int64
    lib.vits_last_error.argtypes = [H]
    lib.vits_last_error.restype = C.c_char_p
    lib.vits_destroy.argtypes = [H]
    lib.vits_destroy.restype = None
    lib.vits_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    lib.vits_host_free.argtypes = [C.c_void_p]
    lib.vits_wait_output.argtypes = [H, C.c_int]
    lib.vits_describe.argtypes = [H, C.POINTER(CInfo)]
    lib.vits_max_output_samples.argtypes = [H, C.c_int64, C.c_float]
    lib.vits_max_output_samples.restype = C.c_int64
    lib.vits_set_stream.argtypes = [H, C.c_void_p]
    lib.vits_output_ticket.argtypes = [H]
    lib.vits_output_ticket.restype = C.c_int64
    lib.vits_wait_ticket.argtypes = [H, C.c_int64]
    for fn in ("vits_describe", "vits_set_stream", "vits_wait_ticket"):
        getattr(lib, fn).restype = C.c_int
    for fn in ("vits_create", "vits_upload", "vits_finalize", "vits_set_option", "vits_prepare", "vits_decode",
               "vits_timer_start", "vits_timer_stop", "vits_stage_ms", "vits_kernel_ms", "vits_host_alloc", "vits_host_free", "vits_wait_output"):
        getattr(lib, fn).restype = C.c_int
    if path is None:
        _lib = lib
    return lib


EXPORTED_SYMBOLS = (
    "vits_abi_version", "vits_create", "vits_upload", "vits_finalize", "vits_set_option", "vits_prepare",
    "vits_decode", "vits_fetch", "vits_timer_start", "vits_timer_stop", "vits_stage_ms", "vits_kernel_ms", "vits_launch_count",
    "vits_last_error", "vits_destroy", "vits_host_alloc", "vits_host_free", "vits_wait_output",
    "vits_describe", "vits_max_output_samples", "vits_set_stream", "vits_output_ticket", "vits_wait_ticket", "vits_open",
    "vits_set_output_offsets", "vits_host_register", "vits_host_unregister",
)
# include/vits_b200_test.h: test-only hooks, not part of the drop-in boundary
TEST_SYMBOLS = ("vits_test_conv", "vits_test_mma_probe", "vits_test_mma_probe_mode", "vits_test_file_arch", "vits_test_file_blob",
                "vits_test_mrf3_plan", "vits_test_conv_plan")
VITS_OUT_ASYNC = 0x100

_DT = {np.dtype(np.float32): 0, np.dtype(np.uint16): 1, np.dtype(np.int32): 2}


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _PinnedBlock:
    """One page-locked host allocation; exposes the buffer protocol numpy needs and goes back to its pool when the
    last array over it is garbage-collected (so the array run() returns is owned by Python, as with onnxruntime)."""

    def __init__(self, pool, ptr: int, nbytes: int):
        self.pool, self.ptr, self.nbytes = pool, ptr, nbytes

    def array(self, n: int, dtype) -> np.ndarray:
        dt = np.dtype(dtype)
        raw = (C.c_uint8 * (n * dt.itemsize)).from_address(self.ptr)
        arr = np.frombuffer(raw, dtype=dt, count=n)
        weakref.finalize(raw, self.pool._give_back, self)       # `raw` lives exactly as long as arrays over it
        return arr


class PinnedPool:
    """Size-classed pool of cudaHostAlloc'd blocks (vits_host_alloc): allocation costs milliseconds, reuse is free.
    Classes are powers of two up to 256 MiB and multiples of 64 MiB above (a 2.2 GB result pins 2.25 GB, not 4)."""

    BIG = 256 << 20
    STEP = 64 << 20

    def __init__(self, lib, max_cached_bytes: int = 8 << 30):
        self.lib, self.free, self.cached, self.max_cached = lib, {}, 0, max_cached_bytes
        self.closed = False

    @classmethod
    def size_class(cls, nbytes: int) -> int:
        if nbytes <= cls.BIG:
            return 1 << max(16, (nbytes - 1).bit_length())
        return (nbytes + cls.STEP - 1) // cls.STEP * cls.STEP

    def take(self, n: int, dtype) -> np.ndarray:
        nbytes = max(1, n * np.dtype(dtype).itemsize)
        cls = self.size_class(nbytes)
        lst = self.free.get(cls)
        if lst:
            blk = lst.pop()
            self.cached -= cls
        else:
            p = C.c_void_p()
            if self.lib.vits_host_alloc(cls, C.byref(p)) != 0 or not p.value:
                return np.empty((n,), dtype)                    # pageable fallback: still correct, just a staged copy
            blk = _PinnedBlock(self, p.value, cls)
        return blk.array(n, dtype)

    def _give_back(self, blk):
        # after close() nothing is cached any more: a block whose last array dies later is freed on the spot
        if self.closed or self.cached + blk.nbytes > self.max_cached:
            self.lib.vits_host_free(C.c_void_p(blk.ptr))
            return
        self.free.setdefault(blk.nbytes, []).append(blk)
        self.cached += blk.nbytes

    def drain(self, close: bool = False):
        if close:
            self.closed = True
        for lst in self.free.values():
            for blk in lst:
                self.lib.vits_host_free(C.c_void_p(blk.ptr))
        self.free, self.cached = {}, 0


class Engine:
    """One GPU, one stream, one voice."""

    def __init__(self, arch: VitsArch, blobs: Dict[str, np.ndarray], options: Dict[str, float],
                 device: int = 0, precision: str = "fp32"):
        self.lib = load_library()
        self.arch = arch
        self.hop = arch.hop
        self._h = C.c_void_p()
        ca = to_c_arch(arch)
        rc = self.lib.vits_create(C.byref(ca), int(device), C.byref(self._h))
        if rc != 0 or not self._h:
            self._h = C.c_void_p()
            raise RuntimeError(f"vits_create failed (code {rc}): no usable sm_100 CUDA device {device}; "
                               "this engine has no CPU fallback")
        for name, arr in blobs.items():
            arr = np.ascontiguousarray(arr)
            self._check(self.lib.vits_upload(self._h, name.encode(), _ptr(arr), arr.nbytes, _DT[arr.dtype]))
        for k, v in options.items():
            self.set_option(k, v)
        self.set_precision(precision)
        self._check(self.lib.vits_finalize(self._h))
        self._B = 0
        self._ylen = None
        self._pool = PinnedPool(self.lib)
        self.last_ticket = 0

    @classmethod
    def open(cls, path: str, device: int = 0, precision: str = "fp32") -> "Engine":
        """The library opens the voice file by itself (vits_open: C++ reader, architecture inference and packing inside the
        .so); this class then only needs what vits_describe reports."""
        from .weights import VitsArch
        lib = load_library()
        self = cls.__new__(cls)
        self.lib = lib
        self._h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = lib.vits_open(str(path).encode(), int(device), 0 if precision == "fp32" else 1, C.byref(self._h), err, 512)
        if rc != 0 or not self._h:
            self._h = C.c_void_p()
            msg = err.value.decode("utf-8", "replace")
            if rc == -1:
                raise ValueError(msg)
            raise RuntimeError(f"vits_open failed (code {rc}): {msg}")
        self.precision = precision
        self._B, self._ylen = 0, None
        self._pool = PinnedPool(lib)
        self.last_ticket = 0
        info = self.describe()
        self.info = info
        self.hop = info["hop"]
        self.arch = VitsArch(n_vocab=info["n_vocab"], hidden=info["hidden"], inter=info["inter"], n_speakers=info["n_speakers"],
                             sample_rate=info["sample_rate"], use_sdp=bool(info["use_sdp"]), resblock=str(info["resblock_type"]))
        return self

    # ------------------------------------------------------------------
    def _check(self, rc: int):
        if rc == 0:
            return
        msg = self.lib.vits_last_error(self._h).decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise RuntimeError(f"libvits_b200 error {rc}: {msg}")

    def set_option(self, key: str, value: float):
        self._check(self.lib.vits_set_option(self._h, key.encode(), float(value)))

    def set_precision(self, precision: str):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        self.set_option("precision", 0 if precision == "fp32" else 1)

    def prepare(self, ids: np.ndarray, lengths: np.ndarray, scales: Sequence[float], sid: Optional[np.ndarray] = None,
                noise_dp: Optional[np.ndarray] = None, logw: Optional[np.ndarray] = None, seed: int = 0) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
        lengths = np.ascontiguousarray(lengths, dtype=np.int64).reshape(-1)
        B = int(lengths.shape[0])
        if B < 1 or int(lengths.sum()) != ids.shape[0]:
            raise ValueError("ids must hold exactly sum(lengths) values (packed) and B >= 1")
        sc = np.ascontiguousarray(scales, dtype=np.float32).reshape(-1)
        if sc.shape[0] != 3:
            raise ValueError("scales must be [noise_scale, length_scale, noise_w]")
        sid_a = None if sid is None else np.ascontiguousarray(sid, dtype=np.int64).reshape(-1)
        if sid_a is not None and sid_a.shape[0] != B:
            raise ValueError("sid must have one entry per utterance")
        nd, stride = None, 0
        if noise_dp is not None:
            nd = np.ascontiguousarray(noise_dp, dtype=np.float32)
            if nd.ndim != 3 or nd.shape[0] != B or nd.shape[1] != 2:
                raise ValueError("noise_dp must be [B, 2, T]")
            stride = nd.shape[2]
        lw = None
        if logw is not None:
            lw = np.ascontiguousarray(logw, dtype=np.float32).reshape(-1)
            if lw.shape[0] != ids.shape[0]:
                raise ValueError("logw override must be packed [sum(lengths)]")
        ylen = np.zeros((B,), np.int64)
        tot = C.c_int64(0)
        self._check(self.lib.vits_prepare(self._h, _ptr(ids), _ptr(lengths), B, _ptr(sc), _ptr(sid_a), _ptr(nd),
                                          stride, _ptr(lw), C.c_uint64(seed & (2 ** 64 - 1)), _ptr(ylen), C.byref(tot)))
        self._B, self._ylen, self._R = B, ylen, int(ids.shape[0])
        return ylen

    def decode(self, noise_z: Optional[np.ndarray] = None, out: str = "f32", volume: float = 1.0,
               normalize: bool = True, asynchronous: bool = False, dest: Optional[np.ndarray] = None,
               dest_offsets=None) -> Optional[np.ndarray]:
        """``asynchronous=True`` (host outputs only): returns as soon as the device->host transfer is enqueued; the array is
        complete after ``wait_ticket(self.last_ticket)``.  It is a per-call flag -- a blocking decode() issued while an
        asynchronous result is still in flight waits for ITS OWN buffer only and does not disturb the other one."""
        if self._ylen is None:
            raise RuntimeError("decode() before prepare()")
        total = int(self._ylen.sum()) * self.hop
        nz, stride = None, 0
        if noise_z is not None:
            nz = np.ascontiguousarray(noise_z, dtype=np.float32)
            if nz.ndim != 3 or nz.shape[0] != self._B or nz.shape[1] != self.arch.inter:
                raise ValueError("noise_z must be [B, inter_channels, T_y]")
            stride = nz.shape[2]
        if out == "none":
            self._check(self.lib.vits_decode(self._h, _ptr(nz), stride, 0, None, 0, volume, int(normalize)))
            return None
        if out not in ("f32", "i16"):
            raise ValueError("out must be 'none', 'f32' or 'i16'")
        # page-locked result: the device->host transfer is a DMA on the copy stream, chunk by chunk
        kind = (1 if out == "f32" else 2) | (VITS_OUT_ASYNC if asynchronous else 0)
        if dest is not None:
            # scatter output: utterance b lands at dest[dest_offsets[b]:...] (vits_set_output_offsets); `dest` is the caller's own
            # buffer (page-lock it with host_register() for asynchronous DMA) and is returned as is
            offs = np.ascontiguousarray(dest_offsets, dtype=np.int64).reshape(-1)
            if offs.shape[0] != self._B or dest.dtype != (np.float32 if out == "f32" else np.int16) or not dest.flags["C_CONTIGUOUS"]:
                raise ValueError("dest must be a contiguous array of the output dtype and dest_offsets one sample offset per utterance")
            self._check(self.lib.vits_set_output_offsets(self._h, _ptr(offs), self._B))
            self._check(self.lib.vits_decode(self._h, _ptr(nz), stride, kind, _ptr(dest), int(dest.size), volume, int(normalize)))
            self.last_ticket = int(self.lib.vits_output_ticket(self._h))
            return dest
        buf = self._pool.take(total, np.float32 if out == "f32" else np.int16)
        self._check(self.lib.vits_decode(self._h, _ptr(nz), stride, kind, _ptr(buf), total, volume, int(normalize)))
        self.last_ticket = int(self.lib.vits_output_ticket(self._h))
        return buf

    def host_register(self, arr: np.ndarray):
        """Page-lock memory the caller owns (e.g. a shared-memory segment) so DMAs into it are asynchronous."""
        self._check(self.lib.vits_host_register(_ptr(arr), arr.nbytes))

    def host_unregister(self, arr: np.ndarray):
        self.lib.vits_host_unregister(_ptr(arr))

    def decode_to_device(self, dev_ptr: int, capacity: int, noise_z: Optional[np.ndarray] = None):
        """float32 audio into a caller-owned DEVICE buffer (out_kind 3): stream-ordered, no host synchronisation."""
        if self._ylen is None:
            raise RuntimeError("decode() before prepare()")
        nz, stride = None, 0
        if noise_z is not None:
            nz = np.ascontiguousarray(noise_z, dtype=np.float32)
            stride = nz.shape[2]
        self._check(self.lib.vits_decode(self._h, _ptr(nz), stride, 3, C.c_void_p(int(dev_ptr)), int(capacity), 1.0, 1))

    def wait_ticket(self, ticket: int):
        self._check(self.lib.vits_wait_ticket(self._h, int(ticket)))

    def set_stream(self, cuda_stream: Optional[int]):
        """Run on the caller's CUDA stream (``torch.cuda.Stream.cuda_stream`` or any cudaStream_t as an int); None restores."""
        self._check(self.lib.vits_set_stream(self._h, C.c_void_p(int(cuda_stream)) if cuda_stream else None))

    def describe(self) -> Dict[str, int]:
        info = CInfo()
        self._check(self.lib.vits_describe(self._h, C.byref(info)))
        return {n: int(getattr(info, n)) for n, _ in CInfo._fields_ if n != "reserved"}

    def max_output_samples(self, sum_ids: int = 0, length_scale: float = 1.0) -> int:
        n = int(self.lib.vits_max_output_samples(self._h, int(sum_ids), float(length_scale)))
        if n < 0:
            self._check(n)
        return n

    def wait_output(self, older_only: bool = False):
        self._check(self.lib.vits_wait_output(self._h, 1 if older_only else 0))

    def fetch(self, name: str) -> np.ndarray:
        a = self.arch
        frames = int(self._ylen.sum())
        cap = {"x": self._R * a.hidden, "stats": self._R * 2 * a.inter, "logw": self._R, "durations": self._R,
               "cum": self._R, "noise_dp": 2 * self._R,
               # per-chunk workspaces: the C side refuses them after a multi-chunk decode (a fragment would be silently wrong)
               "frame_index": frames, "z_p": frames * a.inter, "z": frames * a.inter}[name]
        is_int = name in ("durations", "cum", "frame_index")
        buf = np.empty((cap,), np.int32 if is_int else np.float32)
        n = self.lib.vits_fetch(self._h, name.encode(), _ptr(buf), cap)
        if n < 0:
            self._check(int(n))
        return buf[:n]

    def timer_start(self):
        self._check(self.lib.vits_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        self._check(self.lib.vits_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def stage_ms(self):
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        self._check(self.lib.vits_stage_ms(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"text": a.value, "flow": b.value, "dec": c.value}

    def kernel_ms(self, which: int):
        """(device ms, launches, algorithmic MACs) of one fused decoder kernel since timer_start(): 0 = last stage, 1 = the others."""
        ms, n, mac = C.c_float(), C.c_int64(), C.c_double()
        self._check(self.lib.vits_kernel_ms(self._h, which, C.byref(ms), C.byref(n), C.byref(mac)))
        return float(ms.value), int(n.value), float(mac.value)

    def launch_count(self) -> int:
        return int(self.lib.vits_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.vits_destroy(self._h)
            self._h = C.c_void_p()
            self._pool.drain(close=True)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
