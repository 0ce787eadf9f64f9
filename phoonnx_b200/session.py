"""B200Session: the object that replaces ``onnxruntime.InferenceSession`` inside phoonnx's
``TTSVoice`` (reference seam: phoonnx/voice.py:105-107; touched at :347 ``get_inputs()`` and
:374-377 ``run(None, feed)``; constructed at :167-171).

Same constructor shape ``(path, sess_options=None, providers=None)``, same feed names and
dtypes (``input`` int64 [B,T], ``input_lengths`` int64 [B], ``scales`` float32 [3] =
[noise_scale, length_scale, noise_w], ``sid`` int64 [B] for multi-speaker voices; unknown
keys such as ``langid`` are rejected by the caller's own filter, voice.py:373), same output:
a list whose first element is float32 ``[B, 1, 1, T_max * hop]`` (export_onnx.py:269-278).

Semantics: every utterance is synthesised exactly as if it were alone in the batch (B=1,
which is what TTSVoice always sends, voice.py:350-351); rows are zero beyond their own length.
Errors: bad shapes / ids / sid -> ValueError, CUDA problems -> RuntimeError.  No CPU fallback.
"""
from __future__ import annotations

import threading
from typing import Dict, List, Optional, Sequence

import numpy as np

from .engine import Engine
from .packing import pack_model
from .weights import load_model


class _NoLock:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_LOCK = _NoLock()


class NodeArg:
    """Mirror of onnxruntime.NodeArg as far as TTSVoice looks (``.name``; voice.py:347)."""

    def __init__(self, name: str, type_: str, shape):
        self.name, self.type, self.shape = name, type_, shape

    def __repr__(self):
        return f"NodeArg(name='{self.name}', type='{self.type}', shape={self.shape})"


class B200Session:
    TEST_FEEDS = ("noise_dp", "noise_z", "logw")
    _has_scales = True                                   # the exporter's graph declares `scales` (export_onnx.py:300-304)
    default_scales = np.asarray((0.667, 1.0, 0.8), np.float32)     # config.py:9-11

    def __init__(self, path_or_bytes, sess_options=None, providers=None, provider_options=None, *,
                 device: int = 0, precision: str = "fp32", sample_rate: Optional[int] = None,
                 max_chunk_frames: Optional[int] = None, seed: int = 0, default_scales=None,
                 native_loader: bool = False):
        self._path = str(path_or_bytes)
        if native_loader:
            # the library opens the file itself (vits_open); the Python loader is not involved
            self.engine = Engine.open(self._path, device=device, precision=precision)
            arch = self.engine.arch
            if sample_rate:
                arch.sample_rate = int(sample_rate)
            info = self.engine.info
            hdr = type("Hdr", (), {"metadata": {}, "inputs": ["input", "input_lengths"] + (["scales"] if info["has_scales"] else []) +
                                   (["sid"] if info["has_sid"] else []) + (["langid"] if info["has_langid"] else [])})()
        else:
            W, arch, hdr = load_model(self._path, sample_rate)
            blobs, opts = pack_model(W, arch)
            self.engine = Engine(arch, blobs, opts, device=device, precision=precision)
        self.arch = arch
        self.metadata = dict(hdr.metadata)
        if max_chunk_frames:
            self.engine.set_option("max_chunk_frames", max_chunk_frames)
        # the inputs THIS file's graph declares, in its order (voice.py:347 reads exactly this list and filters its feed by it,
        # voice.py:358-373): voices exported without a `scales` input get none (their scales are constants of the graph: the
        # session then uses `default_scales`), single-language voices that still declare `langid` get it accepted and ignored
        known = {"input": ("tensor(int64)", ["batch_size", "phonemes"]), "input_lengths": ("tensor(int64)", ["batch_size"]),
                 "scales": ("tensor(float)", [3]), "sid": ("tensor(int64)", ["batch_size"]), "langid": ("tensor(int64)", ["batch_size"])}
        declared = [n for n in (hdr.inputs or []) if n in known] or ["input", "input_lengths", "scales"]
        for must in ("input", "input_lengths"):
            if must not in declared:
                raise ValueError(f"voice file declares no '{must}' input: not a VITS voice this engine can run")
        if arch.n_speakers > 1 and "sid" not in declared:
            declared.append("sid")
        self._inputs = [NodeArg(n, known[n][0], known[n][1]) for n in declared if n != "sid" or arch.n_speakers > 1]
        self._has_scales = "scales" in declared
        self.default_scales = np.asarray(default_scales if default_scales is not None else (0.667, 1.0, 0.8), np.float32)   # config.py:9-11
        self._outputs = [NodeArg("output", "tensor(float)", ["batch_size", 1, 1, "time"])]
        self._seed = int(seed)
        self._calls = 0
        self._lock = threading.RLock()                   # run() is re-entrant like ORT's (voice.py:374): calls on one session are serialised
        self.last_lengths: Optional[np.ndarray] = None   # samples per utterance of the last run

    # ---- the InferenceSession surface TTSVoice uses ---------------------------------
    def get_inputs(self) -> List[NodeArg]:
        return list(self._inputs)

    def get_outputs(self) -> List[NodeArg]:
        return list(self._outputs)

    def get_providers(self) -> List[str]:
        return ["B200ExecutionProvider"]

    def run(self, output_names: Optional[Sequence[str]], input_feed: Dict[str, np.ndarray], run_options=None):
        if output_names is not None and list(output_names) not in ([], ["output"]):
            raise ValueError(f"unknown output names {output_names!r}; the graph has one output: 'output'")
        audio, lens = self.synthesize_packed(input_feed)
        B = lens.shape[0]
        tmax = int(lens.max()) if B else 0
        out = np.zeros((B, 1, 1, tmax), np.float32)
        off = 0
        for b in range(B):
            n = int(lens[b])
            out[b, 0, 0, :n] = audio[off:off + n]
            off += n
        return [out]

    # ---- batched entry used by run() and by the throughput harness ------------------
    def _unpack_feed(self, feed):
        known = {a.name for a in self._inputs} | set(self.TEST_FEEDS)
        for k in feed:
            if k not in known:
                raise ValueError(f"Invalid input name: {k}")
        for a in self._inputs:
            if a.name not in feed and a.name != "langid":
                raise ValueError(f"Required input '{a.name}' is missing from the feed")
        x = np.asarray(feed["input"])
        lens = np.asarray(feed["input_lengths"])
        if x.dtype != np.int64 or lens.dtype != np.int64:
            raise ValueError("'input' and 'input_lengths' must be int64 (voice.py:350-351)")
        if x.ndim != 2 or lens.ndim != 1 or lens.shape[0] != x.shape[0]:
            raise ValueError(f"expected input [B,T] and input_lengths [B], got {x.shape} / {lens.shape}")
        if x.shape[0] < 1 or x.shape[1] < 1:
            raise ValueError("empty batch / empty phoneme sequence")
        if lens.min() < 1 or lens.max() > x.shape[1]:
            raise ValueError("input_lengths must lie in [1, T]")
        if self._has_scales:
            scales = np.asarray(feed["scales"])
            if scales.dtype != np.float32 or scales.shape != (3,):
                raise ValueError("'scales' must be float32 [3] = [noise_scale, length_scale, noise_w] (voice.py:364-367)")
        else:
            scales = self.default_scales                 # the graph has no such input (voice.py:358): its scales are fixed
        if "langid" in feed:
            lg = np.asarray(feed["langid"])              # single-language voice: accepted (voice.py:369), nothing to select
            if lg.dtype != np.int64 or lg.shape != (x.shape[0],) or (lg != 0).any():
                raise ValueError("'langid' must be int64 [B] of zeros: this voice has one language")
        sid = None
        if self.arch.n_speakers > 1:
            sid = np.asarray(feed["sid"])
            if sid.dtype != np.int64 or sid.shape != (x.shape[0],):
                raise ValueError("'sid' must be int64 [B] (voice.py:370)")
        return x, lens, scales, sid

    def synthesize_packed(self, feed, out: str = "f32", volume: float = 1.0, normalize: bool = True, asynchronous: bool = False):
        """Returns (packed audio of all utterances, samples per utterance).  ``asynchronous=True``: the array is still being
        filled by the copy stream when this returns; ``self.engine.wait_ticket(self.engine.last_ticket)`` completes it."""
        x, lens, scales, sid = self._unpack_feed(feed)
        B, T = x.shape
        mask = np.arange(T)[None, :] < lens[:, None]
        ids = x[mask]                                    # packed, positions >= length ignored (commons.py:109-113)
        with getattr(self, "_lock", _NO_LOCK):           # text side + frame side of ONE call must not interleave with another thread's
            self._calls += 1
            ylen = self.engine.prepare(ids, lens, scales, sid, feed.get("noise_dp"), feed.get("logw"),
                                       seed=self._seed + self._calls)
            audio = self.engine.decode(feed.get("noise_z"), out=out, volume=volume, normalize=normalize, asynchronous=asynchronous)
            self.last_lengths = ylen * self.engine.hop
            return audio, self.last_lengths

    def prepare_feed(self, feed) -> np.ndarray:
        """Text side only (vits_prepare): returns the samples per utterance the following ``decode_prepared`` will produce -- the
        two-phase form a multi-process job uses to agree on every utterance's place in a shared result buffer before any audio
        exists (bench.py --scaling strong)."""
        x, lens, scales, sid = self._unpack_feed(feed)
        mask = np.arange(x.shape[1])[None, :] < lens[:, None]
        self._calls += 1
        ylen = self.engine.prepare(x[mask], lens, scales, sid, feed.get("noise_dp"), feed.get("logw"), seed=self._seed + self._calls)
        self.last_lengths = ylen * self.engine.hop
        return self.last_lengths

    def decode_prepared(self, feed=None, out: str = "f32", dest: Optional[np.ndarray] = None, dest_offsets=None, asynchronous: bool = False,
                        volume: float = 1.0, normalize: bool = True):
        """Frame side of the batch ``prepare_feed`` prepared; with ``dest`` / ``dest_offsets`` every utterance is DMA'd straight to
        its own place in the caller's buffer."""
        return self.engine.decode(None if feed is None else feed.get("noise_z"), out=out, volume=volume, normalize=normalize,
                                  asynchronous=asynchronous, dest=dest, dest_offsets=dest_offsets)

    def synthesize_many(self, feeds, out: str = "f32", volume: float = 1.0, normalize: bool = True):
        """Batched form of the serial loop in ``TTSVoice.synthesize`` (voice.py:265-269): yields
        ``(packed audio, samples per utterance)`` per feed, in order, with the device->host transfer of batch k
        overlapping the kernels of batch k+1 (page-locked results, copy stream; include/vits_b200.h).

        Asynchrony is a per-call flag and every result is waited for by ITS OWN ticket, so ``run()`` /
        ``synthesize_packed()`` calls made on this session from inside the consumer loop stay blocking and correct, and a
        feed that raises does not lose the result that was already in flight (it is yielded before the error propagates)."""
        eng = self.engine
        prev = None                                     # (audio, lengths, ticket) of the batch whose transfer is in flight
        try:
            for feed in feeds:
                try:
                    audio, alen = self.synthesize_packed(feed, out=out, volume=volume, normalize=normalize, asynchronous=True)
                except Exception:
                    if prev is not None:
                        p, prev = prev, None
                        eng.wait_ticket(p[2])
                        yield p[0], p[1]
                    raise
                cur = (audio, alen, eng.last_ticket)
                if prev is not None:
                    p, prev = prev, cur
                    eng.wait_ticket(p[2])
                    yield p[0], p[1]
                prev = cur
            if prev is not None:
                p, prev = prev, None
                eng.wait_ticket(p[2])
                yield p[0], p[1]
        finally:
            if prev is not None:                        # consumer abandoned the generator: do not leave a DMA into a dead buffer
                eng.wait_ticket(prev[2])

    def end_profiling(self):
        return None
