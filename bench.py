#!/usr/bin/env python
"""Benchmark of the phoneme-ids -> audio hot path (BASELINE.json metric: audio-seconds synthesised per second at 1/2/4/8 B200 +
HiFi-GAN decoder fraction of tensor peak).

  python bench.py --gpus N --steps K --warmup W                  # this engine, BASELINE config 5 (the driver's line)
  python bench.py --config C1|C2|C3|C4|C5 ...                    # every BASELINE.json config; lines under profiles/
  python bench.py --scaling strong --gpus N ...                  # C5 as ONE 4096-utterance job sharded over the ranks
  python bench.py --impl reference --gpus N ...                  # the reference's own SynthesizerTrn.infer on the host cores

Workloads (random-init voices written by phoonnx_b200.modelgen in the exporter's file format):
  C1  medium, ONE utterance of 128 ids per run() -- what TTSVoice sends (voice.py:350-351): per-call latency p50 / p99
  C2  x_low, 32 utterances of randint(64,257) ids (seed 0), one device batch per step
  C3  medium, 8 speakers, 64 utterances (seed 1), sid = arange % 8
  C4  high (ResBlock1, kernels 3/7/11), 16 utterances of 512 ids
  C5  medium, 4096 utterances of randint(64,257) ids (seed 2), length-bucketed.  Default ("weak"): every rank runs its own 4096
      (seed 2 + rank) -- per-GPU work fixed, the line the driver compares across rounds.  "strong": ONE list (seed 2) dealt to the
      ranks by scheduler.plan(lengths, world, rank); the end-to-end number then includes the host gather: every rank's audio lands
      in ONE host buffer (POSIX shared memory, owned by rank 0) in original utterance order.  No collective on the data path
      (SURVEY.md 8e): NCCL carries the start/stop barriers and the max-over-ranks of the timings only; the gather's hand-shake
      is a flag array in the same shared segment.

One JSON line on stdout (rank 0).  `value` = audio-seconds / second with inputs resident on the device (CUDA events on the engine's
stream, max over ranks); `e2e` = the same through the session call with HOST buffers (ids in, float32 audio out, copies inside the
timed region; median of three passes).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCALES = (0.667, 1.0, 0.8)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

CONFIGS = {
    # name: preset, speakers, utterances, (lo, hi) ids, lengths seed, description
    "C1": dict(preset="medium", ns=1, utts=1, lo=128, hi=129, seed=0, what="single utterance of 128 phoneme ids per call (latency)"),
    # `repeat`: passes over the batch per step -- a 3 ms batch timed alone is at the mercy of one host hiccup (a step is >= 50 ms)
    "C2": dict(preset="x_low", ns=1, utts=32, lo=64, hi=257, seed=0, repeat=16, what="batch of 32 utterances of randint(64,257) ids"),
    "C3": dict(preset="medium", ns=8, utts=64, lo=64, hi=257, seed=1, repeat=8, what="8 speakers, batch of 64 utterances of randint(64,257) ids"),
    "C4": dict(preset="high", ns=1, utts=16, lo=512, hi=513, seed=0, repeat=2, what="ResBlock1 decoder, batch of 16 utterances of 512 ids"),
    "C5": dict(preset="medium", ns=1, utts=4096, lo=64, hi=257, seed=2, what="utterances of randint(64,257) ids, length-bucketed"),
}

# DRAM bytes per FRAME of the two fused kernels (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch over a
# 262 144-frame chunk, divided by its frames) -- copied from the committed capture, not measured by this script (a number taken
# under a profiler is never a bench value).  bench.py scales them by the frames per launch it measured.
KERNEL_TRAFFIC = {
    "k_mrf3_tc<32>": {"bytes_per_frame": (2.150258e9 + 0.264484e9) / 262144, "algorithmic_bytes_per_frame": 64 * 64 * 2 + 256 * 4,
                      "source": "profiles/r01g_ncu_mrf3.txt"},
    "k_mrf3_tc<64>": {"bytes_per_frame": (2.150496e9 + 2.104694e9) / 262144, "algorithmic_bytes_per_frame": 64 * 64 * 2 * 2,
                      "source": "profiles/r01g_ncu_mrf3.txt"},
}


def make_workload(n_utts: int, seed: int, n_vocab: int = 256, lo: int = 64, hi: int = 257):
    rs = np.random.RandomState(seed)
    lengths = rs.randint(lo, hi, size=(n_utts,)).astype(np.int64)
    utts = [rs.randint(0, n_vocab, size=(int(L),)).astype(np.int64) for L in lengths]
    return utts, lengths


class ClockSampler:
    """SM clock / throttle reasons of this rank's GPU during the timed region (B200_PROFILING.md's clocks line), read through NVML
    in a thread of this process.  (A polling `nvidia-smi -lms` child per rank re-enumerates every GPU of the box on each sample and
    contends for the driver lock with the kernel launches: with two ranks it slowed the timed loop 2.4x -- r01 2-GPU run.)"""

    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20))

    def __init__(self, device: int, period_s: float = 0.1):
        self.device, self.period = device, period_s
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self.t = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")      # NVML enumerates physical GPUs
            phys = self.device
            if vis:
                ent = [v.strip() for v in vis.split(",") if v.strip()]
                if self.device < len(ent) and ent[self.device].isdigit():
                    phys = int(ent[self.device])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        except Exception as e:          # no NVML: say so in the JSON instead of guessing
            self.err = f"NVML unavailable: {e}"

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.REASONS:
                    if r & bit:
                        self.reasons.add(name)
            except Exception as e:
                self.err = str(e)
                return
            self._stop.wait(self.period)

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"]}
        self._stop.set()
        self.t.join(timeout=2)
        busy = [s for s in self.sm if s > 0]
        out = {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(self.mx) if self.mx else None,
               "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "NVML, this process, 100 ms period"}
        if self.err:
            out["error"] = self.err
        return out


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


# ---------------------------------------------------------------------------------------------------------------------------
# Reference arm: the UNMODIFIED reference (pip-installed into baseline/_ref from /root/reference), its own
# SynthesizerTrn.infer (models.py:681-722 -- the graph export_onnx.py traces and onnxruntime would run; onnxruntime itself is not
# installable offline) on the host cores, B=1 per call like TTSVoice (voice.py:265-269, 350-351), in a pool of worker processes
# of REF_THREADS torch threads each (a single B=1 process cannot use a whole host: r01 timed exactly that and under-reported).
# Falls back to the oracle port (kind "port") only if baseline/_ref is missing.
# ---------------------------------------------------------------------------------------------------------------------------
REF_THREADS = 4
_ref_state = {}


def _ref_build_model(path: str):
    """The reference's SynthesizerTrn carrying exactly the weights of the voice file the engine loads."""
    import types
    import warnings
    import torch
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    name = "phoonnx_train.vits.monotonic_align"           # Cython, training-only (models.py:646); its cp310 .so is absent here
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
    from phoonnx_train.vits.models import SynthesizerTrn
    from phoonnx_b200.weights import load_model
    W, a, _ = load_model(path)
    kw = dict(n_vocab=a.n_vocab, spec_channels=513, segment_size=32, inter_channels=a.inter, hidden_channels=a.hidden,
              filter_channels=a.filter, n_heads=a.n_heads, n_layers=a.n_layers, kernel_size=a.enc_kernel, p_dropout=0.1,
              resblock=a.resblock, resblock_kernel_sizes=tuple(a.rb_kernels), resblock_dilation_sizes=tuple(tuple(d) for d in a.rb_dilations),
              upsample_rates=tuple(a.up_rates), upsample_initial_channel=a.up_init, upsample_kernel_sizes=tuple(a.up_kernels),
              n_speakers=a.n_speakers, gin_channels=a.gin, use_sdp=a.use_sdp)                 # lightning.py:86-106
    import contextlib
    with warnings.catch_warnings(), contextlib.redirect_stdout(open(os.devnull, "w")):      # the reference print()s while removing weight norm
        warnings.simplefilter("ignore")
        m = SynthesizerTrn(**kw).eval()
        m.dec.remove_weight_norm()                          # export_onnx.py:242-245
        for f in m.flow.flows:                              # the exporter folds the flows' weight norm into constants
            if hasattr(f, "enc"):
                f.enc.remove_weight_norm()
    sd = m.state_dict()
    missing = []
    with torch.no_grad():
        for k, v in sd.items():
            if k in W:
                v.copy_(torch.from_numpy(np.ascontiguousarray(W[k])).reshape(v.shape))
            elif not (k.startswith("enc_q.") or k.startswith("dp.post_") or k.startswith("dp.flows.1.")):
                missing.append(k)                           # everything infer() touches must come from the file
    if missing:
        raise RuntimeError(f"voice file lacks tensors of the reference model: {missing[:5]}")
    return m, a


def _ref_worker_init(path: str, threads: int):
    import torch
    torch.set_num_threads(threads)
    m, a = _ref_build_model(path)
    _ref_state.update(model=m, arch=a)


def _ref_worker_run(job):
    import torch
    ids, sid = job
    m, a = _ref_state["model"], _ref_state["arch"]
    with torch.no_grad():
        x = torch.from_numpy(np.asarray(ids, np.int64))[None]
        o = m.infer(x, torch.tensor([x.shape[1]]), sid=None if sid is None else torch.tensor([int(sid)]),
                    noise_scale=SCALES[0], length_scale=SCALES[1], noise_scale_w=SCALES[2])[0]
    return int(o.shape[-1])


class ReferencePool:
    def __init__(self, path: str, cores: int):
        import multiprocessing as mp
        self.kind = "reference" if os.path.isdir(os.path.join(REF_DIR, "phoonnx_train", "vits")) else "port"
        self.path, self.cores = path, cores
        self.threads = min(REF_THREADS, cores)
        self.procs = max(1, cores // self.threads)
        self.pool = None
        if self.kind == "reference":
            ctx = mp.get_context("spawn")                   # the parent may already hold a CUDA context
            self.pool = ctx.Pool(self.procs, initializer=_ref_worker_init, initargs=(path, self.threads))
            self.pool.map(_ref_worker_run, [(np.arange(16) % 200, 0 if self._ns() > 1 else None)] * self.procs)     # build + first call

    def _ns(self):
        from phoonnx_b200.weights import load_model
        if not hasattr(self, "_arch"):
            self._arch = load_model(self.path)[1]
        return self._arch.n_speakers

    def run(self, utts, sids):
        """(audio seconds, wall seconds) for these utterances, B=1 per call."""
        arch_sr = self._arch.sample_rate if hasattr(self, "_arch") else None
        if self.kind == "reference":
            jobs = [(u, None if sids is None else int(s)) for u, s in zip(utts, sids if sids is not None else [None] * len(utts))]
            jobs.sort(key=lambda j: -len(j[0]))             # longest first: the pool's tail stays short
            t0 = time.perf_counter()
            samples = self.pool.map(_ref_worker_run, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            self._ns()
            return sum(samples) / self._arch.sample_rate, dt
        # fallback: the oracle port in this process (all threads) -- reported as kind "port"
        import torch
        from oracle.vits_oracle import VitsOracle
        from phoonnx_b200.weights import load_model
        torch.set_num_threads(self.cores)
        W, arch, _ = load_model(self.path)
        orc = VitsOracle(W, arch)
        rs = np.random.RandomState(99)
        noise = [(rs.randn(2, len(u)).astype(np.float32), rs.randn(arch.inter, 24 * len(u)).astype(np.float32)) for u in utts]
        t0 = time.perf_counter()
        n = 0
        for i, u in enumerate(utts):
            n += orc.infer(u, SCALES, None if sids is None else int(sids[i]), noise[i][0], noise[i][1], stages=False)["audio"].shape[0]
        return n / arch.sample_rate, time.perf_counter() - t0

    def describe(self, sample: str):
        if self.kind == "reference":
            how = (f"the reference's own phoonnx_train.vits.models.SynthesizerTrn.infer (baseline/_ref, unmodified; torch-CPU; onnxruntime is "
                   f"not installable offline), B=1 per call like voice.py:265-269, {self.procs} worker processes x {self.threads} torch threads")
        else:
            how = "oracle/vits_oracle.py (torch-CPU port; baseline/_ref missing), B=1 loop, one process"
        return {"cores": self.procs * self.threads if self.kind == "reference" else self.cores, "kind": self.kind, "sample": f"{sample}; {how}"}

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def spot_check(sess, path, utts, sids, picks):
    """Checker leg (outside every timed region): the engine at scales (0, 1, 0) vs the CPU oracle on a few utterances of the
    benched workload -- a fast kernel whose results differ from the reference's is not done."""
    from oracle.vits_oracle import VitsOracle
    from phoonnx_b200 import scheduler
    from phoonnx_b200.weights import load_model
    W, arch, _ = load_model(path)
    orc = VitsOracle(W, arch)
    z = np.asarray((0.0, 1.0, 0.0), np.float32)
    x, lens = scheduler.pad_batch([utts[i] for i in picks])
    feed = {"input": x, "input_lengths": lens, "scales": z}
    if sids is not None:
        feed["sid"] = np.asarray([sids[i] for i in picks], np.int64)
    audio, alen = sess.synthesize_packed(feed)
    audio = np.array(audio)
    off, worst, compared = 0, 1e9, 0
    for j, i in enumerate(picks):
        r = orc.infer(utts[i], z, None if sids is None else int(sids[i]), None, None, stages=False)
        n = int(alen[j])
        if r["audio"].shape[0] == n:                        # (a ceil tie changes the length: skipped, counted below)
            d = audio[off:off + n] - r["audio"]
            worst = min(worst, 10 * np.log10(float((r["audio"].astype(np.float64) ** 2).sum()) / max(float((d.astype(np.float64) ** 2).sum()), 1e-30)))
            compared += 1
        off += n
    return {"utterances": len(picks), "compared": compared, "worst_snr_db": None if compared == 0 else round(float(worst), 1),
            "scales": [0.0, 1.0, 0.0], "checker": "oracle/vits_oracle.py"}


# ---------------------------------------------------------------------------------------------------------------------------
# Strong scaling: host gather into one buffer in original order (shared memory owned by rank 0)
# ---------------------------------------------------------------------------------------------------------------------------
class HostGather:
    """ONE host buffer for the whole job's audio in ORIGINAL utterance order, filled by every rank in parallel.  Layout of the
    shared segment: flags int64[2 * world] | frames int64[n_utts] | audio float32[capacity].  Hand-shake without a collective:
    a rank publishes its utterances' frame counts, then its flag; when every flag shows the step number each rank computes the
    same exclusive scan and copies its utterances to their final places (thread pool: numpy slice copies release the GIL)."""

    def __init__(self, name: str, world: int, rank: int, n_utts: int, capacity_samples: int, create: bool):
        from multiprocessing import shared_memory
        self.world, self.rank, self.n = world, rank, n_utts
        hdr = 8 * (2 * world + n_utts)
        hdr = (hdr + 4095) // 4096 * 4096
        size = hdr + 4 * capacity_samples
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=size if create else 0)
        buf = self.shm.buf
        self.flags = np.frombuffer(buf, np.int64, 2 * world, 0)
        self.frames = np.frombuffer(buf, np.int64, n_utts, 8 * 2 * world)
        self.audio = np.frombuffer(buf, np.float32, capacity_samples, hdr)
        self.capacity = capacity_samples
        if create:
            self.flags[:] = 0
            self.frames[:] = 0
            self.audio[:] = 0.0                             # first touch: the pages exist before anything is timed
        from concurrent.futures import ThreadPoolExecutor
        self.pool = ThreadPoolExecutor(max_workers=max(2, min(16, (os.cpu_count() or 8) // max(1, world))))

    def _wait_all(self, col: int, step: int):
        spins = 0
        while int(self.flags[col * self.world:(col + 1) * self.world].min()) < step:
            spins += 1
            if spins > 200:
                time.sleep(0.0001)
            if spins > 2_000_000:
                raise RuntimeError("host gather: a rank never published its lengths")

    def gather(self, step: int, hop: int, results):
        """results: [(global utterance indices, packed audio, samples per utterance)] of this rank's device batches."""
        for idx, _, alen in results:
            self.frames[idx] = alen // hop
        self.flags[self.rank] = step
        self._wait_all(0, step)
        offs = np.concatenate([[0], np.cumsum(self.frames)]) * hop
        if int(offs[-1]) > self.capacity:
            raise RuntimeError("host gather: result buffer too small")
        jobs = []
        for idx, audio, alen in results:
            src = np.concatenate([[0], np.cumsum(alen)])
            for j, g in enumerate(idx):
                jobs.append((int(offs[g]), audio, int(src[j]), int(alen[j])))
        dst = self.audio

        def cp(chunk):
            for o, a, s, n in chunk:
                dst[o:o + n] = a[s:s + n]

        k = max(1, len(jobs) // (4 * self.pool._max_workers))
        list(self.pool.map(cp, [jobs[i:i + k] for i in range(0, len(jobs), k)]))
        self.flags[self.world + self.rank] = step
        if self.rank == 0:
            self._wait_all(1, step)
        return int(offs[-1])

    def gather_direct(self, step: int, hop: int, sess, bidx, feed):
        """Same result without the host memcpy (one device batch per rank): the text side runs first and yields every utterance's
        frame count; once all ranks have published theirs each rank knows where its utterances belong and the frame side's
        device->host DMAs write them THERE (vits_set_output_offsets; the segment is page-locked in every process)."""
        alen = sess.prepare_feed(feed)
        self.frames[bidx] = alen // hop
        self.flags[self.rank] = step
        self._wait_all(0, step)
        offs = np.concatenate([[0], np.cumsum(self.frames)]) * hop
        if int(offs[-1]) > self.capacity:
            raise RuntimeError("host gather: result buffer too small")
        sess.decode_prepared(feed, out="f32", dest=self.audio, dest_offsets=offs[bidx])      # returns when its DMAs are complete
        self.flags[self.world + self.rank] = step
        if self.rank == 0:
            self._wait_all(1, step)
        return int(offs[-1]), alen

    def page_lock(self, engine):
        engine.host_register(self.audio)
        self._locked_by = engine

    def close(self, unlink: bool):
        if getattr(self, "_locked_by", None) is not None:
            self._locked_by.host_unregister(self.audio)
        self.pool.shutdown()
        del self.flags, self.frames, self.audio
        self.shm.close()
        if unlink:
            try:
                self.shm.unlink()
            except Exception:
                pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C5", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="C5 only: per-GPU work fixed (driver line) or ONE job sharded over the ranks")
    ap.add_argument("--utts", type=int, default=None, help="override the config's utterance count (per GPU in weak mode)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--preset", default=None, help="override the config's voice preset")
    ap.add_argument("--max-ids", type=int, default=262144, help="phoneme ids per device batch")
    ap.add_argument("--max-utts", type=int, default=2048, help="utterances per device batch")
    ap.add_argument("--chunk-frames", type=int, default=None, help="frames per decode chunk (default 262144; 65536 under --scaling strong so "
                    "that a rank's device->host DMAs overlap its kernels chunk by chunk: 8-GPU strong e2e 405k -> 527k audio-s/s, profiles/r02q)")
    ap.add_argument("--cpu-sample", type=int, default=96, help="utterances in the CPU-baseline sample")
    ap.add_argument("--calls", type=int, default=300, help="C1: timed run() calls per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather-memcpy", action="store_true", help="strong scaling: gather through host memcpy even when direct DMA placement is possible")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling section of the default C5 line")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="engine option (A/B experiments), repeatable")
    args = ap.parse_args()

    rank, world, local = dist_env()
    cfg = dict(CONFIGS[args.config])
    if args.preset:
        cfg["preset"] = args.preset
    if args.utts:
        cfg["utts"] = args.utts
    strong = args.scaling == "strong"
    if args.chunk_frames is None:
        args.chunk_frames = 65536 if strong else 262144
    if strong and args.config != "C5":
        raise SystemExit("--scaling strong applies to C5 (the sharded throughput sweep)")
    from phoonnx_b200 import modelgen, scheduler
    tmp = tempfile.mkdtemp(prefix="vits_bench_")
    path = os.path.join(tmp, f"{cfg['preset']}.onnx")
    _, arch = modelgen.make_voice(path, cfg["preset"], n_speakers=cfg["ns"], seed=1234)
    wl_seed = cfg["seed"] + (rank if (args.config == "C5" and not strong) else 0)
    utts, lengths = make_workload(cfg["utts"], seed=wl_seed, n_vocab=arch.n_vocab, lo=cfg["lo"], hi=cfg["hi"])
    sids = (np.arange(cfg["utts"]) % cfg["ns"]).astype(np.int64) if cfg["ns"] > 1 else None
    n_threads = os.cpu_count() or 1
    sample_idx = list(range(0, cfg["utts"], max(1, cfg["utts"] // max(1, args.cpu_sample))))[: args.cpu_sample]
    per = "" if args.config != "C5" else (" in total" if strong else "/GPU")
    workload = (f"{args.config}: VITS {cfg['preset']} (random-init, exporter-format file), "
                f"{(str(cfg['utts']) + per + ' ') if args.config == 'C5' else ''}{cfg['what']}, scales {SCALES}, {arch.sample_rate} Hz")

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        pool = ReferencePool(path, n_threads)
        # a step = the bounded sample of the workload (small configs: the whole workload; C1: 16 calls of its one utterance)
        su = [utts[i] for i in sample_idx] if args.config != "C1" else [utts[0]] * 16
        ss = None if sids is None else [sids[i] for i in sample_idx]
        vals = []
        for s in range(args.warmup + args.steps):
            if s < args.warmup:
                pool.run(su[:max(2, pool.procs)], None if ss is None else ss[:max(2, pool.procs)])
            else:
                vals.append(pool.run(su, ss))
        a_tot = sum(v[0] for v in vals); t_tot = sum(v[1] for v in vals)
        v = a_tot / t_tot
        sample = (f"{len(su)} of the {cfg['utts']} utterances per step" if args.config != "C1" else "16 calls of the single 128-id utterance per step")
        cb = dict(pool.describe(sample), value=v, unit="audio-s/s")
        line = {"impl": "reference", "metric": "audio_seconds_per_second", "value": v, "unit": "audio-s/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
                "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "sample": sample}, "cpu_baseline": cb,
                "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        pool.close()
        return

    # ------------------------------------------------------------------ this engine
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        # one rank per GPU: run (and allocate the page-locked result pool) on the CPUs / NUMA node next to this rank's GPU
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from phoonnx_b200.session import B200Session
    sess = B200Session(path, device=local, precision=args.precision, max_chunk_frames=args.chunk_frames, seed=1000 * rank)
    eng = sess.engine
    for kv in args.opt:
        k, v = kv.split("=", 1)
        eng.set_option(k, float(v))

    def make_feeds(w, r):
        batches = scheduler.plan(lengths, w, r, max_ids=args.max_ids, max_utts=args.max_utts)
        feeds = []
        for bidx in batches:
            x, lens = scheduler.pad_batch([utts[i] for i in bidx])
            f = {"input": x, "input_lengths": lens, "scales": np.asarray(SCALES, np.float32)}
            if sids is not None:
                f["sid"] = sids[bidx]
            feeds.append(f)
        return batches, feeds

    batches, feeds = make_feeds(world, rank) if strong else make_feeds(1, 0)
    rep = int(cfg.get("repeat", 1))
    if rep > 1:
        batches, feeds = batches * rep, feeds * rep
    h2d = sum(int(f["input_lengths"].sum()) * 8 + f["input_lengths"].size * 8 + 12 + (f["input_lengths"].size * 8 if "sid" in f else 0) for f in feeds)

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(fs):
        frames = 0
        for f in fs:
            _, alen = sess.synthesize_packed(f, out="none")
            frames += int(alen.sum()) // arch.hop
        return frames

    def e2e_pass(fs, nsteps):
        """K steps through ONE synthesize_many stream, as a service would run them: host ids in, complete host float32 audio out,
        the device->host DMA of batch k under the kernels of batch k+1; every result is checked before the timed region ends."""
        frames = nbytes = 0
        for audio, alen in sess.synthesize_many(fs * nsteps, out="f32"):
            frames += int(alen.sum()) // arch.hop
            nbytes += audio.nbytes
            head = audio[:1 << 20]              # a NaN tile or an unbounded sample is a failed run
            if not (np.isfinite(head).all() and float(np.abs(head).max()) <= 1.0):
                raise SystemExit("bench.py: synthesised audio is not finite / not tanh-bounded")
        return frames, nbytes

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"

    # ================================================================== C1: per-call latency of B200Session.run
    if args.config == "C1":
        feed = {"input": utts[0][None], "input_lengths": np.asarray([len(utts[0])], np.int64), "scales": np.asarray(SCALES, np.float32)}
        for _ in range(20 + 5 * args.warmup):
            sess.run(None, feed)
        barrier()
        sampler = ClockSampler(local); sampler.start()
        lat, samples_out = [], 0
        l0 = eng.launch_count()
        t_all = time.perf_counter()
        for _ in range(args.steps * args.calls):
            t0 = time.perf_counter()
            out = sess.run(None, feed)[0]                    # host ids -> host float32 [1,1,1,T]; exactly voice.py:374-377
            lat.append(time.perf_counter() - t0)
            samples_out += out.shape[-1]
        t_all = time.perf_counter() - t_all
        launches = eng.launch_count() - l0
        # device-resident rate of the same call (events on the engine's stream)
        eng.timer_start()
        fr = 0
        for _ in range(args.steps * args.calls):
            fr += device_step([feed])
        dev_ms = eng.timer_stop()
        stage = eng.stage_ms()
        clocks = sampler.stop()
        lat_ms = np.asarray(lat) * 1e3
        ncalls = args.steps * args.calls
        audio_s = samples_out / arch.sample_rate
        value = (fr * arch.hop / arch.sample_rate) / (dev_ms * 1e-3)
        dec_tf = 2.0 * arch.dec_mac_per_frame() * fr / (stage["dec"] * 1e-3) / 1e12 if stage["dec"] > 0 else 0.0
        line = {"metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": workload, "precision": args.precision, "calls_per_step": args.calls,
                           "l2_policy": "one utterance: the working set is far below L2 by construction; latency is quoted warm, as a serving loop sees it",
                           "frames_per_call": fr // ncalls},
                "latency_ms": {"p50": float(np.percentile(lat_ms, 50)), "p90": float(np.percentile(lat_ms, 90)), "p99": float(np.percentile(lat_ms, 99)),
                               "mean": float(lat_ms.mean()), "min": float(lat_ms.min()), "calls": ncalls, "launches_per_call": launches / ncalls,
                               "device_ms_per_call": dev_ms / ncalls, "what": "B200Session.run(None, feed): host int64 ids -> host float32 audio, one 128-id utterance"},
                "clocks": clocks,
                "e2e": {"value": audio_s / t_all, "unit": "audio-s/s", "h2d_bytes_per_step": args.calls * (128 * 8 + 8 + 12),
                        "d2h_bytes_per_step": int(samples_out * 4 // args.steps), "note": "serial run() calls, wall clock"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "achieved": dec_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": dec_tf / peak_tf, "traffic": None,
                             "kernel": "HiFi-GAN decoder launches of a single ~450-frame utterance: latency-bound (3 tiles for 148 SMs), quoted for completeness",
                             "dec_ms": stage["dec"], "flow_ms": stage["flow"], "text_ms": stage["text"], "peak_source": peak_src}}
        if not args.no_cpu_baseline and world == 1:
            pool = ReferencePool(path, n_threads)
            a_s, dt = pool.run([utts[0]] * (4 * pool.procs), None)
            line["cpu_baseline"] = dict(pool.describe(f"{4 * pool.procs} calls of the same utterance"), value=a_s / dt, unit="audio-s/s")
            # the latency a reference user sees: ONE process, all threads, one call at a time
            pool.close()
            line["spot_check"] = spot_check(sess, path, utts, sids, [0])
        if rank == 0:
            print(json.dumps(line))
        if use_dist:
            dist.destroy_process_group()
        return

    # ================================================================== C2-C5: throughput
    # settle phase (untimed, before the W warm-up steps): one end-to-end pass and one device pass so that every buffer has its
    # final size, every kernel module is loaded and the page-locked result pool exists before anything is counted
    for _ in sess.synthesize_many(feeds, out="f32"):
        pass
    device_step(feeds)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        device_step(feeds)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = eng.launch_count()
    eng.timer_start()
    frames = 0
    for _ in range(args.steps):
        frames += device_step(feeds)
    dev_ms = eng.timer_stop()
    barrier()
    launches = eng.launch_count() - launches0
    stage = eng.stage_ms()
    kern = [eng.kernel_ms(0), eng.kernel_ms(1)]       # (ms, launches, algorithmic MACs) of the two fused decoder kernels
    clocks = sampler.stop()

    # ---- end to end, three passes, the MEDIAN reported (r01 reported the faster of two)
    gather = None
    if strong:
        shm_name = f"vits_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if use_dist else os.getpid()}"
        # capacity every rank agrees on: the job's frames per step as just measured (durations are redrawn every step: 30 % head-room)
        tot = torch.tensor([float(frames) / args.steps], device="cuda", dtype=torch.float64)
        if use_dist:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        est = int(tot.item() * arch.hop * 1.3) + (1 << 20)
        if rank == 0:
            gather = HostGather(shm_name, world, rank, cfg["utts"], est, create=True)
        barrier()
        if rank != 0:
            gather = HostGather(shm_name, world, rank, cfg["utts"], est, create=False)
        barrier()
        direct = len(feeds) == 1 and not args.gather_memcpy
        if direct:
            gather.page_lock(eng)
        barrier()
    gstep = [0]
    if not strong:
        direct = False

    def e2e_timed(nsteps):
        if not strong:
            return e2e_pass(feeds, nsteps)
        # strong: per step, synthesize this rank's share, then place it in the job-wide host buffer in original order
        frames = nbytes = 0
        for _ in range(nsteps):
            gstep[0] += 1
            if direct:
                total, alen = gather.gather_direct(gstep[0], arch.hop, sess, batches[0], feeds[0])
                frames += int(alen.sum()) // arch.hop
                nbytes += int(alen.sum()) * 4
            else:
                res = []
                for bidx, (audio, alen) in zip(batches, sess.synthesize_many(feeds, out="f32")):
                    res.append((bidx, audio, alen))
                    frames += int(alen.sum()) // arch.hop
                    nbytes += audio.nbytes
                total = gather.gather(gstep[0], arch.hop, res)
            if rank == 0:
                head = gather.audio[:1 << 20]
                tail = gather.audio[max(0, total - (1 << 20)):total]
                if not (np.isfinite(head).all() and np.isfinite(tail).all() and float(np.abs(tail).max()) <= 1.0 and float(np.abs(tail).max()) > 0.0):
                    raise SystemExit("bench.py: gathered audio is not finite / not tanh-bounded / empty")
        return frames, nbytes

    e2e_timed(args.steps)        # untimed: the same K-step stream once, so the page-locked result pool has every block the stream needs
    passes = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        e_frames, d2h = e2e_timed(args.steps)
        torch.cuda.synchronize()
        passes.append((time.perf_counter() - t0, e_frames, d2h))
    barrier()

    audio_s = frames * arch.hop / arch.sample_rate
    e_audio_s = passes[0][1] * arch.hop / arch.sample_rate
    d2h = passes[0][2]
    pass_s = [p[0] for p in passes]
    if use_dist:
        t = torch.tensor([dev_ms] + pass_s, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = torch.tensor([audio_s, e_audio_s, float(frames), float(launches), float(d2h)], device="cuda", dtype=torch.float64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dev_ms, *pass_s = t.tolist()
        audio_s, e_audio_s, frames_all, launches_all, d2h_all = s.tolist()
    else:
        frames_all, launches_all, d2h_all = float(frames), float(launches), float(d2h)
    e2e_s = float(np.median(pass_s))            # each pass: slowest rank; then the median pass

    # ---- the default C5 line also carries the strong-scaling figures of the same box (rank 0 alone = the N=1 time of the ONE job)
    strong_sec = None
    if args.config == "C5" and not strong and not args.no_strong:
        g_utts, g_len = make_workload(cfg["utts"], seed=cfg["seed"], n_vocab=arch.n_vocab, lo=cfg["lo"], hi=cfg["hi"])

        def feeds_for(w, r):
            out = []
            for bidx in scheduler.plan(g_len, w, r, max_ids=args.max_ids, max_utts=args.max_utts):
                x, lens = scheduler.pad_batch([g_utts[i] for i in bidx])
                out.append({"input": x, "input_lengths": lens, "scales": np.asarray(SCALES, np.float32)})
            return out

        def timed(fs):
            for _ in sess.synthesize_many(fs, out="f32"):
                pass
            ts = []
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fr, _ = e2e_pass(fs, 1)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            return float(np.median(ts)), fr

        t1 = fr1 = None
        if use_dist:
            if rank == 0:
                t1, fr1 = timed(feeds_for(1, 0))
            barrier()
        mine = feeds_for(world, rank)
        barrier()
        tn, frn = timed(mine)
        if use_dist:
            tt = torch.tensor([tn], device="cuda", dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ff = torch.tensor([float(frn)], device="cuda", dtype=torch.float64); dist.all_reduce(ff, op=dist.ReduceOp.SUM)
            tn, frn = tt.item(), ff.item()
        else:
            t1, fr1 = tn, frn
        if rank == 0:
            strong_sec = {"what": f"the ONE {cfg['utts']}-utterance list (seed {cfg['seed']}) dealt to {world} rank(s) by scheduler.plan; host ids -> host audio per rank, "
                                  "no job-wide gather in this section (see --scaling strong for the gathered number)",
                          "e2e_audio_s_per_s": frn * arch.hop / arch.sample_rate / tn, "seconds": tn, "seconds_one_gpu_same_box": t1,
                          "efficiency_vs_one_gpu": (t1 / (world * tn)) if t1 else None}

    if rank != 0:
        if gather is not None:
            barrier()
            gather.close(False)
        if use_dist:
            dist.destroy_process_group()
        return
    if gather is not None:
        if use_dist:
            barrier()
        gather.close(True)

    dec_flops = 2.0 * arch.dec_mac_per_frame() * frames            # this rank
    dec_tflops = dec_flops / (stage["dec"] * 1e-3) / 1e12 if stage["dec"] > 0 else 0.0
    value = audio_s / (dev_ms * 1e-3)
    kernels = []
    for (ms, n, mac), nm in zip(kern, ("k_mrf3_tc<32>: last stage = ConvTranspose1d + 3 ResBlock2 + lrelu/conv_post/tanh in one kernel (dominant kernel)",
                                       "k_mrf3_tc<64>: other fused multi-receptive-field stages")):
        if n > 0 and ms > 0:
            tf = 2.0 * mac / (ms * 1e-3) / 1e12
            fpl = frames / n if n else 0.0
            tr = KERNEL_TRAFFIC.get(nm[:13])
            traffic = None if tr is None else {"bytes_per_launch": tr["bytes_per_frame"] * fpl, "algorithmic_bytes_per_launch": tr["algorithmic_bytes_per_frame"] * fpl,
                                               "frames_per_launch": fpl, "source": tr["source"] + " (bytes per frame of the ncu capture x the frames per launch measured here)"}
            kernels.append({"name": nm, "launches": n, "ms_per_launch": ms / n, "achieved": tf, "frac": tf / peak_tf,
                            "algorithmic_flops_per_launch": 2.0 * mac / n, "traffic": traffic})
    line = {
        "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
        "data": "synthetic",
        "config": {"workload": workload, "precision": args.precision,
                   "l2_policy": "inputs larger than L2 (per-step activations >> 126 MB)" if frames // args.steps > 4096 else
                                "per-step activations of this small batch are of the order of L2; W warm-up steps, steps back to back",
                   "frames_per_step_per_gpu": frames // args.steps, "ids_per_step_per_gpu": int(sum(int(f["input_lengths"].sum()) for f in feeds)),
                   "device_batches": len(feeds), "passes_over_the_batch_per_step": rep, "chunk_frames": args.chunk_frames, "x_realtime": value,
                   "device_busy_ms_per_step": (stage["text"] + stage["flow"] + stage["dec"]) / args.steps},
        "clocks": clocks,
        "e2e": {"value": e_audio_s / e2e_s, "unit": "audio-s/s", "h2d_bytes_per_step": h2d * (world if use_dist else 1) if strong else h2d,
                "d2h_bytes_per_step": int(d2h_all if strong else d2h) // args.steps,
                "passes_s": [round(x, 4) for x in pass_s], "note": "three K-step passes, the MEDIAN reported; each pass is its slowest rank"
                + (("; includes the host gather of every rank's audio into one buffer in original utterance order (shared memory, rank 0): "
                    + ("every utterance DMA'd straight to its final place (text side first, lengths exchanged through the shared segment, vits_set_output_offsets)"
                       if direct else "per-rank pinned results copied into place by a host thread pool (several device batches per rank)")) if strong else "")},
        "gpu_launches": int(launches_all),
        "roofline": {"bound": "tensor", "achieved": dec_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": dec_tflops / peak_tf, "traffic": None,
                     "traffic_note": "per-kernel DRAM traffic of the two fused kernels is under roofline.kernels[].traffic; the decoder as a whole is a "
                                     "launch family, see profiles/README.md for the per-launch dram bytes of every member",
                     "kernel": "HiFi-GAN decoder = BASELINE.json's 'decoder % of tensor-core peak': every launch between conv_pre and the "
                               "tanh output (k_conv_tc<...> + k_mrf3_tc<64> + k_mrf3_tc<32>), rank 0, CUDA events on the engine's stream",
                     "algorithmic_flops_per_frame": 2 * arch.dec_mac_per_frame(), "dec_ms": stage["dec"],
                     "flow_ms": stage["flow"], "text_ms": stage["text"], "peak_source": peak_src,
                     "kernels": kernels},
    }
    if strong_sec is not None:
        line["strong_scaling"] = strong_sec
    if strong:
        line["config"]["sharding"] = f"scheduler.plan(lengths, {world}, rank): whole length buckets dealt to ranks by estimated cost"
    if not args.no_cpu_baseline and world == 1:
        pool = ReferencePool(path, n_threads)
        su = [utts[i] for i in sample_idx]
        ss = None if sids is None else [sids[i] for i in sample_idx]
        a_s, dt = pool.run(su, ss)
        line["cpu_baseline"] = dict(pool.describe(f"{len(su)} of the {cfg['utts']} utterances"), value=a_s / dt, unit="audio-s/s")
        pool.close()
        rs = np.random.RandomState(0)
        line["spot_check"] = spot_check(sess, path, utts, sids, sorted(rs.choice(cfg["utts"], size=min(3, cfg["utts"]), replace=False).tolist()))
    print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
