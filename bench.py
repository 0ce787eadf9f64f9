#!/usr/bin/env python
"""Throughput benchmark of the phoneme-ids -> audio hot path (BASELINE.json metric:
audio-seconds synthesised per second at 1/2/4/8 B200 + HiFi-GAN decoder fraction of tensor peak).

  python bench.py --gpus N --steps K --warmup W            # this engine (CUDA, sm_100a)
  python bench.py --impl reference --gpus N ...             # the reference's CPU path (oracle port), host cores

Workload (BASELINE.json configs[4], "VITS medium throughput sweep"): Piper-style VITS medium,
random-init (phoonnx_b200.modelgen, exporter-format file), U utterances per GPU of
randint(64,257) phoneme ids (seed 2 + rank), length-bucketed, scales (0.667, 1.0, 0.8), noise
generated on device.  A "step" is one pass over the rank's U utterances.  Utterances shard across
ranks with no data-path collective: weak scaling, per-GPU work fixed.

One JSON line on stdout (rank 0).  `value` = audio-seconds / second with inputs resident on the
device (CUDA events on the engine's stream, max over ranks); `e2e` = the same through the
session call with HOST buffers (ids in, float32 audio out, copies inside the timed region).
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

# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum, `ncu --set full`, one launch over a 131072-frame chunk of this workload)
# -- constants copied from the committed captures, not measured by this script (a number taken under a profiler is never a
# bench value; these only say whether the kernels re-read more than the algorithm needs)
KERNEL_TRAFFIC = {
    "k_mrf3_tc<32>": {"bytes_per_launch": 2.150258e9 + 0.264484e9, "algorithmic_bytes_per_launch": 262144 * (64 * 64 * 2 + 256 * 4),
                      "frames_per_launch": 262144, "source": "profiles/r01g_ncu_mrf3.txt"},
    "k_mrf3_tc<64>": {"bytes_per_launch": 2.150496e9 + 2.104694e9, "algorithmic_bytes_per_launch": 262144 * (64 * 64 * 2) * 2,
                      "frames_per_launch": 262144, "source": "profiles/r01g_ncu_mrf3.txt"},
}
DEC_TRAFFIC = {"bytes": None, "note": "per-kernel DRAM traffic of the two fused kernels is under roofline.kernels[].traffic; the decoder as a "
                                      "whole is a launch family, see profiles/README.md for the per-launch dram bytes of every member"}


def make_workload(n_utts: int, seed: int, n_vocab: int = 256):
    rs = np.random.RandomState(seed)
    lengths = rs.randint(64, 257, size=(n_utts,)).astype(np.int64)
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
            # CUDA_VISIBLE_DEVICES-aware: NVML enumerates physical GPUs
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
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
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_baseline(path: str, utts, sample_idx, threads: int):
    """Oracle (CPU port of the reference arithmetic) on a bounded sample, all host threads."""
    import torch
    from oracle.vits_oracle import VitsOracle
    from phoonnx_b200.weights import load_model
    torch.set_num_threads(threads)
    W, arch, _ = load_model(path)
    orc = VitsOracle(W, arch)
    rs = np.random.RandomState(99)
    audio_s, t0 = 0.0, time.perf_counter()
    frames = 0
    for i in sample_idx:
        ids = utts[i]
        T = len(ids)
        nd = rs.randn(2, T).astype(np.float32)
        nz = rs.randn(arch.inter, 24 * T).astype(np.float32)
        r = orc.infer(ids, SCALES, None, nd, nz, stages=False)
        audio_s += r["audio"].shape[0] / arch.sample_rate
        frames += int(r["y_len"])
    dt = time.perf_counter() - t0
    return audio_s, dt, frames


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--utts", type=int, default=4096, help="utterances per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--preset", default="medium")
    ap.add_argument("--max-ids", type=int, default=262144, help="phoneme ids per device batch")
    ap.add_argument("--max-utts", type=int, default=2048, help="utterances per device batch")
    ap.add_argument("--chunk-frames", type=int, default=262144)
    ap.add_argument("--cpu-sample", type=int, default=24, help="utterances in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="engine option (A/B experiments), repeatable")
    args = ap.parse_args()

    rank, world, local = dist_env()
    from phoonnx_b200 import modelgen, scheduler
    tmp = tempfile.mkdtemp(prefix="vits_bench_")
    path = os.path.join(tmp, f"{args.preset}.onnx")
    _, arch = modelgen.make_voice(path, args.preset, n_speakers=1, seed=1234)
    utts, lengths = make_workload(args.utts, seed=2 + rank, n_vocab=arch.n_vocab)
    n_threads = os.cpu_count() or 1
    sample_idx = list(range(0, args.utts, max(1, args.utts // max(1, args.cpu_sample))))[: args.cpu_sample]
    workload = (f"C5: VITS {args.preset} (random-init, exporter-format file), {args.utts} utterances/GPU of "
                f"randint(64,257) phoneme ids, length-bucketed, scales {SCALES}, {arch.sample_rate} Hz")

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        vals = []
        for s in range(args.warmup + args.steps):
            a_s, dt, _ = cpu_baseline(path, utts, sample_idx if s >= args.warmup else sample_idx[:2], n_threads)
            if s >= args.warmup:
                vals.append((a_s, dt))
        a_tot = sum(v[0] for v in vals); t_tot = sum(v[1] for v in vals)
        v = a_tot / t_tot
        line = {"impl": "reference", "metric": "audio_seconds_per_second", "value": v, "unit": "audio-s/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "sample": f"{len(sample_idx)} utterances (every {args.utts // max(1, len(sample_idx))}th) per step"},
                "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": n_threads, "kind": "port",
                                 "sample": f"{len(sample_idx)} of the {args.utts} utterances per step, B=1 loop like voice.py:265-269; "
                                           "oracle/vits_oracle.py (torch-CPU restatement of SynthesizerTrn.infer; onnxruntime is not installable offline)"},
                "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
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
        # (8-GPU check of round 1: device-resident rate scaled linearly, the end-to-end rate lost 15 % to the host side)
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
    batches = scheduler.plan(lengths, 1, 0, max_ids=args.max_ids, max_utts=args.max_utts)
    feeds = []
    for bidx in batches:
        x, lens = scheduler.pad_batch([utts[i] for i in bidx])
        feeds.append({"input": x, "input_lengths": lens, "scales": np.asarray(SCALES, np.float32)})
    h2d = sum(int(f["input_lengths"].sum()) * 8 + f["input_lengths"].size * 8 + 12 for f in feeds)

    def one_step(out_kind: str):
        frames = 0
        nbytes = 0
        for f in feeds:
            audio, alen = sess.synthesize_packed(f, out=out_kind)
            frames += int(alen.sum()) // arch.hop
            if audio is not None:
                nbytes += audio.nbytes
        return frames, nbytes

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # settle phase (untimed, before the W warm-up steps): one end-to-end pass and one device pass so that every buffer has its
    # final size, every kernel module is loaded and the page-locked result pool exists before anything is counted -- the first
    # CUDA process on a fresh box otherwise showed host-side gaps between launches well into the timed steps (r01 2-GPU runs)
    for _ in sess.synthesize_many(feeds, out="f32"):
        pass
    one_step("none")
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        one_step("none")
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = eng.launch_count()
    eng.timer_start()
    frames = 0
    for _ in range(args.steps):
        fr, _ = one_step("none")
        frames += fr
    dev_ms = eng.timer_stop()
    barrier()
    launches = eng.launch_count() - launches0
    stage = eng.stage_ms()
    kern = [eng.kernel_ms(0), eng.kernel_ms(1)]       # (ms, launches, algorithmic MACs) of the two fused decoder kernels
    clocks = sampler.stop()
    # end-to-end: host ids -> host float32 audio through the session's batch call (B200Session.synthesize_many: the
    # device->host transfer of batch k overlaps the kernels of batch k+1; every result is complete when it is yielded)
    def e2e_step(nsteps=1):
        # the K steps go through ONE synthesize_many stream, as a service would run them: the device->host DMA of the last batch of
        # step s runs under the kernels of step s + 1 (every result is complete and checked before the timed region ends)
        frames = nbytes = 0
        for audio, alen in sess.synthesize_many(feeds * nsteps, out="f32"):
            frames += int(alen.sum()) // arch.hop
            nbytes += audio.nbytes
            head = audio[:1 << 20]              # cheap sanity check of every result: a NaN tile or an unbounded sample is a failed run
            if not (np.isfinite(head).all() and float(np.abs(head).max()) <= 1.0):
                raise SystemExit("bench.py: synthesised audio is not finite / not tanh-bounded")
        return frames, nbytes

    e2e_step(args.steps)        # untimed: the same K-step stream once, so the page-locked result pool has every block the stream needs
    barrier()
    t0 = time.perf_counter()
    e_frames, d2h = e2e_step(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    # the end-to-end number is wall-clock on the host: a second K-step pass guards it against a one-off host stall (seen on the first
    # CUDA process of a fresh box); the faster pass is reported, both are in the JSON line
    t0 = time.perf_counter()
    e_frames2, d2h2 = e2e_step(args.steps)
    torch.cuda.synchronize()
    e2e_s2 = time.perf_counter() - t0
    barrier()
    e2e_passes = [e2e_s, e2e_s2]
    if e2e_s2 < e2e_s:
        e2e_s, e_frames, d2h = e2e_s2, e_frames2, d2h2

    audio_s = frames * arch.hop / arch.sample_rate
    e_audio_s = e_frames * arch.hop / arch.sample_rate
    if use_dist:
        t = torch.tensor([dev_ms, e2e_s] + e2e_passes, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = torch.tensor([audio_s, e_audio_s, float(frames), float(launches), stage["dec"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s, p1, p2 = t.tolist()
        e2e_passes = [p1, p2]
        e2e_s = min(p1, p2)                     # the faster pass by its slowest rank
        audio_s, e_audio_s, frames_all, launches_all, dec_ms_sum = s.tolist()
    else:
        frames_all, launches_all, dec_ms_sum = float(frames), float(launches), stage["dec"]
    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    dec_flops = 2.0 * arch.dec_mac_per_frame() * frames            # this rank
    dec_tflops = dec_flops / (stage["dec"] * 1e-3) / 1e12 if stage["dec"] > 0 else 0.0
    value = audio_s / (dev_ms * 1e-3)
    e2e_passes_max = [round(x, 4) for x in e2e_passes]
    # the two fused decoder kernels on their own (events around each launch; algorithmic FLOPs = no halo, no padding)
    kernels = []
    for (ms, n, mac), nm in zip(kern, ("k_mrf3_tc<32>: last stage = ConvTranspose1d + 3 ResBlock2 + lrelu/conv_post/tanh in one kernel (dominant kernel)",
                                       "k_mrf3_tc<64>: other fused multi-receptive-field stages")):
        if n > 0 and ms > 0:
            tf = 2.0 * mac / (ms * 1e-3) / 1e12
            kernels.append({"name": nm, "launches": n, "ms_per_launch": ms / n, "achieved": tf, "frac": tf / peak_tf,
                            "algorithmic_flops_per_launch": 2.0 * mac / n,
                            "traffic": KERNEL_TRAFFIC.get(nm[:13])})
    line = {
        "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
        "data": "synthetic",
        "config": {"workload": workload, "precision": args.precision, "l2_policy": "inputs larger than L2 (per-step activations >> 126 MB)",
                   "frames_per_step_per_gpu": frames // args.steps, "ids_per_step_per_gpu": int(lengths.sum()),
                   "device_batches": len(feeds), "chunk_frames": args.chunk_frames, "x_realtime": value,
                   "device_busy_ms_per_step": (stage["text"] + stage["flow"] + stage["dec"]) / args.steps},
        "clocks": clocks,
        "e2e": {"value": e_audio_s / e2e_s, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h // args.steps,
                "passes_s": e2e_passes_max, "note": "two K-step passes, the faster one reported; max over ranks each"},
        "gpu_launches": int(launches_all),
        "roofline": {"bound": "tensor", "achieved": dec_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": dec_tflops / peak_tf, "traffic": DEC_TRAFFIC["bytes"],
                     "traffic_note": DEC_TRAFFIC["note"],
                     "kernel": "HiFi-GAN decoder = BASELINE.json's 'decoder % of tensor-core peak': every launch between conv_pre and the "
                               "tanh output (k_conv_tc<...> + k_mrf3_tc<64> + k_mrf3_tc<32>), rank 0, CUDA events on the engine's stream",
                     "algorithmic_flops_per_frame": 2 * arch.dec_mac_per_frame(), "dec_ms": stage["dec"],
                     "flow_ms": stage["flow"], "text_ms": stage["text"], "peak_source": peak_src,
                     "kernels": kernels},
    }
    if not args.no_cpu_baseline and world == 1:
        a_s, dt, _ = cpu_baseline(path, utts, sample_idx, n_threads)
        line["cpu_baseline"] = {"value": a_s / dt, "unit": "audio-s/s", "cores": n_threads, "kind": "port",
                                "sample": f"{len(sample_idx)} of the {args.utts} utterances, B=1 loop (voice.py:265-269), "
                                          "oracle/vits_oracle.py on torch-CPU (onnxruntime absent offline)"}
    print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
