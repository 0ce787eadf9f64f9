"""Utterance-level batching and sharding.

The path shards by utterance: no cross-utterance dependence exists anywhere in
``SynthesizerTrn.infer`` (masks isolate batch rows, models.py:705-709), so the multi-GPU
strategy is replicas + host-side partitioning, with NO collective on the data path:

  * ``bucket_by_length``   sort by phoneme count, cut into buckets of similar length (the engine
                           is varlen, so this only improves tile occupancy, never correctness);
  * ``shard_buckets``      deal whole buckets to ranks greedily by estimated cost so that the
                           expected frame count (the decoder dominates) balances;
  * ``device_batches``     split one rank's work into device batches bounded by total ids.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def bucket_by_length(lengths: Sequence[int], width: int = 16) -> List[np.ndarray]:
    """Indices grouped into buckets whose lengths fall in the same ``width``-wide bin, longest first."""
    lengths = np.asarray(lengths, dtype=np.int64)
    if lengths.size == 0:
        return []
    order = np.argsort(-lengths, kind="stable")
    bins = (lengths[order] - 1) // int(width)
    out: List[np.ndarray] = []
    start = 0
    for i in range(1, len(order) + 1):
        if i == len(order) or bins[i] != bins[start]:
            out.append(order[start:i])
            start = i
    return out


def estimate_cost(lengths: np.ndarray, frames_per_id: float = 3.5, a: float = 1.0, b: float = 4.0) -> np.ndarray:
    """cost ~ a*T (text side) + b*T*frames_per_id (frame side; the decoder is ~90% of the FLOPs)."""
    lengths = np.asarray(lengths, dtype=np.float64)
    return a * lengths + b * lengths * frames_per_id


def shard_buckets(buckets: List[np.ndarray], lengths: Sequence[int], world_size: int,
                  max_bucket: int = 64) -> List[List[np.ndarray]]:
    """Greedy longest-processing-time assignment of (sub-)buckets to ranks.  Returns, per rank, its list
    of index arrays.  Buckets larger than ``max_bucket`` are split so the deal is fine-grained enough
    to balance; every utterance is assigned to exactly one rank."""
    lengths = np.asarray(lengths, dtype=np.int64)
    pieces: List[np.ndarray] = []
    for bk in buckets:
        for s in range(0, len(bk), max_bucket):
            pieces.append(bk[s:s + max_bucket])
    costs = [float(estimate_cost(lengths[p]).sum()) for p in pieces]
    order = np.argsort(-np.asarray(costs), kind="stable")
    load = np.zeros((world_size,), np.float64)
    out: List[List[np.ndarray]] = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        out[r].append(pieces[i])
        load[r] += costs[i]
    return out


def device_batches(pieces: List[np.ndarray], lengths: Sequence[int], max_ids: int = 32768,
                   max_utts: int = 512) -> List[np.ndarray]:
    """Merge a rank's pieces (already length-sorted within themselves) into device batches."""
    lengths = np.asarray(lengths, dtype=np.int64)
    flat = np.concatenate(pieces) if pieces else np.zeros((0,), np.int64)
    if flat.size == 0:
        return []
    flat = flat[np.argsort(-lengths[flat], kind="stable")]
    out, cur, tot = [], [], 0
    for i in flat:
        L = int(lengths[i])
        if cur and (tot + L > max_ids or len(cur) >= max_utts):
            out.append(np.asarray(cur, np.int64))
            cur, tot = [], 0
        cur.append(int(i))
        tot += L
    if cur:
        out.append(np.asarray(cur, np.int64))
    return out


def plan(lengths: Sequence[int], world_size: int = 1, rank: int = 0, bucket_width: int = 16,
         max_ids: int = 32768, max_utts: int = 512) -> List[np.ndarray]:
    """Device batches (index arrays into the utterance list) for ``rank`` of ``world_size``."""
    buckets = bucket_by_length(lengths, bucket_width)
    shards = shard_buckets(buckets, lengths, world_size)
    return device_batches(shards[rank], lengths, max_ids, max_utts)


def pad_batch(utterances: List[np.ndarray]):
    """List of id arrays -> (input [B, Tmax] int64 zero-padded, input_lengths [B]) -- the feed of voice.py:350-355."""
    B = len(utterances)
    lens = np.asarray([len(u) for u in utterances], np.int64)
    x = np.zeros((B, int(lens.max())), np.int64)
    for b, u in enumerate(utterances):
        x[b, : len(u)] = u
    return x, lens
