"""Segment-level data parallelism: the windows the reference cuts long audio into, and how they
are spread over GPUs.  Segments are independent until the host-side text merge, so there is no
collective on the path."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

SR = 16000


def segment_windows(n_samples: int, segment_s: float = 60.0, overlap_s: float = 4.0, sr: int = SR) -> List[Tuple[int, int]]:
    """[start, end) sample windows exactly as core/orchestrator.py:64-67,128-136,143 cuts them:
    one pass if duration <= segment+2 s, else windows of `segment_s` stepping by segment-overlap."""
    duration = n_samples / sr
    if duration <= segment_s + 2.0:
        return [(0, n_samples)]
    out, step, curr = [], segment_s - overlap_s, 0.0
    while curr < duration:
        end = min(curr + segment_s, duration)
        out.append((int(curr * sr), int(end * sr)))
        if end >= duration:
            break
        curr += step
    return out


def shard(n_items: int, world: int, rank: int) -> List[int]:
    """Round-robin ownership of items by rank: no rank holds more than ceil(n/world)."""
    return list(range(rank, n_items, world))


def pack_batches(lengths: List[int], max_batch: int) -> List[List[int]]:
    """Group segment indices into batches of similar length (longest first) to limit padding."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    return [order[i:i + max_batch] for i in range(0, len(order), max_batch)]


def pad_batch(audio: np.ndarray, windows: List[Tuple[int, int]], idx: List[int]):
    """-> (batch [B][S_max] zero-padded, ilens [B])."""
    lens = [windows[i][1] - windows[i][0] for i in idx]
    out = np.zeros((len(idx), max(lens)), np.float32)
    for r, i in enumerate(idx):
        out[r, :lens[r]] = audio[windows[i][0]:windows[i][1]]
    return out, lens
