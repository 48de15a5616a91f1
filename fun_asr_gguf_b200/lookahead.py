"""Look-ahead orchestration: run the front half of EVERY segment of a long file as a few
batches before the reference's per-segment loop starts, so that `transcribe()` itself gets batch throughput.

The reference cuts long audio into windows (core/orchestrator.py:128-136) and, per window, calls
`encode_audio` (nano_onnx.py:78-133) and then the CTC session (core/decoder.py:27), interleaved with LLM
decoding.  Nothing in those calls depends on the previous window, so all of them can be computed up front:

    from fun_asr_gguf_b200 import ort_shim, lookahead
    ort_shim.install()
    ...
    lookahead.prefetch(audio)          # one extra line before _transcribe_long's loop
    # the unchanged per-segment encoder / CTC session calls now return the precomputed arrays

Physical lengths are the ones `encode_audio` would feed (the native window length, or 1 s for a shorter
one — the CPU-provider rule at nano_onnx.py:90-99).  The unmasked CTC head's ids depend on the physical length
(SURVEY F7), so padding a short window up to a longer one would change them; the engine's ragged call
(`fa_front_half_ragged`) lets every window keep its own physical length inside a shared batch, so ALL windows of a
file — the short tail included — go through together (an engine without it, the fp32 arbiter mode, falls back to
batches of equal physical length).  Results are therefore identical to the per-segment calls.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import segments as S
from . import weights as W


@dataclass
class SegmentFront:
    window: Tuple[int, int]          # [start, end) samples in the file
    n_valid: int                     # true samples
    n_phys: int                      # samples fed (>= n_valid)
    enc_output: np.ndarray           # (1, T, 512)   what encoder_sess returns first
    adaptor_output: np.ndarray       # (1, T, 1024)  rows >= target_len zero
    ids: np.ndarray                  # (1, T) int32  what ctc_sess returns
    target_len: int

    @property
    def audio_embd(self) -> np.ndarray:          # nano_onnx.py:124-131
        return self.adaptor_output[0, :self.target_len, :]


def physical_samples(n_valid: int) -> int:
    """Samples `encode_audio` feeds for a window of n_valid samples when the provider is the CPU one."""
    return max(n_valid, W.SAMPLE_RATE)


def _key(audio_1d: np.ndarray, n_valid: int) -> str:
    h = hashlib.blake2b(digest_size=16)
    h.update(np.int64(n_valid).tobytes())
    h.update(np.ascontiguousarray(audio_1d[:n_valid], dtype=np.float32).tobytes())
    return h.hexdigest()


class FrontCache:
    """Results of `run_file`, looked up by the content of a segment (valid samples + count)."""

    def __init__(self, capacity: int = 512):
        self.capacity = capacity
        self._by_audio: "OrderedDict[str, SegmentFront]" = OrderedDict()
        self._by_enc: Dict[int, SegmentFront] = {}
        self.hits = 0

    def put(self, audio_1d: np.ndarray, seg: SegmentFront) -> None:
        self._by_audio[_key(audio_1d, seg.n_valid)] = seg
        self._by_enc[id(seg.enc_output)] = seg
        while len(self._by_audio) > self.capacity:
            _, old = self._by_audio.popitem(last=False)
            self._by_enc.pop(id(old.enc_output), None)

    def encoder_lookup(self, audio_phys: np.ndarray, n_valid: int) -> Optional[SegmentFront]:
        if not self._by_audio:               # nothing was prefetched: do not hash 3.8 MB per call for a certain miss
            return None
        seg = self._by_audio.get(_key(audio_phys, n_valid))
        if seg is not None and seg.n_phys == audio_phys.shape[0]:
            self.hits += 1
            return seg
        return None

    def ctc_lookup(self, enc: np.ndarray) -> Optional[SegmentFront]:
        """The CTC session is fed the very array the encoder session returned (core/decoder.py:27)."""
        seg = self._by_enc.get(id(enc)) if self._by_enc else None
        if seg is not None and seg.enc_output is enc:
            self.hits += 1
            return seg
        return None

    def clear(self) -> None:
        self._by_audio.clear()
        self._by_enc.clear()


def run_file(engine, audio: np.ndarray, segment_s: float = 60.0, overlap_s: float = 4.0, world: int = 1, rank: int = 0,
             cache: Optional[FrontCache] = None) -> List[Optional[SegmentFront]]:
    """Front half of every window of `audio` owned by `rank` (round-robin over `world`), in batches of windows
    that share a physical length.  Returns one entry per window (None for windows other ranks own)."""
    audio = np.ascontiguousarray(audio, dtype=np.float32).reshape(-1)
    windows = S.segment_windows(audio.shape[0], segment_s, overlap_s)
    mine = S.shard(len(windows), world, rank)
    out: List[Optional[SegmentFront]] = [None] * len(windows)
    def emit(group, batch, lens, n_phys, enc, ad, ids):
        for r, i in enumerate(group):
            t = W.lfr_frames(n_phys[r])
            seg = SegmentFront(window=windows[i], n_valid=lens[r], n_phys=n_phys[r],
                               enc_output=np.ascontiguousarray(enc[r:r + 1, :t]), adaptor_output=np.ascontiguousarray(ad[r:r + 1, :t]),
                               ids=np.ascontiguousarray(ids[r:r + 1, :t]), target_len=W.adaptor_target_len(lens[r]))
            out[i] = seg
            if cache is not None:
                cache.put(batch[r, :n_phys[r]], seg)

    phys_of = {i: physical_samples(windows[i][1] - windows[i][0]) for i in mine}
    if getattr(engine, "supports_ragged", False):
        # every window keeps ITS OWN physical length inside a shared batch (fa_front_half_ragged), so windows of any
        # lengths — the short tail of a file included — go through together, in evenly sized batches
        order = sorted(mine, key=lambda i: -phys_of[i])
        n_batches = -(-len(order) // engine.max_batch) if order else 0
        for k in range(n_batches):
            group = order[k::n_batches]
            n_phys = [phys_of[i] for i in group]
            batch = np.zeros((len(group), max(n_phys)), np.float32)
            lens = []
            for r, i in enumerate(group):
                a, b = windows[i]
                batch[r, :b - a] = audio[a:b]
                lens.append(b - a)
            enc, ad, ids = engine.front_half(batch, lens, phys=n_phys)
            emit(group, batch, lens, n_phys, enc, ad, ids)
        return out
    by_phys: Dict[int, List[int]] = {}
    for i in mine:
        by_phys.setdefault(phys_of[i], []).append(i)
    for n_phys, idx in sorted(by_phys.items(), reverse=True):
        for b0 in range(0, len(idx), engine.max_batch):
            group = idx[b0:b0 + engine.max_batch]
            batch = np.zeros((len(group), n_phys), np.float32)
            lens = []
            for r, i in enumerate(group):
                a, b = windows[i]
                batch[r, :b - a] = audio[a:b]
                lens.append(b - a)
            enc, ad, ids = engine.front_half(batch, lens)
            emit(group, batch, lens, [n_phys] * len(group), enc, ad, ids)
    return out


_cache = FrontCache()


def cache() -> FrontCache:
    return _cache


def prefetch(audio: np.ndarray, model_path: str = "Fun-ASR-Nano-Encoder-Adaptor.onnx", segment_s: float = 60.0,
             overlap_s: float = 4.0) -> int:
    """Compute every window of `audio` on the engine the shim's sessions use and park the results where the
    sessions look first.  Returns the number of windows computed."""
    from . import ort_shim
    n = audio.reshape(-1).shape[0]
    longest = max(physical_samples(b - a) for a, b in S.segment_windows(n, segment_s, overlap_s))
    eng = ort_shim._engine_for(model_path, min_samples=longest)
    return sum(s is not None for s in run_file(eng, audio, segment_s, overlap_s, cache=_cache))
