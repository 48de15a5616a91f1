"""FrontHalf — the host-side object over the C ABI: one context = weights + workspace on one B200.

It exposes the two graph executions of the reference's sessions (encoder+adaptor, CTC head) for
batches of independent segments, with numpy (host) and torch.cuda (device-resident) variants.
PyTorch is used only to own device memory / streams for the device variants.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _lib, weights as W


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class FrontHalf:
    def __init__(self, tensors: Optional[Dict[str, "object"]] = None, device: int = 0, max_batch: int = 32,
                 max_samples: int = 62 * W.SAMPLE_RATE, precision: str = "bf16x3", seed: int = 0):
        """tensors: reference state_dict keys -> fp32 arrays/tensors; None => seeded random init."""
        self.lib = _lib.load()
        if precision not in _lib.PREC:
            raise ValueError(f"precision must be one of {sorted(_lib.PREC)}")
        self.device, self.max_batch, self.max_samples, self.precision = device, int(max_batch), int(max_samples), precision
        if tensors is None:
            tensors = W.random_weights(seed)
        handle = C.c_void_p()
        _lib.check(self.lib.fa_ctx_create(device, self.max_batch, self.max_samples, _lib.PREC[precision], C.byref(handle)))
        self._h = handle
        try:
            consts = W.front_end_constants(W.lfr_frames(self.max_samples))
            for name, t in list(tensors.items()) + list(consts.items()):
                arr = np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float32)
                shape = (C.c_int64 * arr.ndim)(*arr.shape)
                _lib.check(self.lib.fa_ctx_load_tensor(self._h, name.encode(), _ptr(arr), shape, arr.ndim))
            _lib.check(self.lib.fa_ctx_finalize(self._h))
        except Exception:
            self.close()
            raise
        v = C.c_int()
        _lib.check(self.lib.fa_ctx_vocab(self._h, C.byref(v)))
        self.vocab = v.value
        self.blank_id = self.vocab - 1

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            self.lib.fa_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def supports_ragged(self) -> bool:
        """Per-segment physical lengths (front_half(..., phys=...)) need the padding-free path: a tensor-core mode."""
        return self.precision != "fp32"

    # ------------------------------------------------------------------ shapes
    @staticmethod
    def frames(samples: int) -> int:
        return W.lfr_frames(samples)

    @staticmethod
    def target_len(n_valid: int) -> int:
        return W.adaptor_target_len(n_valid)

    def _ilens(self, ilens: Sequence[int], batch: int):
        arr = (C.c_int64 * batch)(*[int(v) for v in ilens])
        if len(ilens) != batch:
            raise ValueError("ilens must have one entry per segment")
        return arr

    # ------------------------------------------------------------------ host (numpy) API
    def encode(self, audio: np.ndarray, ilens: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
        """audio [B][S] fp32 (rows zero-padded to S) -> enc [B][T][512], adaptor_output [B][T][1024]."""
        audio = np.ascontiguousarray(audio, dtype=np.float32)
        if audio.ndim != 2:
            raise ValueError("audio must be [batch][samples]")
        b, s = audio.shape
        t = self.frames(s)
        enc = np.empty((b, t, W.D_ENC), np.float32)
        ad = np.empty((b, t, W.D_LLM), np.float32)
        _lib.check(self.lib.fa_encode(self._h, _ptr(audio), b, s, self._ilens(ilens, b), _ptr(enc), _ptr(ad)))
        return enc, ad

    def ctc(self, enc: np.ndarray) -> np.ndarray:
        """enc [B][T][512] -> greedy ids [B][T] int32 over every physical frame."""
        enc = np.ascontiguousarray(enc, dtype=np.float32)
        if enc.ndim != 3 or enc.shape[2] != W.D_ENC:
            raise ValueError("enc must be [batch][frames][512]")
        b, t, _ = enc.shape
        ids = np.empty((b, t), np.int32)
        _lib.check(self.lib.fa_ctc(self._h, _ptr(enc), b, t, _ptr(ids)))
        return ids

    def front_half(self, audio: np.ndarray, ilens: Sequence[int], want_enc: bool = True, want_adaptor: bool = True,
                   phys: Optional[Sequence[int]] = None):
        """Both graphs back to back; enc never leaves the device between them.

        phys (fa_front_half_ragged): per-segment PHYSICAL sample counts, ilens[b] <= phys[b] <= audio.shape[1].  Segment b
        is then computed as the reference computes it when fed at physical length phys[b], whatever else is in the batch;
        ids[b, frames(phys[b]):] are -1."""
        audio = np.ascontiguousarray(audio, dtype=np.float32)
        b, s = audio.shape
        t = self.frames(s)
        enc = np.empty((b, t, W.D_ENC), np.float32) if want_enc else None
        ad = np.empty((b, t, W.D_LLM), np.float32) if want_adaptor else None
        ids = np.empty((b, t), np.int32)
        if phys is None:
            _lib.check(self.lib.fa_front_half(self._h, _ptr(audio), b, s, self._ilens(ilens, b),
                                              _ptr(enc) if want_enc else None, _ptr(ad) if want_adaptor else None, _ptr(ids)))
        else:
            _lib.check(self.lib.fa_front_half_ragged(self._h, _ptr(audio), b, s, self._ilens(ilens, b), self._ilens(phys, b),
                                                     _ptr(enc) if want_enc else None, _ptr(ad) if want_adaptor else None, _ptr(ids)))
        return enc, ad, ids

    def front_half_into(self, audio: np.ndarray, ilens: Sequence[int], embd: np.ndarray, row_offset: int = 0,
                        want_enc: bool = False):
        """Embedding handoff.  Both graphs back to back; of each segment's adaptor_output only the rows
        the LLM reads, [0, target_len), leave the device, written straight into ``embd`` — a C-contiguous float32
        [rows][1024] numpy array such as a view of ``llama_batch.embd``, or a torch tensor in host or device memory — one segment after another from ``row_offset``
        (core/decoder.py:199 concatenates prefix, audio and suffix embeddings and llama.py:547 memmoves the result; here
        the audio part lands in place).  Returns (rows per segment, ids [B][T], enc [B][T][512] or None)."""
        audio = np.ascontiguousarray(audio, dtype=np.float32)
        b, s = audio.shape
        if isinstance(embd, np.ndarray):
            ok = (embd.dtype == np.float32 and embd.ndim == 2 and embd.shape[1] == W.D_LLM and embd.flags["C_CONTIGUOUS"]
                  and embd.flags["WRITEABLE"])
            base = embd.ctypes.data if ok else 0
        else:                                   # a torch tensor, host or CUDA (the copy is cudaMemcpyDefault)
            import torch
            ok = (isinstance(embd, torch.Tensor) and embd.dtype == torch.float32 and embd.dim() == 2
                  and embd.shape[1] == W.D_LLM and embd.is_contiguous())
            base = embd.data_ptr() if ok else 0
        if not ok:
            raise ValueError("embd must be a writable C-contiguous float32 array or tensor of shape [rows][1024]")
        rows = [int(self.lib.fa_adaptor_rows_for_samples(int(n))) for n in ilens]
        if len(rows) != b:
            raise ValueError("ilens must have one entry per segment")
        if row_offset < 0 or row_offset + sum(rows) > embd.shape[0]:
            raise ValueError(f"embd has {embd.shape[0]} rows; {row_offset} + {sum(rows)} are needed")
        t = self.frames(s)
        enc = np.empty((b, t, W.D_ENC), np.float32) if want_enc else None
        ids = np.empty((b, t), np.int32)
        dst = (C.c_void_p * b)()
        off = row_offset
        for i, r in enumerate(rows):
            dst[i] = base + off * W.D_LLM * 4
            off += r
        got = (C.c_int64 * b)()
        _lib.check(self.lib.fa_front_half_embd(self._h, _ptr(audio), b, s, self._ilens(ilens, b),
                                               _ptr(enc) if want_enc else None, dst, got, _ptr(ids)))
        assert list(got) == rows
        return rows, ids, enc

    # ------------------------------------------------------------------ device (torch.cuda) API
    def use_torch_stream(self):
        import torch
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fa_ctx_set_stream(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def encode_cuda(self, audio, ilens: Sequence[int], enc=None, adaptor=None):
        """audio: torch.float32 CUDA tensor [B][S] on this device; asynchronous on the context's stream."""
        import torch
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous() and audio.dim() == 2
        b, s = audio.shape
        t = self.frames(s)
        if enc is None:
            enc = torch.empty((b, t, W.D_ENC), dtype=torch.float32, device=audio.device)
        if adaptor is None:
            adaptor = torch.empty((b, t, W.D_LLM), dtype=torch.float32, device=audio.device)
        _lib.check(self.lib.fa_encode_dev(self._h, C.c_void_p(audio.data_ptr()), b, s, self._ilens(ilens, b),
                                          C.c_void_p(enc.data_ptr()), C.c_void_p(adaptor.data_ptr())))
        return enc, adaptor

    def ctc_cuda(self, enc, ids=None):
        import torch
        assert enc.is_cuda and enc.dtype == torch.float32 and enc.is_contiguous() and enc.dim() == 3
        b, t, _ = enc.shape
        if ids is None:
            ids = torch.empty((b, t), dtype=torch.int32, device=enc.device)
        _lib.check(self.lib.fa_ctc_dev(self._h, C.c_void_p(enc.data_ptr()), b, t, C.c_void_p(ids.data_ptr())))
        return ids

    def front_half_cuda(self, audio, ilens: Sequence[int], enc=None, adaptor=None, ids=None, phys: Optional[Sequence[int]] = None):
        """Both graphs back to back on CUDA tensors (fa_front_half_dev); asynchronous on the context's stream.  Mixed-length
        batches stay padding-free through the CTC head as well.  Returns (enc, adaptor, ids)."""
        import torch
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous() and audio.dim() == 2
        b, s = audio.shape
        t = self.frames(s)
        if enc is None:
            enc = torch.empty((b, t, W.D_ENC), dtype=torch.float32, device=audio.device)
        if adaptor is None:
            adaptor = torch.empty((b, t, W.D_LLM), dtype=torch.float32, device=audio.device)
        if ids is None:
            ids = torch.empty((b, t), dtype=torch.int32, device=audio.device)
        if phys is None:
            _lib.check(self.lib.fa_front_half_dev(self._h, C.c_void_p(audio.data_ptr()), b, s, self._ilens(ilens, b),
                                                  C.c_void_p(enc.data_ptr()), C.c_void_p(adaptor.data_ptr()), C.c_void_p(ids.data_ptr())))
        else:
            _lib.check(self.lib.fa_front_half_ragged_dev(self._h, C.c_void_p(audio.data_ptr()), b, s, self._ilens(ilens, b),
                                                         self._ilens(phys, b), C.c_void_p(enc.data_ptr()),
                                                         C.c_void_p(adaptor.data_ptr()), C.c_void_p(ids.data_ptr())))
        return enc, adaptor, ids

    def collapse_cuda(self, ids):
        """ids [B][T] int32 CUDA -> (tokens [B][T], start_frames [B][T], counts [B]); nano_ctc.py:70-99."""
        import torch
        b, t = ids.shape
        tokens, starts = torch.empty_like(ids), torch.empty_like(ids)
        counts = torch.empty((b,), dtype=torch.int32, device=ids.device)
        _lib.check(self.lib.fa_ctc_collapse_dev(self._h, C.c_void_p(ids.data_ptr()), b, t, C.c_void_p(tokens.data_ptr()),
                                                C.c_void_p(starts.data_ptr()), C.c_void_p(counts.data_ptr())))
        return tokens, starts, counts

    def sync(self):
        _lib.check(self.lib.fa_ctx_sync(self._h))

    def launch_count(self) -> int:
        return int(self.lib.fa_launch_count())

    # ------------------------------------------------------------------ debug taps (tests)
    def enable_taps(self, on: bool = True):
        _lib.check(self.lib.fa_debug_enable_taps(self._h, 1 if on else 0))

    def read_tap(self, name: str) -> np.ndarray:
        r, c = C.c_int64(), C.c_int64()
        _lib.check(self.lib.fa_debug_read_tap(self._h, name.encode(), None, 0, C.byref(r), C.byref(c)))
        out = np.empty((r.value, c.value), np.float32)
        _lib.check(self.lib.fa_debug_read_tap(self._h, name.encode(), _ptr(out), out.size, C.byref(r), C.byref(c)))
        return out


def greedy_tokens(ids: np.ndarray, blank_id: int):
    """Host mirror of the device collapse for one segment: [(token, start_frame, start_seconds)].
    start = max((frame*60 - 240)/1000, 0) as in nano_ctc.py:67-68,99."""
    out, prev = [], None
    for i, tok in enumerate(ids.tolist()):
        if tok != prev:
            if tok != blank_id:
                out.append((tok, i, max((i * 60 - 240) / 1000.0, 0.0)))
            prev = tok
    return out


def profile_begin():
    """Start per-launch CUDA-event timing of this thread's kernel launches (bench.py roofline leg)."""
    _lib.check(_lib.load().fa_prof_begin())


def profile_end() -> dict:
    import json
    buf = C.create_string_buffer(1 << 16)
    _lib.check(_lib.load().fa_prof_end(buf, len(buf)))
    return json.loads(buf.value.decode())
