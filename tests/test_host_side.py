"""CPU tests of the host logic: the C ABI surface, the onnxruntime-shaped shim run against the
REAL nano_onnx.py / nano_ctc.py when the reference is mounted, and segment sharding."""
import ctypes
import importlib.util
import os
import re
import sys

import numpy as np
import pytest
import torch

from fun_asr_gguf_b200 import _lib, ort_shim, segments, weights as Wm
from oracle import oracle as O
from tests import signals

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/fun_asr_gguf"


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "funasr_b200.h")).read()
    declared = set(re.findall(r"FA_API\s+[\w\s\*]+?\b(fa_\w+)\s*\(", header))
    assert declared and declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.fa_abi_version() == 2


def test_shape_helpers_match_reference_arithmetic():
    lib = _lib.load()
    for s in (1, 159, 160, 700, 16000, 80000, 960000, 62 * 16000, 63877):
        assert lib.fa_frames_for_samples(s) == Wm.lfr_frames(s) == (s // 160 + 1 + 5) // 6
        assert lib.fa_adaptor_rows_for_samples(s) == Wm.adaptor_target_len(s) == O.target_len(s)
    assert lib.fa_frames_for_samples(960000) == 1001 and lib.fa_adaptor_rows_for_samples(960000) == 126


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_gpu_means_a_loud_error_not_a_fallback():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.fa_ctx_create(0, 1, 16000, 1, ctypes.byref(h)) != 0
    assert b"cuda" in lib.fa_last_error().lower()
    from fun_asr_gguf_b200 import FrontHalf
    with pytest.raises(RuntimeError):
        FrontHalf({}, device=0, max_batch=1, max_samples=16000)


def test_segment_windows_follow_the_orchestrator():
    sr = 16000
    assert segments.segment_windows(62 * sr) == [(0, 62 * sr)]                 # <= segment + 2 s: one pass
    w = segments.segment_windows(3600 * sr, 60.0, 4.0)                         # BASELINE config 4
    assert len(w) == 65 and w[0] == (0, 60 * sr) and w[1][0] == 56 * sr
    assert w[-1] == (3584 * sr, 3600 * sr) and all(b - a == 60 * sr for a, b in w[:-1])
    assert Wm.lfr_frames(w[-1][1] - w[-1][0]) == 267
    owned = [segments.shard(65, 8, r) for r in range(8)]
    assert sorted(i for o in owned for i in o) == list(range(65)) and max(map(len, owned)) == 9
    batches = segments.pack_batches([b - a for a, b in w], 32)
    assert [len(b) for b in batches] == [32, 32, 1] and batches[-1] == [64]
    audio = np.arange(3600 * sr, dtype=np.float32)
    batch, lens = segments.pad_batch(audio, w, [64, 0])
    assert batch.shape == (2, 60 * sr) and lens == [16 * sr, 60 * sr] and batch[0, 16 * sr:].sum() == 0


class _OracleEngine:
    """Stand-in for FrontHalf so the shim's session logic can run on a CPU-only box (tests only)."""
    max_samples = 62 * 16000

    def __init__(self, w, c):
        self.w, self.c = w, c

    @staticmethod
    def frames(s):
        return Wm.lfr_frames(s)

    def encode(self, audio, ilens):
        e, a = O.encode_batch(torch.from_numpy(audio), ilens, self.w, self.c)
        return e.numpy(), a.numpy()

    def ctc(self, enc):
        return O.ctc_ids_batch(torch.from_numpy(enc), self.w).numpy()

    max_batch = 2

    def front_half(self, audio, ilens):
        e, a = self.encode(audio, ilens)
        return e, a, self.ctc(e)


@pytest.fixture()
def shim(monkeypatch, weights, consts):
    eng = _OracleEngine(weights, consts)
    monkeypatch.setattr(ort_shim, "_engine_for", lambda path, min_samples=0: eng)
    monkeypatch.setitem(sys.modules, "onnxruntime", ort_shim)
    return ort_shim


def test_shim_surface(shim):
    so = shim.SessionOptions()
    so.add_session_config_entry("session.intra_op.allow_spinning", "0")
    so.graph_optimization_level = shim.GraphOptimizationLevel.ORT_ENABLE_ALL
    assert "DmlExecutionProvider" not in shim.get_available_providers()
    enc = shim.InferenceSession("model/Fun-ASR-Nano-Encoder-Adaptor.fp32.onnx", sess_options=so, providers=["CPUExecutionProvider"])
    ctc = shim.InferenceSession("model/Fun-ASR-Nano-CTC.fp32.onnx", sess_options=so, providers=["CPUExecutionProvider"])
    assert [(i.name, i.type) for i in enc.get_inputs()] == [("audio", "tensor(float)"), ("ilens", "tensor(int64)")]
    assert [o.name for o in enc.get_outputs()] == ["enc_output", "adaptor_output"]
    assert ctc.get_inputs()[0].name == "enc_output" and "float16" not in ctc.get_inputs()[0].type
    assert enc.get_providers()[0] == "CPUExecutionProvider"
    audio = signals.white(16000, 0).numpy().reshape(1, 1, -1)
    e, a = enc.run(None, {"audio": audio, "ilens": np.array([16000], np.int64)})
    assert e.shape == (1, 17, 512) and a.shape == (1, 17, 1024) and e.dtype == np.float32
    only = enc.run(["adaptor_output"], {"audio": audio, "ilens": np.array([16000], np.int64)})
    assert len(only) == 1 and np.array_equal(only[0], a)
    ids = ctc.run(None, {"enc_output": e})[0]
    assert ids.shape == (1, 17) and ids.dtype == np.int32
    with pytest.raises(ValueError):
        enc.run(None, {"wave": audio})
    with pytest.raises(ValueError):
        ctc.run(["logits"], {"enc_output": e})


def test_model_file_name_selects_the_precision_mode(monkeypatch):
    """02-Quantize-ONNX.py:14,34 names the variants; 04-Inference.py:42-43 loads the fp16 ones by default."""
    monkeypatch.delenv("FUNASR_B200_PRECISION", raising=False)
    assert ort_shim.precision_for("m/Fun-ASR-Nano-Encoder-Adaptor.fp32.onnx") == "bf16x3"
    assert ort_shim.precision_for("m/Fun-ASR-Nano-Encoder-Adaptor.fp16.onnx") == "bf16x3"
    assert ort_shim.precision_for("m/Fun-ASR-Nano-CTC.int8.onnx") == "fp8"
    monkeypatch.setenv("FUNASR_B200_PRECISION", "fp32")
    assert ort_shim.precision_for("m/Fun-ASR-Nano-CTC.int8.onnx") == "fp32"


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not mounted")
def test_unmodified_reference_call_sites_run_on_the_shim(shim, weights, consts):
    """load_onnx_models / encode_audio / decode_ctc of the reference, byte-for-byte, over our sessions."""
    def load(name):
        spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    nano_onnx, nano_ctc = load("nano_onnx"), load("nano_ctc")
    assert nano_onnx.onnxruntime is shim
    enc_sess, ctc_sess, _ = nano_onnx.load_onnx_models("m/Fun-ASR-Nano-Encoder-Adaptor.fp32.onnx",
                                                       "m/Fun-ASR-Nano-CTC.fp32.onnx", padding_secs=1)
    sig = signals.structured(2 * 16000 + 77, 3).numpy()
    audio_embd, enc_output = nano_onnx.encode_audio(sig, enc_sess)
    e_o, a_o = O.encode_one(torch.from_numpy(sig), sig.shape[0], weights, consts)
    assert audio_embd.shape == (Wm.adaptor_target_len(sig.shape[0]), 1024)
    assert np.array_equal(enc_output[0], e_o.numpy()) and np.array_equal(audio_embd, a_o[: audio_embd.shape[0]].numpy())
    ids = ctc_sess.run(None, {"enc_output": enc_output})[0]
    id2token = {i: chr(0x4E00 + i % 2000) for i in range(60515)}
    text, tokens, _ = nano_ctc.decode_ctc(ids, id2token)
    want = O.greedy_collapse(ids[0], 60514)
    assert [t.start for t in tokens] == [s for _, _, s in want] and len(text) == len(want)


def test_lookahead_prefetch_feeds_the_unchanged_sessions(shim, weights, consts):
    """every window of a long file is computed up front in equal-length batches; the per-segment
    session calls of the reference's loop then return those arrays, identical to computing them one by one."""
    from fun_asr_gguf_b200 import lookahead
    sr = 16000
    audio = signals.structured(7 * sr - 123, 5).numpy()
    lookahead.cache().clear()
    n = lookahead.prefetch(audio, segment_s=2.0, overlap_s=0.5)
    windows = segments.segment_windows(audio.shape[0], 2.0, 0.5)
    assert n == len(windows) == 5 and windows[-1][1] - windows[-1][0] < sr          # the last window is under 1 s
    enc_sess = shim.InferenceSession("m/Fun-ASR-Nano-Encoder-Adaptor.fp32.onnx")
    ctc_sess = shim.InferenceSession("m/Fun-ASR-Nano-CTC.fp32.onnx")
    hits0 = lookahead.cache().hits
    for a, b in windows:
        chunk = audio[a:b]
        n_phys = lookahead.physical_samples(b - a)                                 # encode_audio pads to 1 s on the CPU provider
        fed = np.zeros((1, 1, n_phys), np.float32)
        fed[0, 0, :b - a] = chunk
        e, ad = enc_sess.run(None, {"audio": fed, "ilens": np.array([b - a], np.int64)})
        ids = ctc_sess.run(None, {"enc_output": e})[0]
        e_o, a_o = O.encode_one(torch.from_numpy(fed[0, 0]), b - a, weights, consts)
        assert np.array_equal(e[0], e_o.numpy()) and np.array_equal(ad[0], a_o.numpy())
        assert np.array_equal(ids[0], O.ctc_ids_one(e_o, weights).numpy())
    assert lookahead.cache().hits - hits0 == 2 * len(windows)                      # every call was served from the cache
    # a segment that was not prefetched still computes
    other = signals.white(sr, 9).numpy().reshape(1, 1, -1)
    e2 = enc_sess.run(None, {"audio": other, "ilens": np.array([sr], np.int64)})[0]
    assert e2.shape == (1, 17, 512) and lookahead.cache().hits - hits0 == 2 * len(windows)
    lookahead.cache().clear()


def test_checkpoint_key_mapping_follows_load_weights(weights, monkeypatch):
    """load_checkpoint applies HybridSenseVoice.load_weights' mapping (model_definition.py:231-238): the three module
    prefixes pass through, ctc.ctc_lo.* becomes ctc_proj.ctc_lo.*, everything else (the LLM, the tokenizer
    embeddings) is ignored; a missing or mis-shaped tensor is an error, not a silent random init."""
    sd = {}
    for k, v in weights.items():
        sd[k.replace("ctc_proj.ctc_lo", "ctc.ctc_lo")] = v
    for k in ("ctc.ctc_lo.bias", "audio_adaptor.blocks.1.norm2.weight"):
        sd[k] = sd[k].to(torch.float64)                                            # dtype is normalised to fp32
    sd["llm.model.embed_tokens.weight"] = torch.zeros(4, 4)
    sd["ctc.ctc_lo.extra"] = torch.zeros(1)
    monkeypatch.setattr(torch, "load", lambda path, map_location=None: {"state_dict": sd})
    got = Wm.load_checkpoint("model.pt")
    assert set(got) == set(weights)
    for k in ("audio_encoder.encoders0.0.self_attn.linear_q_k_v.weight", "ctc_proj.ctc_lo.bias", "audio_adaptor.blocks.1.norm2.weight"):
        assert got[k].dtype == torch.float32 and torch.equal(got[k], weights[k])
    bad = dict(sd)
    del bad["audio_encoder.tp_norm.bias"]
    monkeypatch.setattr(torch, "load", lambda path, map_location=None: bad)
    with pytest.raises(KeyError):
        Wm.load_checkpoint("model.pt")
    bad = dict(sd)
    bad["ctc.ctc_lo.weight"] = torch.zeros(10, 512)
    monkeypatch.setattr(torch, "load", lambda path, map_location=None: bad)
    with pytest.raises(ValueError):
        Wm.load_checkpoint("model.pt")


def test_sessions_survive_an_engine_rebuild(monkeypatch):
    """a segment longer than FUNASR_B200_MAX_SECONDS replaces the shared engine.  Sessions resolve the engine
    on every run, the old engine is closed only after the new one exists, the capacity is rounded up to whole seconds so
    the CTC session's request does not force a second rebuild, and the checkpoint tensors are not re-read."""
    from fun_asr_gguf_b200 import engine as E
    built, loads = [], []

    class FakeFrontHalf:
        def __init__(self, tensors, device=0, max_batch=4, max_samples=0, precision="bf16x3"):
            self.max_samples, self.max_batch, self.closed = max_samples, max_batch, False
            built.append(self)

        def close(self):
            self.closed = True

        @staticmethod
        def frames(s):
            return Wm.lfr_frames(s)

        def encode(self, audio, ilens):
            assert not self.closed and audio.shape[1] <= self.max_samples
            b, t = audio.shape[0], Wm.lfr_frames(audio.shape[1])
            return np.zeros((b, t, 512), np.float32), np.zeros((b, t, 1024), np.float32)

        def ctc(self, enc):
            assert not self.closed and enc.shape[1] <= Wm.lfr_frames(self.max_samples)
            return np.zeros(enc.shape[:2], np.int32)

    monkeypatch.setattr(E, "FrontHalf", FakeFrontHalf)
    monkeypatch.setattr(Wm, "random_weights", lambda seed: loads.append(seed) or {})
    monkeypatch.setattr(ort_shim, "_engines", {})
    monkeypatch.setattr(ort_shim, "_tensors", {})
    monkeypatch.delenv("FUNASR_B200_MAX_SECONDS", raising=False)
    enc = ort_shim.InferenceSession("m/Fun-ASR-Nano-Encoder-Adaptor.fp32.onnx")
    ctc = ort_shim.InferenceSession("m/Fun-ASR-Nano-CTC.fp32.onnx")
    assert len(built) == 1 and built[0].max_samples == 62 * 16000
    s = int(63.4 * 16000) + 7
    e, _ = enc.run(None, {"audio": np.zeros((1, 1, s), np.float32), "ilens": np.array([s], np.int64)})
    assert len(built) == 2 and built[0].closed and not built[1].closed and built[1].max_samples == 64 * 16000
    ctc.run(None, {"enc_output": e})                       # same engine: no second rebuild, and not the closed one
    assert len(built) == 2 and loads == [0]
    enc.run(None, {"audio": np.zeros((1, 1, 16000), np.float32)})
    assert len(built) == 2
    for t in (1, 17, 1001, 1034):
        assert Wm.lfr_frames(ort_shim.samples_for_frames(t)) == t and (t == 1 or Wm.lfr_frames(ort_shim.samples_for_frames(t) - 1) == t - 1)


def test_lookahead_puts_every_window_of_a_file_into_ragged_batches(weights, consts):
    """With an engine that takes per-segment physical lengths (fa_front_half_ragged), run_file batches ALL windows of a
    file — the short tail included — instead of only windows of equal physical length, and hands every window back in
    the shapes of its own physical length.  (Stand-in engine: the oracle run row by row at phys[b]; the CUDA engine's
    ragged call is checked against that same definition in tests/test_gpu_parity.py.)"""
    from fun_asr_gguf_b200 import lookahead
    calls = []

    class Ragged(_OracleEngine):
        supports_ragged = True
        max_batch = 3

        def front_half(self, audio, ilens, phys=None):
            calls.append((audio.shape, list(ilens), list(phys)))
            b, s = audio.shape
            t = Wm.lfr_frames(s)
            enc, ad, ids = np.zeros((b, t, 512), np.float32), np.zeros((b, t, 1024), np.float32), np.full((b, t), -1, np.int32)
            for r in range(b):
                e, a = O.encode_one(torch.from_numpy(audio[r, :phys[r]]), ilens[r], self.w, self.c)
                tp = e.shape[0]
                enc[r, :tp], ad[r, :tp], ids[r, :tp] = e.numpy(), a.numpy(), O.ctc_ids_one(e, self.w).numpy()
            return enc, ad, ids

    sr = 16000
    audio = signals.structured(7 * sr - 123, 5).numpy()
    res = lookahead.run_file(Ragged(weights, consts), audio, segment_s=2.0, overlap_s=0.5)
    windows = segments.segment_windows(audio.shape[0], 2.0, 0.5)
    assert len(res) == len(windows) == 5 and len(calls) == 2                       # 5 windows, batches of <= 3: sizes 3 + 2
    assert sorted(len(c[1]) for c in calls) == [2, 3] and all(c[2] != [c[2][0]] * len(c[2]) or len(set(c[2])) == 1 for c in calls)
    assert any(len(set(c[2])) > 1 for c in calls)                                  # the tail shares a batch with full windows
    for (a, b), r in zip(windows, res):
        n_phys = lookahead.physical_samples(b - a)
        fed = np.zeros(n_phys, np.float32)
        fed[:b - a] = audio[a:b]
        e_o, a_o = O.encode_one(torch.from_numpy(fed), b - a, weights, consts)
        assert r.enc_output.shape == (1, Wm.lfr_frames(n_phys), 512) and r.n_phys == n_phys
        assert np.array_equal(r.enc_output[0], e_o.numpy()) and np.array_equal(r.adaptor_output[0], a_o.numpy())
        assert np.array_equal(r.ids[0], O.ctc_ids_one(e_o, weights).numpy())
