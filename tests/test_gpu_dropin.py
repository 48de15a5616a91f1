"""The reference's UNMODIFIED host-side call sites over the CUDA engine.

`tools/stage_reference.py` (run by `__graft_entry__.build()` wherever /root/reference is mounted) copies
nano_onnx.py, nano_ctc.py, core/decoder.py and the small modules they import, byte for byte, into the
git-ignored baseline/_ref/fun_asr_gguf/, which travels to the GPU box with the snapshot.  Here
`ort_shim.install()` makes `import onnxruntime` resolve to the shim, and then the reference's own code runs:

    nano_onnx.load_onnx_models  (:21-76, warm-up with 60 s of zeros)
    nano_onnx.encode_audio      (:78-133, CPU-provider padding rule, OrtValue feed, target_len slice)
    core.decoder.CTCDecoder.decode -> ctc_sess.run + nano_ctc.decode_ctc   (decoder.py:19-48)
    the per-segment loop of core/orchestrator.py:128-171 (windows cut exactly as there) after lookahead.prefetch

all against FrontHalf on cuda:0; results are compared with the oracle.
"""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

from fun_asr_gguf_b200 import lookahead, ort_shim, segments, synth, weights as Wm
from oracle import oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref", "fun_asr_gguf")
SR = 16000
ACT_TOL = 3e-4


@pytest.fixture(scope="module")
def ref():
    """The staged reference package, imported WITHOUT its __init__ (which pulls in llama.cpp's shared library):
    a bare package object whose __path__ is the staged directory, and a stub for the one sibling module that needs
    libllama."""
    if not os.path.isfile(os.path.join(STAGED, "nano_onnx.py")):
        pytest.skip("baseline/_ref/fun_asr_gguf is not staged (python tools/stage_reference.py in the build container)")
    saved = {k: sys.modules.get(k) for k in ("onnxruntime", "fun_asr_gguf", "fun_asr_gguf.llama", "fun_asr_gguf.core",
                                             "fun_asr_gguf.core.model_manager")}
    ort_shim.install(force=True)
    pkg = types.ModuleType("fun_asr_gguf"); pkg.__path__ = [STAGED]
    core = types.ModuleType("fun_asr_gguf.core"); core.__path__ = [os.path.join(STAGED, "core")]
    llama = types.ModuleType("fun_asr_gguf.llama")                   # LLM side: out of scope, never called here
    mm = types.ModuleType("fun_asr_gguf.core.model_manager"); mm.ModelManager = object
    sys.modules.update({"fun_asr_gguf": pkg, "fun_asr_gguf.core": core, "fun_asr_gguf.llama": llama,
                        "fun_asr_gguf.core.model_manager": mm})
    pkg.llama = llama
    mods = types.SimpleNamespace(nano_onnx=importlib.import_module("fun_asr_gguf.nano_onnx"),
                                 nano_ctc=importlib.import_module("fun_asr_gguf.nano_ctc"),
                                 decoder=importlib.import_module("fun_asr_gguf.core.decoder"))
    assert mods.nano_onnx.onnxruntime is ort_shim
    yield mods
    ort_shim.shutdown()
    lookahead.cache().clear()
    for k in list(sys.modules):
        if k == "fun_asr_gguf" or k.startswith("fun_asr_gguf."):
            del sys.modules[k]
    for k, v in saved.items():
        if v is not None:
            sys.modules[k] = v
        else:
            sys.modules.pop(k, None)


@pytest.fixture(scope="module")
def sessions(ref):
    os.environ.pop("FUNASR_B200_PRECISION", None)
    os.environ["FUNASR_B200_MAX_BATCH"] = "8"
    enc_sess, ctc_sess, t_cost = ref.nano_onnx.load_onnx_models("model/Fun-ASR-Nano-Encoder-Adaptor.fp32.onnx",
                                                                "model/Fun-ASR-Nano-CTC.fp32.onnx", padding_secs=60)
    assert t_cost > 0
    return enc_sess, ctc_sess


class _Models:
    """What CTCDecoder reads from ModelManager (core/decoder.py:19-48)."""
    corrector = None

    def __init__(self, ctc_sess):
        self.ctc_sess = ctc_sess
        self.ctc_id2token = {i: chr(0x4E00 + i % 20000) for i in range(Wm.VOCAB - 1)}
        self.ctc_id2token[Wm.VOCAB - 1] = "<blk>"


def _oracle(sig, n_phys, weights, consts):
    fed = torch.zeros(n_phys)
    fed[: sig.shape[0]] = torch.from_numpy(sig)
    enc_o, ad_o = O.encode_one(fed, sig.shape[0], weights, consts)
    return enc_o, ad_o, O.ctc_logits_one(enc_o, weights)


def test_load_models_encode_audio_and_ctc_decoder_on_the_cuda_engine(ref, sessions, weights, consts):
    enc_sess, ctc_sess = sessions
    assert enc_sess.get_providers()[0] == "CPUExecutionProvider"            # => encode_audio pads to 1 s only (nano_onnx.py:90-93)
    dec = ref.decoder.CTCDecoder(_Models(ctc_sess))
    for seconds, seed in ((7.3, 3), (0.4, 4), (60.0, 5)):
        sig = synth.structured(int(seconds * SR), seed).numpy()
        audio_embd, enc_output = ref.nano_onnx.encode_audio(sig, enc_sess)
        n_phys = max(sig.shape[0], SR)
        enc_o, ad_o, logits = _oracle(sig, n_phys, weights, consts)
        tl = Wm.adaptor_target_len(sig.shape[0])
        assert enc_output.shape == (1, Wm.lfr_frames(n_phys), 512) and audio_embd.shape == (tl, 1024)
        e_err = float(np.abs(enc_output[0] - enc_o.numpy()).max())
        a_err = float(np.abs(audio_embd - ad_o[:tl].numpy()).max())
        print(f"[dropin {seconds} s] enc max|d| {e_err:.3e} adaptor max|d| {a_err:.3e}")
        assert e_err <= ACT_TOL and a_err <= ACT_TOL
        results, hotwords, stats = dec.decode(enc_output, True, 10)
        ids_o = logits.argmax(-1).numpy()
        want = O.greedy_collapse(ids_o, Wm.VOCAB - 1)
        assert [r.start for r in results] == [s for _, _, s in want]
        assert "".join(r.text for r in results) == "".join(chr(0x4E00 + t % 20000) for t, _, _ in want)
        assert hotwords == [] and stats["infer"] > 0
        # the raw session output is what decode_ctc's ids branch needs: int32 (1, T)
        ids = ctc_sess.run(None, {"enc_output": enc_output})[0]
        assert ids.dtype == np.int32 and ids.shape == (1, enc_output.shape[1]) and np.array_equal(ids[0], ids_o)


def test_a_segment_longer_than_planned_rebuilds_the_engine_and_both_sessions_survive(ref, sessions, weights, consts, monkeypatch):
    """segment_size is user-configurable in the reference; the shim plans for FUNASR_B200_MAX_SECONDS (62).  A longer
    segment replaces the shared engine; the encoder AND the CTC session must keep working afterwards."""
    enc_sess, ctc_sess = sessions
    sig = synth.structured(int(63.5 * SR), 6).numpy()
    before = dict(ort_shim._engines)
    audio_embd, enc_output = ref.nano_onnx.encode_audio(sig, enc_sess)
    ids = ctc_sess.run(None, {"enc_output": enc_output})[0]
    after = dict(ort_shim._engines)
    assert list(before) == list(after) and all(before[k] is not after[k] for k in before), "the engine was not replaced"
    assert all(e._h is None for e in before.values()), "the replaced engine was not closed"
    assert len({id(e) for e in after.values()}) == 1, "the CTC session forced a second rebuild"
    enc_o, ad_o, logits = _oracle(sig, sig.shape[0], weights, consts)
    assert np.abs(enc_output[0] - enc_o.numpy()).max() <= ACT_TOL
    assert np.array_equal(ids[0], logits.argmax(-1).numpy())
    # and a short segment afterwards, through both sessions again
    sig2 = synth.white(2 * SR, 8).numpy()
    embd2, enc2 = ref.nano_onnx.encode_audio(sig2, enc_sess)
    ids2 = ctc_sess.run(None, {"enc_output": enc2})[0]
    enc_o2, _, logits2 = _oracle(sig2, 2 * SR, weights, consts)
    assert np.abs(enc2[0] - enc_o2.numpy()).max() <= ACT_TOL and np.array_equal(ids2[0], logits2.argmax(-1).numpy())


def test_lookahead_prefetch_then_the_orchestrators_segment_loop(ref, sessions, weights, consts):
    """core/orchestrator.py:128-171: a long file is cut into windows and each goes through encode_audio and the CTC
    session in turn.  lookahead.prefetch(audio) computes every window up front in equal-length batches on the engine the
    sessions share; the unchanged per-segment calls are then served from the cache — same arrays as computing them
    one by one — and a window that was not prefetched still computes."""
    enc_sess, ctc_sess = sessions
    dec = ref.decoder.CTCDecoder(_Models(ctc_sess))
    seg_s, ov_s = 20.0, 2.0
    audio = synth.structured(int(70.7 * SR), 12).numpy()
    lookahead.cache().clear()
    n = lookahead.prefetch(audio, "model/Fun-ASR-Nano-Encoder-Adaptor.fp32.onnx", segment_s=seg_s, overlap_s=ov_s)
    # the loop of _transcribe_long, verbatim arithmetic
    duration = len(audio) / SR
    info, step, curr = [], seg_s - ov_s, 0.0
    while curr < duration:
        end = min(curr + seg_s, duration)
        info.append((curr, end))
        if end >= duration:
            break
        curr += step
    assert n == len(info) == len(segments.segment_windows(len(audio), seg_s, ov_s)) == 4
    hits0 = lookahead.cache().hits
    for s_s, e_s in info:
        chunk = audio[int(s_s * SR):int(e_s * SR)]
        audio_embd, enc_output = ref.nano_onnx.encode_audio(chunk, enc_sess)
        results, _, _ = dec.decode(enc_output, True, 10)
        enc_o, ad_o, logits = _oracle(chunk, max(len(chunk), SR), weights, consts)
        tl = Wm.adaptor_target_len(len(chunk))
        assert np.abs(enc_output[0] - enc_o.numpy()).max() <= ACT_TOL
        assert np.abs(audio_embd - ad_o[:tl].numpy()).max() <= ACT_TOL
        want = O.greedy_collapse(logits.argmax(-1).numpy(), Wm.VOCAB - 1)
        assert [r.start for r in results] == [s for _, _, s in want]
    assert lookahead.cache().hits - hits0 == 2 * len(info), "the per-segment calls were not served from the prefetch cache"
    # prefetched results are bit-identical to computing the same window without the cache
    a, b = int(info[1][0] * SR), int(info[1][1] * SR)
    embd_cached, enc_cached = ref.nano_onnx.encode_audio(audio[a:b], enc_sess)
    lookahead.cache().clear()
    embd_direct, enc_direct = ref.nano_onnx.encode_audio(audio[a:b], enc_sess)
    assert np.array_equal(enc_cached, enc_direct) and np.array_equal(embd_cached, embd_direct)
    assert np.array_equal(ctc_sess.run(None, {"enc_output": enc_direct})[0][0], O.ctc_logits_one(
        _oracle(audio[a:b], b - a, weights, consts)[0], weights).argmax(-1).numpy())
