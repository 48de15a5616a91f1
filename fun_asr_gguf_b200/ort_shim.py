"""Drop-in for the slice of the ``onnxruntime`` Python API that the reference touches on this path.

The reference runs the front half through two ``onnxruntime.InferenceSession`` objects
(fun_asr_gguf/nano_onnx.py:21-76 builds and warms them, :78-133 runs the encoder,
core/decoder.py:27 runs the CTC head).  This module mirrors exactly that surface — same names,
argument meaning and error behaviour (exceptions) — and executes on the B200 library instead:

    SessionOptions().add_session_config_entry / .graph_optimization_level   nano_onnx.py:26-29
    GraphOptimizationLevel.ORT_ENABLE_ALL                                   nano_onnx.py:29
    get_available_providers()                                               nano_onnx.py:32
    InferenceSession(path, sess_options=..., providers=[...])               nano_onnx.py:35-45
        .get_inputs() / .get_outputs()  -> objects with .name .type .shape  nano_onnx.py:54,60,67,82,104
        .get_providers()                                                    nano_onnx.py:90
        .run(None, {name: ndarray})                                         nano_onnx.py:62,73; decoder.py:27
        .run_with_ort_values(names, {name: OrtValue}) -> [.numpy()]         nano_onnx.py:117-129
    OrtValue.ortvalue_from_numpy(arr, 'cpu', 0)                             nano_onnx.py:109-114

``install()`` registers the module as ``onnxruntime`` so that ``nano_onnx.py`` imports it unchanged.

Two things differ from a real ORT session and are deliberate:
  * the "model file" is only a name: which graph a session runs is read from the file name
    ("CTC" in it -> CTC head, else encoder+adaptor, the names 01-Export...py:36-37 writes) and the
    weights come from a FunASR ``model.pt`` found next to it (or $FUNASR_B200_WEIGHTS); with no
    checkpoint present a seeded random initialisation is used and a warning is logged;
  * ``get_providers()[0]`` is 'CPUExecutionProvider'.  nano_onnx.py:90 uses that string to mean
    "no shape-recompile cost, pad only to 1 s", which is true here, and it makes the physical
    length — which the unmasked CTC head's ids depend on — the one the reference's
    FP32 CPU run sees.  The second entry, 'B200ExecutionProvider', says what actually runs.
"""
from __future__ import annotations

import logging
import os
import sys
import threading
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import weights as W

log = logging.getLogger("fun_asr_gguf_b200")

__version__ = "0.0-b200"


class GraphOptimizationLevel:
    ORT_DISABLE_ALL, ORT_ENABLE_BASIC, ORT_ENABLE_EXTENDED, ORT_ENABLE_ALL = 0, 1, 2, 99


class SessionOptions:
    def __init__(self):
        self.graph_optimization_level = GraphOptimizationLevel.ORT_ENABLE_ALL
        self.intra_op_num_threads = 0
        self.inter_op_num_threads = 0
        self.log_severity_level = 2
        self._entries: Dict[str, str] = {}

    def add_session_config_entry(self, key: str, value: str) -> None:
        self._entries[str(key)] = str(value)

    def get_session_config_entry(self, key: str) -> str:
        return self._entries[key]


def get_available_providers() -> List[str]:
    return ["B200ExecutionProvider", "CPUExecutionProvider"]


def get_device() -> str:
    return "GPU-B200"


class NodeArg:
    def __init__(self, name: str, type_: str, shape):
        self.name, self.type, self.shape = name, type_, shape

    def __repr__(self):
        return f"NodeArg(name='{self.name}', type='{self.type}', shape={self.shape})"


class OrtValue:
    def __init__(self, array: np.ndarray, device: str = "cpu"):
        self._a, self._device = array, device

    @staticmethod
    def ortvalue_from_numpy(array: np.ndarray, device_type: str = "cpu", device_id: int = 0) -> "OrtValue":
        if not isinstance(array, np.ndarray):
            raise TypeError("ortvalue_from_numpy expects a numpy.ndarray")
        return OrtValue(array, device_type)

    def numpy(self) -> np.ndarray:
        return self._a

    def shape(self):
        return list(self._a.shape)

    def device_name(self) -> str:
        return self._device


# --------------------------------------------------------------------------------------- shared engine

_engines: Dict[tuple, "object"] = {}
_lock = threading.Lock()
_run_lock = threading.RLock()          # serialises session calls on the shared context (and an engine swap against a running call)


def _find_checkpoint(model_path: str) -> Optional[str]:
    env = os.environ.get("FUNASR_B200_WEIGHTS")
    if env:
        return env
    d = os.path.dirname(os.path.abspath(model_path))
    for cand in (os.path.join(d, "model.pt"), os.path.join(d, "..", "Fun-ASR-Nano-2512", "model.pt"),
                 os.path.join(d, "Fun-ASR-Nano-2512", "model.pt")):
        if os.path.isfile(cand):
            return os.path.normpath(cand)
    return None


def precision_for(model_path: str) -> str:
    """The reference ships three builds of each graph and picks one by file name (02-Quantize-ONNX.py:14,34;
    04-Inference.py:42-43 defaults to ``.fp16.onnx``): ``.fp32.`` the traced FP32 graph, ``.fp16.`` everything but
    LayerNorm in half precision, ``.int8.`` per-channel dynamic QUInt8 MatMuls.  Here: the fp32 and fp16
    names run the bf16x3 mode — 16 mantissa bits per operand, i.e. at least the fp16 graph's accuracy and token-exact
    against the FP32 one — and the int8 name, whose user has asked for 8-bit MatMuls with per-channel weight scales,
    runs the fp8 mode (e4m3 x e4m3 on tcgen05 kind::f8f6f4, per-output-channel weight scales, LayerNorm and softmax in
    fp32 as 02-Quantize-ONNX.py:26 keeps them).  ``$FUNASR_B200_PRECISION`` overrides (fp32, bf16x3, bf16, fp8)."""
    env = os.environ.get("FUNASR_B200_PRECISION")
    if env:
        return env
    return "fp8" if ".int8." in os.path.basename(model_path).lower() else "bf16x3"


_tensors: Dict[tuple, "object"] = {}      # checkpoint tensors by key: growing an engine does not re-read the file


def samples_for_frames(frames: int) -> int:
    """Fewest samples whose segment has `frames` LFR frames: ceil((s//160 + 1) / 6) >= T  <=>  s >= (6T - 6) * 160."""
    return max(1, (W.LFR_N * frames - W.LFR_N) * W.HOP)


def _engine_for(model_path: str, min_samples: int = 0):
    """The engine the sessions of `model_path` share, with room for segments of `min_samples`.

    Sessions call this on EVERY run and never keep the result: when a longer segment than planned for arrives the
    engine is replaced (the new one is built first, then the old one is closed), and a session holding on to the old
    one would be left with a closed context.  Capacity is rounded up to whole seconds, so the CTC session's request
    for the frames the encoder session just produced never forces a second rebuild."""
    from .engine import FrontHalf

    ckpt = _find_checkpoint(model_path)
    device = int(os.environ.get("FUNASR_B200_DEVICE", "0"))
    precision = precision_for(model_path)
    max_batch = int(os.environ.get("FUNASR_B200_MAX_BATCH", "4"))
    key = (ckpt or "random:0", device, precision)
    with _lock:
        eng = _engines.get(key)
        want = max(min_samples, int(os.environ.get("FUNASR_B200_MAX_SECONDS", "62")) * W.SAMPLE_RATE)
        want = -(-want // W.SAMPLE_RATE) * W.SAMPLE_RATE
        if eng is not None and eng.max_samples >= want:
            return eng
        tensors = _tensors.get(key)
        if tensors is None:
            if ckpt is None:
                log.warning("no FunASR checkpoint near %s: using seeded random-init weights of the architecture", model_path)
                tensors = W.random_weights(0)
            else:
                tensors = W.load_checkpoint(ckpt)
            _tensors[key] = tensors
        if eng is not None:
            log.info("segment of %d samples exceeds the engine's %d: rebuilding with more room", min_samples, eng.max_samples)
        new = FrontHalf(tensors, device=device, max_batch=max_batch, max_samples=want, precision=precision)
        _engines[key] = new
        if eng is not None:
            eng.close()
        return new


def shutdown() -> None:
    with _lock:
        for e in _engines.values():
            e.close()
        _engines.clear()
        _tensors.clear()


# --------------------------------------------------------------------------------------- sessions

class InferenceSession:
    """One of the reference's two sessions, picked by file name, over a shared B200 context."""

    def __init__(self, path_or_bytes, sess_options: Optional[SessionOptions] = None,
                 providers: Optional[Sequence] = None, provider_options=None, **kwargs):
        if not isinstance(path_or_bytes, (str, os.PathLike)):
            raise TypeError("this backend identifies the graph by model file name; pass a path")
        self._path = os.fspath(path_or_bytes)
        self._options = sess_options or SessionOptions()
        self._requested = list(providers or [])
        self._role = "ctc" if "ctc" in os.path.basename(self._path).lower() else "encoder"
        _engine_for(self._path)                  # builds (or finds) the shared context now, like ORT loads the model now
        if self._role == "encoder":
            self._inputs = [NodeArg("audio", "tensor(float)", [1, 1, "samples"]),
                            NodeArg("ilens", "tensor(int64)", ["batch"])]
            self._outputs = [NodeArg("enc_output", "tensor(float)", [1, "enc_frames", W.D_ENC]),
                             NodeArg("adaptor_output", "tensor(float)", [1, "adaptor_frames", W.D_LLM])]
        else:
            self._inputs = [NodeArg("enc_output", "tensor(float)", [1, "enc_len", W.D_ENC])]
            self._outputs = [NodeArg("indices", "tensor(int32)", [1, "enc_len"])]

    # -- introspection
    def get_inputs(self): return list(self._inputs)
    def get_outputs(self): return list(self._outputs)
    def get_providers(self): return ["CPUExecutionProvider", "B200ExecutionProvider"]
    def get_session_options(self): return self._options
    def get_modelmeta(self): return type("ModelMeta", (), {"producer_name": "fun_asr_gguf_b200", "graph_name": self._role})()

    # -- execution
    def _select(self, output_names, results: Dict[str, np.ndarray]):
        names = [o.name for o in self._outputs] if not output_names else list(output_names)
        for n in names:
            if n not in results:
                raise ValueError(f"Invalid output name: {n}")
        return [results[n] for n in names]

    def _run(self, feed: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
        for name in feed:
            if name not in {i.name for i in self._inputs}:
                raise ValueError(f"Invalid input name: {name}")
        if self._role == "encoder":
            if "audio" not in feed:
                raise ValueError("Required input 'audio' is missing")
            audio = np.asarray(feed["audio"])
            if audio.ndim != 3 or audio.shape[1] != 1:
                raise ValueError(f"audio must be (batch, 1, samples); got {audio.shape}")
            b, _, s = audio.shape
            ilens = np.asarray(feed.get("ilens", np.full((b,), s, np.int64))).astype(np.int64).reshape(-1)
            if ilens.shape[0] != b:
                raise ValueError("ilens must have one entry per batch row")
            if b == 1:                          # a window lookahead.prefetch() has already computed
                from . import lookahead
                hit = lookahead.cache().encoder_lookup(audio.reshape(s), int(ilens[0]))
                if hit is not None:
                    return {"enc_output": hit.enc_output, "adaptor_output": hit.adaptor_output}
            with _run_lock:          # one context serves both sessions: one call at a time (ORT sessions are re-entrant)
                enc, ad = _engine_for(self._path, min_samples=s).encode(audio.reshape(b, s).astype(np.float32, copy=False), ilens.tolist())
            return {"enc_output": enc, "adaptor_output": ad}
        if "enc_output" not in feed:
            raise ValueError("Required input 'enc_output' is missing")
        enc = np.asarray(feed["enc_output"])
        if isinstance(feed["enc_output"], np.ndarray):
            from . import lookahead
            hit = lookahead.cache().ctc_lookup(feed["enc_output"])
            if hit is not None:
                return {"indices": hit.ids}
        if enc.ndim != 3 or enc.shape[2] != W.D_ENC:
            raise ValueError(f"enc_output must be (batch, frames, {W.D_ENC}); got {enc.shape}")
        with _run_lock:
            eng = _engine_for(self._path, min_samples=samples_for_frames(enc.shape[1]))
            return {"indices": eng.ctc(enc.astype(np.float32, copy=False))}

    def run(self, output_names, input_feed: Dict[str, np.ndarray], run_options=None):
        return self._select(output_names, self._run(dict(input_feed)))

    def run_with_ort_values(self, output_names, input_dict_ort_values: Dict[str, OrtValue], run_options=None):
        feed = {k: (v.numpy() if isinstance(v, OrtValue) else np.asarray(v)) for k, v in input_dict_ort_values.items()}
        return [OrtValue(a) for a in self._select(output_names, self._run(feed))]


def install(force: bool = False) -> None:
    """Make ``import onnxruntime`` resolve to this module (the reference imports it by that name,
    nano_onnx.py:1).  A real onnxruntime already imported is left alone unless force=True."""
    if "onnxruntime" in sys.modules and sys.modules["onnxruntime"] is not sys.modules[__name__] and not force:
        return
    sys.modules["onnxruntime"] = sys.modules[__name__]
