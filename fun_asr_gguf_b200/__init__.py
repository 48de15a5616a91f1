"""B200-native audio front half of Fun-ASR-Nano (fbank/LFR -> SAN-M encoder -> adaptor -> CTC greedy ids).

    from fun_asr_gguf_b200 import FrontHalf           # batched engine over the C ABI
    from fun_asr_gguf_b200 import ort_shim            # onnxruntime-shaped sessions for nano_onnx.py

The compute lives in libfunasr_b200.so (hand-written CUDA for sm_100a, built by
`python -m fun_asr_gguf_b200.build`); there is no CPU fallback.
"""
from . import weights  # noqa: F401


def __getattr__(name):
    if name == "FrontHalf":
        from .engine import FrontHalf
        return FrontHalf
    if name in ("ort_shim", "lookahead", "segments"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
