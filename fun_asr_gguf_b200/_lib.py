"""ctypes binding of include/funasr_b200.h.  There is no CPU fallback: if the shared library is
missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# $FUNASR_B200_LIB: another build of the same sources (tuning aid for same-box A/Bs: tools/_ab/*.so)
LIB_PATH = os.environ.get("FUNASR_B200_LIB") or os.path.join(HERE, "libfunasr_b200.so")

PREC = {"fp32": 0, "bf16x3": 1, "bf16": 2, "fp8": 3}

c_float_p = C.POINTER(C.c_float)
c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)

# name -> (restype, argtypes); must list every function include/funasr_b200.h declares
SIGNATURES = {
    "fa_abi_version": (C.c_int, []),
    "fa_last_error": (C.c_char_p, []),
    "fa_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "fa_frames_for_samples": (C.c_int64, [C.c_int64]),
    "fa_adaptor_rows_for_samples": (C.c_int64, [C.c_int64]),
    "fa_ctx_create": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "fa_ctx_destroy": (C.c_int, [C.c_void_p]),
    "fa_ctx_load_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, c_i64_p, C.c_int]),
    "fa_ctx_finalize": (C.c_int, [C.c_void_p]),
    "fa_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fa_ctx_sync": (C.c_int, [C.c_void_p]),
    "fa_ctx_vocab": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "fa_launch_count": (C.c_int64, []),
    "fa_prof_begin": (C.c_int, []),
    "fa_prof_end": (C.c_int, [C.c_char_p, C.c_int64]),
    "fa_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, c_i64_p, C.c_void_p, C.c_void_p]),
    "fa_encode_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, c_i64_p, C.c_void_p, C.c_void_p]),
    "fa_ctc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "fa_ctc_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "fa_front_half": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, c_i64_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fa_front_half_ragged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, c_i64_p, c_i64_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fa_front_half_ragged_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, c_i64_p, c_i64_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "fa_front_half_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, c_i64_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fa_front_half_embd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, c_i64_p, C.c_void_p, C.POINTER(C.c_void_p), c_i64_p,
                           C.c_void_p]),
    "fa_ctc_collapse_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fa_debug_enable_taps": (C.c_int, [C.c_void_p, C.c_int]),
    "fa_debug_read_tap": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, c_i64_p, c_i64_p]),
    "fa_test_linear": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fa_test_vocab_argmax": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p]),
    "fa_test_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p]),
    "fa_test_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                     C.c_void_p, C.c_void_p]),
    "fa_test_fsmn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fa_test_front_end": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, c_i64_p, C.c_void_p, C.c_void_p]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the in-tree library (never a site-packages copy) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} is missing: build it with `python -m fun_asr_gguf_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.fa_abi_version() != 2:
        raise RuntimeError(f"ABI mismatch: library reports {lib.fa_abi_version()}, binding expects 2")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("funasr_b200: " + load().fa_last_error().decode("utf-8", "replace"))
