"""Helpers for the -m gpu tests: a bare context for the kernel-level hooks, torch references."""
import ctypes as C

import numpy as np

from fun_asr_gguf_b200 import _lib


def P(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class RawContext:
    """An un-finalised context: enough for the fa_test_* kernel hooks (they need a device + stream)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        _lib.check(self.lib.fa_ctx_create(device, 1, 16000, 1, C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.fa_ctx_destroy(self.h)
            self.h = None

    def linear(self, a, w, bias, resid=None, relu=False, precision="fp32", planes=False):
        m, k = a.shape
        n = w.shape[0]
        out = np.empty((m, n), np.float32)
        pl = np.empty((m, n), np.float32) if planes else None
        _lib.check(self.lib.fa_test_linear(self.h, P(a), P(w), P(bias), P(resid), m, n, k, int(relu),
                                           _lib.PREC[precision], P(out), P(pl)))
        return (out, pl) if planes else out

    def vocab_argmax(self, a, w, bias, precision="bf16x3"):
        m, k = a.shape
        ids = np.empty((m,), np.int32)
        _lib.check(self.lib.fa_test_vocab_argmax(self.h, P(a), P(w), P(bias), m, w.shape[0], k, _lib.PREC[precision], P(ids)))
        return ids

    def attention(self, qkv, batch, frames, heads, dk, kv_len=None, precision="fp32"):
        out = np.empty((batch * frames, heads * dk), np.float32)
        kv = None if kv_len is None else np.asarray(kv_len, np.int32)
        _lib.check(self.lib.fa_test_attention(self.h, P(qkv), batch, frames, heads, dk, P(kv), _lib.PREC[precision], P(out)))
        return out

    def layernorm(self, x, gamma, beta, eps):
        out = np.empty_like(x)
        pl = np.empty_like(x)
        _lib.check(self.lib.fa_test_layernorm(self.h, P(x), x.shape[0], x.shape[1], P(gamma), P(beta), C.c_float(eps), P(out), P(pl)))
        return out, pl

    def fsmn(self, v, w, t_valid, resid=None):
        b, t, _ = v.shape
        out = np.empty_like(v)
        tv = np.asarray(t_valid, np.int32)
        _lib.check(self.lib.fa_test_fsmn(self.h, P(v), P(w), P(tv), b, t, P(resid), P(out)))
        return out


def rel_err(got, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))
