#!/usr/bin/env python
"""Time the attention kernel alone at the encoder's shape through the profile hooks (experiments only).

    AB=32 AH=4 ADK=128 python tools/attn_bench.py
Short isolated launches run at whatever clock the idle GPU ramps to: compare runs of this tool with each other,
not with the per-kernel times of a full step.
"""
import os, sys
os.environ.setdefault("FUNASR_B200_TEST_ATTN_OUT", "planes")   # what the engine's launches write
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests.gpu_util import RawContext
from fun_asr_gguf_b200 import engine as E


def main():
    b, t, heads, dk = int(os.environ.get("AB", 8)), int(os.environ.get("AT", 1001)), int(os.environ.get("AH", 4)), int(os.environ.get("ADK", 128))
    raw = RawContext()
    rng = np.random.default_rng(0)
    qkv = (rng.standard_normal((b * t, 3 * heads * dk)) * 0.7).astype(np.float32)
    raw.attention(qkv, b, t, heads, dk, None, precision="bf16x3")
    E.profile_begin()
    for _ in range(5):
        raw.attention(qkv, b, t, heads, dk, None, precision="bf16x3")
    prof = E.profile_end()
    v = next(x for k, x in prof.items() if k.startswith("k_attention_tc"))
    flops = 4.0 * b * heads * t * t * dk
    us = v["ms"] / v["launches"] * 1e3
    print(f"batch {b} frames {t} heads {heads} dk {dk}: {us:8.1f} us per launch, {flops / us / 1e6:7.1f} algorithmic TFLOP/s")
    raw.close()


main()
