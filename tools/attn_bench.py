#!/usr/bin/env python
"""Time the attention kernel alone at the encoder's shape through profile hooks (experiments only)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests.gpu_util import RawContext
from fun_asr_gguf_b200 import engine as E

def main():
    b, t, heads, dk = int(os.environ.get("AB", 8)), 1001, int(os.environ.get("AH", 4)), int(os.environ.get("ADK", 128))
    raw = RawContext()
    rng = np.random.default_rng(0)
    qkv = (rng.standard_normal((b * t, 3 * heads * dk)) * 0.7).astype(np.float32)
    for dbg in sys.argv[1:]:
        os.environ["FUNASR_B200_ATTN_DBG"] = dbg
        raw.attention(qkv, b, t, heads, dk, None, precision="bf16x3")
        E.profile_begin()
        for _ in range(3):
            raw.attention(qkv, b, t, heads, dk, None, precision="bf16x3")
        prof = E.profile_end()
        v = prof["k_attention_tc"]
        print(f"dbg={dbg:>3s}  batch {b} heads {heads} dk {dk}: {v['ms'] / v['launches'] * 1e3:8.1f} us per launch", flush=True)
    raw.close()

main()
