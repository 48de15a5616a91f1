#!/usr/bin/env python
"""Per-kernel time of ONE 60 s segment (the reference's per-call batch): where a launch-bound call spends its time."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from fun_asr_gguf_b200 import FrontHalf, weights as Wm
from fun_asr_gguf_b200.engine import profile_begin, profile_end
from tests import signals

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = 960000
eng = FrontHalf(Wm.random_weights(0), device=0, max_batch=B, max_samples=S, precision="bf16x3")
audio = np.stack([signals.structured(S, 21 + i).numpy() for i in range(B)])
for _ in range(2):
    eng.front_half(audio, [S] * B)
profile_begin()
eng.front_half(audio, [S] * B)
p = profile_end()
tot = sum(v["ms"] for v in p.values())
rows = sorted(p.items(), key=lambda kv: -kv[1]["ms"])
print(json.dumps({"batch": B, "total_kernel_ms": tot,
                  "kernels": {k: {"launches": int(v["launches"]), "ms": round(v["ms"], 3), "us_per_launch": round(1e3 * v["ms"] / v["launches"], 2)}
                              for k, v in rows}}))
