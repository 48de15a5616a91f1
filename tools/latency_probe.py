#!/usr/bin/env python
"""Latency of the reference's own call pattern (BASELINE configs[0]): ONE 60 s segment per call, encoder session
then CTC session, host buffers in and out (nano_onnx.py:78-133, core/decoder.py:27).  A call is ~620 kernel
launches for a few milliseconds of GPU work, so the host variants replay it as CUDA graphs; this prints the wall
time per call with graph replay (default) and with eager launches (FUNASR_B200_GRAPH_MAX_BATCH=0)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from fun_asr_gguf_b200 import FrontHalf, weights as Wm
from fun_asr_gguf_b200 import synth as signals

S = 960000
w = Wm.random_weights(0)
audio = signals.structured(S, 21).numpy()[None]
out = {}
for mode in ("graph", "eager"):
    if mode == "eager":
        os.environ["FUNASR_B200_GRAPH_MAX_BATCH"] = "0"
    eng = FrontHalf(w, device=0, max_batch=1, max_samples=S, precision="bf16x3")
    for _ in range(3):
        enc, ad = eng.encode(audio, [S])
        ids = eng.ctc(enc)
    t = []
    for _ in range(20):
        t0 = time.perf_counter()
        enc, ad = eng.encode(audio, [S])
        t1 = time.perf_counter()
        ids = eng.ctc(enc)
        t2 = time.perf_counter()
        t.append((t1 - t0, t2 - t1))
    t = np.array(t) * 1e3
    out[mode] = {"encode_ms_median": float(np.median(t[:, 0])), "ctc_ms_median": float(np.median(t[:, 1])),
                 "call_pair_ms_median": float(np.median(t.sum(1))), "ids_checksum": int(ids.astype(np.int64).sum())}
    eng.close()
out["audio_s_per_s_single_stream"] = {m: 60.0 / (out[m]["call_pair_ms_median"] * 1e-3) for m in ("graph", "eager")}
print(json.dumps(out))
