#!/usr/bin/env python
"""Is the difference between a step and the sum of its kernels launch gaps or clock throttling?
Times single steps separated by idle pauses against back-to-back steps (experiments only)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fun_asr_gguf_b200 import FrontHalf, weights as Wm
from fun_asr_gguf_b200 import synth as signals

B, S = 32, 960000
dev = torch.device("cuda", 0)
eng = FrontHalf(Wm.random_weights(0), device=0, max_batch=B, max_samples=S)
eng.use_torch_stream()
audio = torch.stack([signals.white(S, i) for i in range(B)]).to(dev)
t = eng.frames(S)
enc = torch.empty((B, t, 512), dtype=torch.float32, device=dev)
ad = torch.empty((B, t, 1024), dtype=torch.float32, device=dev)
ids = torch.empty((B, t), dtype=torch.int32, device=dev)
il = [S] * B
def step():
    eng.encode_cuda(audio, il, enc, ad); eng.ctc_cuda(enc, ids)
def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for _ in range(2): step()
torch.cuda.synchronize()
singles = []
for _ in range(5):
    time.sleep(1.0)
    singles.append(timed(1))
print("single steps after 1 s idle:", [round(x, 2) for x in singles])
print("8 back-to-back:", round(timed(8), 2), " 16 back-to-back:", round(timed(16), 2))
