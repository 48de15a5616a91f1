#!/usr/bin/env python
"""Power, clocks and throttle reasons sampled every 20 ms while steps run back to back (experiments only)."""
import os, sys, time, subprocess, statistics
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
for _ in range(2):
    eng.encode_cuda(audio, il, enc, ad); eng.ctc_cuda(enc, ids)
torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,power.draw.instant,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                      "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
t0 = time.time()
n = 0
while time.time() - t0 < 4.0:
    eng.encode_cuda(audio, il, enc, ad); eng.ctc_cuda(enc, ids); n += 1
    if n % 4 == 0: torch.cuda.synchronize()
torch.cuda.synchronize()
el = time.time() - t0
p.terminate()
rows = [l.strip().split(", ") for l in p.stdout.read().strip().splitlines() if l.strip()]
rows = [r for r in rows if len(r) >= 7]
busy = rows[20:-5]
print(f"{n} steps in {el:.2f} s = {el / n * 1e3:.1f} ms/step; {len(busy)} samples under load")
print("sm MHz   median", statistics.median(float(r[0]) for r in busy), "min", min(float(r[0]) for r in busy), "max", max(float(r[0]) for r in busy))
print("power W  avg", round(statistics.mean(float(r[1]) for r in busy), 1), "max", max(float(r[1]) for r in busy),
      "| instant avg", round(statistics.mean(float(r[2]) for r in busy), 1), "max", max(float(r[2]) for r in busy))
print("temp C   max", max(float(r[3]) for r in busy), " sw_power_cap active in", sum(r[4].startswith("Active") for r in busy), "samples; hw_slowdown", sum(r[5].startswith("Active") for r in busy), "; sw_thermal", sum(r[6].startswith("Active") for r in busy))
