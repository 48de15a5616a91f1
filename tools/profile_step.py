#!/usr/bin/env python
"""Short driver for ncu and for per-kernel CUDA-event breakdowns: builds one context, runs
W warm-up steps and K steps of the hot path (encode + CTC) on synthetic 60 s segments.

    python tools/profile_step.py --batch 32 --steps 1 --warmup 1 [--breakdown] [--seconds 60]

Each step launches the same kernel sequence as a bench.py step.  Nothing here is a bench value.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--breakdown", action="store_true", help="print per-kernel CUDA-event totals of one extra step")
    ap.add_argument("--per-launch", action="store_true", help="with --breakdown: also dump every launch")
    ap.add_argument("--cuda-profiler", action="store_true", help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
    ap.add_argument("--mixed", action="store_true", help="BASELINE configs[2]: lengths randint(80 000, 960 001), seed 1234 (padding-free path)")
    args = ap.parse_args()

    import torch
    from fun_asr_gguf_b200 import FrontHalf, weights as Wm, engine as E
    from fun_asr_gguf_b200 import synth as signals

    s = int(args.seconds * 16000)
    dev = torch.device("cuda", 0)
    eng = FrontHalf(Wm.random_weights(0), device=0, max_batch=args.batch, max_samples=s, precision=args.precision)
    eng.use_torch_stream()
    audio = torch.stack([signals.white(s, i) for i in range(args.batch)]).to(dev)
    t = eng.frames(s)
    enc = torch.empty((args.batch, t, 512), dtype=torch.float32, device=dev)
    ad = torch.empty((args.batch, t, 1024), dtype=torch.float32, device=dev)
    ids = torch.empty((args.batch, t), dtype=torch.int32, device=dev)
    ilens = [s] * args.batch
    if args.mixed:
        g = torch.Generator().manual_seed(1234)
        ilens = [min(int(v), s) for v in torch.randint(80_000, 960_001, (args.batch,), generator=g)]
        for i, n in enumerate(ilens):
            audio[i, n:] = 0

    def step():
        eng.front_half_cuda(audio, ilens, enc, ad, ids)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    if args.cuda_profiler:
        torch.cuda.profiler.start()
    e0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    enqueue_ms = (time.perf_counter() - t0) * 1e3 / max(args.steps, 1)
    e1.record()
    torch.cuda.synchronize()
    if args.cuda_profiler:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1) / max(args.steps, 1)
    print(json.dumps({"batch": args.batch, "seconds": args.seconds, "ms_per_step": ms, "cpu_enqueue_ms_per_step": enqueue_ms,
                      "audio_s_per_s": args.batch * args.seconds / (ms * 1e-3),
                      "launches_per_step": (eng.launch_count() - l0) // max(args.steps, 1),
                      "ids_checksum": int(ids.to(torch.int64).sum().item())}))
    if args.breakdown:
        E.profile_begin()
        step()
        torch.cuda.synchronize()
        prof = E.profile_end()
        tot = sum(v["ms"] for v in prof.values())
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0
            print(f"{k:28s} launches {int(v['launches']):4d}  ms {v['ms']:8.3f}  share {v['ms'] / tot:6.1%}  alg TFLOP/s {tf:7.1f}")
    eng.close()


if __name__ == "__main__":
    main()
