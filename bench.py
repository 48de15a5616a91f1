#!/usr/bin/env python
"""bench.py — encoder+adaptor+CTC throughput of the front half, in audio-seconds per second.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference path on the host cores

A step is one pass of the whole hot path (fbank/LFR -> 70 SAN-M layers -> adaptor -> CTC head ->
greedy ids) over one batch of BASELINE.json configs[1]: 32 x 60 s synthetic segments per GPU.
N > 1 is segment-level data parallelism: every rank runs its own batch, no collective on the path
(weak scaling); the timed region is bracketed by a barrier + synchronize and the max over ranks
is reported.  Rank 0 prints ONE JSON line.

`value`  : inputs already resident in HBM, CUDA-event timed on the launching stream.
`e2e`    : the same metric through the host-buffer C-ABI call (fa_front_half): pinned host audio in,
           enc_output + adaptor_output + ids back to host, copies inside the timed region.
`roofline`: the dominant kernel (the tcgen05 projection GEMM), per-launch CUDA-event timed in an
           extra profiled step right after the timed region.
`cpu_baseline` / --impl reference: the oracle (a torch fp32 port of the reference's
           model_definition.py — onnxruntime itself is not installable here) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
SEG_S = 60
BATCH = 32
METRIC = "encoder+adaptor+CTC throughput (fbank -> SAN-M encoder -> adaptor -> CTC greedy ids), batched 60 s segments"
UNIT = "audio-s/s"
WORKLOAD = "configs[1]: batch 32 x 60 s synthetic segments, fbank+encoder+adaptor+CTC greedy on 1 B200 (per GPU)"


def synth_batch(n_seg: int, first_index: int):
    import torch
    from tests import signals
    return torch.stack([signals.white(SEG_S * SR, first_index + i) for i in range(n_seg)])


# ------------------------------------------------------------------------------------------ clocks

class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ reference arm

def oracle_rate(n_steps: int, warmup: int):
    """One step = one 60 s segment of the batch through the oracle on all host threads."""
    import torch
    from fun_asr_gguf_b200 import weights as Wm
    from oracle import oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    w = Wm.random_weights(0)
    consts = Wm.front_end_constants(1100)
    audio = synth_batch(1, 0)[0]
    times = []
    for i in range(warmup + n_steps):
        t0 = time.perf_counter()
        enc, _ = O.encode_one(audio, audio.shape[0], w, consts)
        O.ctc_ids_one(enc, w)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return SEG_S * len(times) / sum(times), sum(times) / len(times), torch.get_num_threads()


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = min(args.steps, 8), min(args.warmup, 2)
    value, sec_per_step, cores = oracle_rate(steps, warmup)
    sample = (f"1 of the {BATCH} segments of the step's batch (one 60 s segment per step), {steps} steps after {warmup} warm-up; "
              f"torch fp32 eager port of model_definition.py on {cpu_model()} — a PyTorch-eager stand-in for the ONNX Runtime CPU "
              "path (onnxruntime is not installable in this image)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "segments_per_step": 1, "segment_s": SEG_S},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm

def run_ours(args) -> None:
    import torch
    import torch.distributed as dist
    from fun_asr_gguf_b200 import FrontHalf, weights as Wm
    from fun_asr_gguf_b200 import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: keep NCCL's version banner / debug output on stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    batch = args.batch
    s = SEG_S * SR

    eng = FrontHalf(Wm.random_weights(0), device=local, max_batch=batch, max_samples=s, precision=args.precision)
    eng.use_torch_stream()
    # two distinct input batches, alternated: 2 x 123 MB of audio plus ~2.5 GB of activations per step,
    # far beyond the 126 MB L2, so no step starts with its inputs cached
    host = [synth_batch(batch, rank * 1000 + j * batch).pin_memory() for j in range(2)]
    dev_in = [h.to(dev, non_blocking=True) for h in host]
    ilens = [s] * batch
    t = eng.frames(s)
    enc = torch.empty((batch, t, 512), dtype=torch.float32, device=dev)
    ad = torch.empty((batch, t, 1024), dtype=torch.float32, device=dev)
    ids = torch.empty((batch, t), dtype=torch.int32, device=dev)

    def step(i):
        eng.encode_cuda(dev_in[i & 1], ilens, enc, ad)
        eng.ctc_cuda(enc, ids)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    fence()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    fence()
    launches = eng.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    clocks = sampler.stop() if rank == 0 else None
    value = world * batch * SEG_S * args.steps / (ms * 1e-3)

    # ---- e2e: host-buffer C-ABI call, copies inside the timed region
    h_enc = torch.empty((batch, t, 512), dtype=torch.float32).pin_memory()
    h_ad = torch.empty((batch, t, 1024), dtype=torch.float32).pin_memory()
    h_ids = torch.empty((batch, t), dtype=torch.int32).pin_memory()
    import ctypes as C
    from fun_asr_gguf_b200 import _lib

    def e2e_step(i):
        arr = (C.c_int64 * batch)(*ilens)
        _lib.check(eng.lib.fa_front_half(eng._h, C.c_void_p(host[i & 1].data_ptr()), batch, s, arr,
                                         C.c_void_p(h_enc.data_ptr()), C.c_void_p(h_ad.data_ptr()), C.c_void_p(h_ids.data_ptr())))

    e2e_step(0)
    fence()
    n_e2e = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for i in range(n_e2e):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = world * batch * SEG_S * n_e2e / e2e_s
    h2d = batch * s * 4
    d2h = h_enc.numel() * 4 + h_ad.numel() * 4 + h_ids.numel() * 4

    # ---- roofline of the dominant kernel: one extra step with per-launch CUDA events
    prof = None
    if rank == 0:
        E.profile_begin()
        step(0)
        torch.cuda.synchronize()
        prof = E.profile_end()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    traffic, traffic_src, ncu_pipe = None, None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        traffic, traffic_src = tr["traffic_bytes_per_launch_mean"], tr["source"]
        ncu_pipe = tr.get("tensor_pipe_active_pct_time_weighted")
    except (OSError, ValueError, KeyError):
        pass
    total_ms = sum(v["ms"] for v in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1]["ms"])
    gemms = [k for k in prof if k.startswith("k_gemm_tc")] or [k for k in prof if "gemm" in k] or [top[0]]
    gemm_name = max(gemms, key=lambda k: prof[k]["ms"])          # the projection kernel that takes most of the step
    g = prof[gemm_name]
    mma_factor = {"bf16x3": 3, "bf16": 1, "fp32": 1}[args.precision]
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PF sustained"
    if args.precision == "fp32":
        peak_tf, peak_src = 75.0, "nominal fp32 CUDA-core peak (fp32 mode is the arbiter, not the product path)"
    ach = g["flops"] / (g["ms"] * 1e-3) / 1e12
    roofline = {
        "kernel": gemm_name, "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
        "traffic": traffic, "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean of the four encoder-layer shapes)",
        "traffic_source": traffic_src, "ncu_tensor_pipe_active_pct": ncu_pipe, "peak_source": peak_src,
        "launches_per_step": g["launches"], "avg_launch_ms": g["ms"] / g["launches"],
        "algorithmic_flops_per_step": g["flops"], "share_of_step": g["ms"] / total_ms,
        "mma_flops_factor": mma_factor, "issued_tflops": ach * mma_factor, "issued_frac": ach * mma_factor / peak_tf,
        "note": "achieved counts the algorithmic 2MNK of the fp32 GEMM the reference runs; bf16x3 issues 3 MMAs per product",
        "step_breakdown_ms": {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
    }

    # the bandwidth-bound kernels against the measured copy bandwidth (algorithmic bytes noted at launch; a per-launch
    # event pair adds a few microseconds to each of these short kernels, so `achieved` is a lower bound)
    hbm_peak = peaks.get("hbm_gbs") or 6500.0
    roofline["hbm_kernels"] = {
        k: {"bound": "hbm", "achieved": v["bytes"] / (v["ms"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": v["bytes"] / (v["ms"] * 1e-3) / 1e9 / hbm_peak, "launches_per_step": v["launches"],
            "avg_launch_us": 1e3 * v["ms"] / v["launches"], "algorithmic_bytes_per_launch": v["bytes"] / v["launches"]}
        for k, v in prof.items() if v.get("bytes", 0) > 0 and v["ms"] > 0}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, sec, cores = oracle_rate(12, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"1 of the {batch} segments of a step (60 s), 12 runs after 1 warm-up, {sec:.2f} s each; torch fp32 eager port of "
                         f"model_definition.py on {cpu_model()} (stand-in for ONNX Runtime CPU, which is not installable here)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16x3 (projections and attention: bf16 hi+lo planes on tcgen05, 3 MMAs per product, fp32 accumulate; fp32 softmax, LayerNorm, FSMN, front end)",
                  "bf16": "bf16 (projections plain bf16 on tcgen05, attention bf16x3, fp32 accumulate; speed mode, not token-exact)", "fp32": "f32"}[args.precision],
        "data": "synthetic (0.1*N(0,1) clipped, seed 1234+i); random-init weights of the architecture (no checkpoint ships)",
        "config": {"workload": WORKLOAD, "segments_per_step_per_gpu": batch, "segment_s": SEG_S, "frames_per_segment": t,
                   "precision": args.precision, "parallelism": f"segment-dp{world}",
                   "l2": "inputs alternate between two 123 MB batches and a step streams ~2.5 GB of activations, both > 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_e2e,
                "api": "fa_front_half (host buffers, pinned)"},
        "gpu_launches": launches,
        "roofline": roofline,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
