"""Deterministic synthetic audio: the workloads bench.py measures and the signals the tests and the golden generator use.

white():       0.1*N(0,1) clipped to [-1,1], seed = 1234 + index (SURVEY §8d "value distribution").
structured():  swept sinusoids under slow amplitude envelopes, -30 dB noise and silence gaps, so
               the mel bins span a realistic dynamic range and the CTC ids vary over time.
"""
import math

import torch


def white(n_samples: int, index: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(1234 + index)
    return (0.1 * torch.randn(n_samples, generator=g)).clamp(-1.0, 1.0)


def structured(n_samples: int, seed: int = 7) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples, dtype=torch.float64) / 16000.0
    dur = max(n_samples / 16000.0, 1e-3)
    x = torch.zeros(n_samples, dtype=torch.float64)
    for k in range(4):
        f0 = 200.0 + 700.0 * k + 300.0 * torch.rand((), generator=g).item()
        f1 = f0 * (1.5 + torch.rand((), generator=g).item())
        phase = 2 * math.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / dur)
        rate = 2.0 + 2.0 * torch.rand((), generator=g).item()
        env = 0.5 * (1 + torch.sin(2 * math.pi * rate * t + 6.28 * torch.rand((), generator=g).item()))
        x += 0.2 * env * torch.sin(phase)
    x += 10 ** (-30 / 20) * torch.randn(n_samples, generator=g, dtype=torch.float64)
    gate = (torch.sin(2 * math.pi * 0.7 * t) > -0.3).to(torch.float64)
    return (x * gate).clamp(-1.0, 1.0).to(torch.float32)


def padded(sig: torch.Tensor, n_phys: int) -> torch.Tensor:
    out = torch.zeros(n_phys, dtype=torch.float32)
    out[: sig.shape[0]] = sig
    return out
