"""The seeded parity cases shared by the golden generator, the oracle tests and the GPU tests."""
from tests import signals

SR = 16000

# name -> (signal builder, n_valid samples, n_phys samples)
CASES = {
    # 3 s of white noise at its native length
    "native3": (lambda: signals.white(3 * SR, 0), 3 * SR, 3 * SR),
    # 5 s of structured audio zero-padded to 8 s: masks, replicate padding, length control
    "padded5in8": (lambda: signals.structured(5 * SR, 7), 5 * SR, 8 * SR),
    # ragged lengths that are multiples of neither the hop nor the LFR stride
    "ragged": (lambda: signals.structured(4 * SR - 37, 9), 4 * SR - 37, 4 * SR + 123),
    # the warm-up input of nano_onnx.py:56 — exercises log(1e-7)
    "zeros2": (lambda: signals.white(2 * SR, 1) * 0.0, 2 * SR, 2 * SR),
    # shortest the engine can meet in practice: under one LFR window of valid audio in 1 s
    "tiny": (lambda: signals.white(700, 2), 700, SR),
}

# full-size case (BASELINE.json config 1, synthetic stand-in for input.mp3 — SURVEY F4)
SIXTY = ("sixty", lambda: signals.structured(60 * SR, 21), 60 * SR, 60 * SR)


def build(name):
    fn, n_valid, n_phys = CASES[name]
    return signals.padded(fn()[:n_valid], n_phys), n_valid
