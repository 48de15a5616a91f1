"""Planted CTC projection (test infrastructure; uses the oracle).

With plain random init the greedy ids are almost constant, which would make
"token-exact" a weak test and leave collapse/timestamps unexercised.  plant() rewrites only
ctc_proj.ctc_lo so that, on real signals, ids change every few frames, many ids repeat in
runs and the blank id wins a share of the frames:
  * only N_ACTIVE vocabulary rows (+ blank) stay live, scaled up by GAIN;
  * the bias is re-centred on the mean hidden state of a calibration signal;
  * the blank logit is a constant placed at a chosen quantile of the per-frame maxima.
"""
import torch

from oracle import oracle as O
from tests import signals

N_ACTIVE, GAIN, BLANK_Q = 48, 4.0, 0.5


def plant(w, consts, seed: int = 11, cal=None):
    """cal: calibration signal (default 6 s of structured audio); plant(..., cal=signals.white(...)) makes the ids vary
    on white noise, which is what bench.py's batches hold."""
    w = dict(w)
    if cal is None:
        cal = signals.structured(16000 * 6, seed=seed)
    enc, _ = O.encode_one(cal, cal.shape[0], w, consts)
    taps = {}
    O.ctc_logits_one(enc, w, taps=taps)
    h = taps["ctc_h"]
    h_mean = h.mean(0)
    W0, vocab = w["ctc_proj.ctc_lo.weight"], w["ctc_proj.ctc_lo.weight"].shape[0]
    g = torch.Generator().manual_seed(seed)
    live = torch.randperm(vocab - 1, generator=g)[:N_ACTIVE]
    W1 = torch.zeros_like(W0)
    W1[live] = W0[live] * GAIN
    b1 = torch.full((vocab,), -50.0)
    b1[live] = -(W1[live] @ h_mean)
    per_frame_max = ((h - h_mean) @ W1[live].t()).max(-1).values
    b1[vocab - 1] = torch.quantile(per_frame_max, BLANK_Q)
    w["ctc_proj.ctc_lo.weight"], w["ctc_proj.ctc_lo.bias"] = W1.contiguous(), b1.contiguous()
    return w
