"""ORACLE — test infrastructure only.  Nothing under ``fun_asr_gguf_b200/`` may import this.

A CPU restatement, in functional torch, of the arithmetic the reference's two exported
graphs perform (the graphs ``nano_onnx.py`` runs through ONNX Runtime).  It exists to check
the CUDA path; it is never the thing shipped or measured (except as the labelled
``cpu_baseline`` / ``--impl reference`` leg of bench.py).

What it follows (all paths relative to /root/reference/fun_asr_gguf/):
  front_end()      model_definition.py:269-311  (EncoderExportWrapperPaddable.forward, steps 0-3)
                   model_definition.py:244-256  (STFT_Process)
  sanm_encoder()   model_definition.py:205-214  (SenseVoiceEncoderSmall.forward)
                   model_definition.py:100-116  (EncoderLayerSANM.forward)
                   model_definition.py:60-90    (MultiHeadedAttentionSANM)
  projector()      model_definition.py:179-185  (CorrectTransformerAdaptor.forward)
                   model_definition.py:154-163  (EncoderLayer.forward), :132-145 (MultiHeadedAttention)
  encode_one()     model_definition.py:313-323  (adaptor + length control)
  ctc_ids_one()    model_definition.py:335-337  (CTCHeadExportWrapper.forward; mask=None)
  greedy_collapse() nano_ctc.py:62-99           (collapse, blank drop, start time)

Pinning status: the reference ships no tests, fixtures or golden vectors, so the
pins are outputs of the reference's own ``model_definition.py`` executed in the build
container (PyTorch eager FP32 — onnxruntime is not installable there), stored under
tests/golden/ by tests/golden/make_golden.py.  tests/test_oracle_golden.py checks this file
against them.  The named oracle "FP32 ONNX graph on ONNX Runtime CPU" itself could not be
run; the residual ORT-vs-eager difference is unmeasured.

Semantics the reference fixes that this file must keep:
  * one segment at a time: batches are a loop of single-row calls (F8);
  * layer 0 returns right after attention+FSMN, no residual, no FFN (F9);
  * the CTC head is unmasked and therefore sees every physical frame (F7);
  * adaptor rows >= target_len are zeroed (F10).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

HOP, N_FFT, N_MELS, LFR_M, LFR_N = 160, 400, 80, 7, 6
W = Dict[str, torch.Tensor]


def _cast(w: W, dtype: torch.dtype) -> W:
    return w if dtype == torch.float32 else {k: v.to(dtype) for k, v in w.items()}


# ----------------------------------------------------------------------------- front end

def front_end(audio: torch.Tensor, n_valid: int, consts: W, taps: Optional[dict] = None
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """audio (S_phys,) -> (x (T,560) masked LFR features, m (T,) frame mask)."""
    dt = audio.dtype
    s_phys = audio.shape[0]
    idx = torch.arange(s_phys)
    smask = (idx < n_valid).to(dt)
    # mean over the valid samples only, removed, then re-masked
    mean = (audio * smask).sum() / n_valid
    a = (audio - mean) * smask
    # y[0] = a[0]; y[n] = a[n] - 0.97 a[n-1]; re-masked
    coef = torch.tensor(0.97, dtype=torch.float32).to(dt)
    a = torch.cat([a[:1], a[1:] - coef * a[:-1]]) * smask
    if taps is not None:
        taps["preemph"] = a.clone()
    # centre-padded framing, windowed DFT as two correlations with stride 160
    xp = F.pad(a.view(1, 1, -1), (N_FFT // 2, N_FFT // 2))
    re = F.conv1d(xp, consts["const.dft_cos"].to(dt).unsqueeze(1), stride=HOP)
    im = F.conv1d(xp, consts["const.dft_sin"].to(dt).unsqueeze(1), stride=HOP)
    power = re * re + im * im                                       # (1, 201, T_mel)
    mel = (torch.matmul(consts["const.mel_fbank"].to(dt).unsqueeze(0), power).transpose(1, 2) + 1e-7).log()[0]
    if taps is not None:
        taps["logmel"] = mel.clone()
    t_mel_phys = mel.shape[0]
    t_mel_valid = n_valid // HOP + 1
    t_valid = (t_mel_valid + LFR_N - 1) // LFR_N
    t_phys = (t_mel_phys + LFR_N - 1) // LFR_N
    # stacked frame t, slot i reads mel row clamp(6t+i-3) with replicate padding both sides,
    # rows past the valid length first folded onto the last valid row
    src = torch.arange(t_phys).unsqueeze(1) * LFR_N + torch.arange(LFR_M).unsqueeze(0) - (LFR_M - 1) // 2
    src = src.clamp(0, t_mel_phys - 1).clamp(max=t_mel_valid - 1)
    x = mel[src].reshape(t_phys, LFR_M * N_MELS)
    m = (torch.arange(t_phys) < t_valid).to(dt)
    x = x * m.unsqueeze(-1)
    if taps is not None:
        taps["lfr"] = x.clone()
    return x, m


# ----------------------------------------------------------------------------- blocks

def _ln(x, w: W, name: str, eps: float):
    return F.layer_norm(x, (x.shape[-1],), w[name + ".weight"], w[name + ".bias"], eps)


def _lin(x, w: W, name: str):
    return F.linear(x, w[name + ".weight"], w[name + ".bias"])


def _key_mask_add(m: Optional[torch.Tensor]):
    return None if m is None else (m - 1.0).view(1, 1, -1) * 10000.0


def _attend(q, k, v, heads: int, m: Optional[torch.Tensor]):
    """q,k,v (T, D) -> (T, D): scaled dot-product over `heads`, additive key mask."""
    t, d = q.shape
    dk = d // heads
    qh = q.view(t, heads, dk).transpose(0, 1) * (dk ** -0.5)
    kh = k.view(t, heads, dk).transpose(0, 1)
    vh = v.view(t, heads, dk).transpose(0, 1)
    scores = torch.matmul(qh, kh.transpose(-2, -1))
    add = _key_mask_add(m)
    if add is not None:
        scores = scores + add.view(1, 1, -1)
    p = torch.softmax(scores, dim=-1)
    return torch.matmul(p, vh).transpose(0, 1).reshape(t, d)


def sanm_layer(x, m, w: W, p: str, first: bool):
    """One EncoderLayerSANM. `first` = the 560->512 layer that skips residual and FFN."""
    h = _ln(x, w, p + ".norm1", 1e-5)
    qkv = _lin(h, w, p + ".self_attn.linear_q_k_v")
    q, k, v = torch.split(qkv, 512, dim=-1)
    vm = v * m.unsqueeze(-1)
    mem = F.conv1d(F.pad(vm.t().unsqueeze(0), (5, 5)), w[p + ".self_attn.fsmn_block.weight"], groups=512)[0].t() + vm
    att = _lin(_attend(q, k, v, 4, m), w, p + ".self_attn.linear_out")
    y = att + mem
    if first:
        return y
    x = x + y
    h = _ln(x, w, p + ".norm2", 1e-5)
    return x + _lin(F.relu(_lin(h, w, p + ".feed_forward.w_1")), w, p + ".feed_forward.w_2")


def sanm_encoder(x, m, w: W, pos: torch.Tensor, taps: Optional[dict] = None, n_layers: Optional[int] = None):
    """x (T,560), m (T,) -> enc (T,512)."""
    x = x * (512 ** 0.5) + pos[: x.shape[0]].to(x.dtype)
    done = 0
    x = sanm_layer(x, m, w, "audio_encoder.encoders0.0", True)
    done += 1
    if taps is not None:
        taps["layer0"] = x.clone()
    for i in range(49):
        if n_layers is not None and done >= n_layers:
            return x
        x = sanm_layer(x, m, w, f"audio_encoder.encoders.{i}", False)
        done += 1
        if taps is not None and i == 0:
            taps["layer1"] = x.clone()
    x = _ln(x, w, "audio_encoder.after_norm", 1e-5) * m.unsqueeze(-1)
    if taps is not None:
        taps["layer49"] = x.clone()
    for i in range(20):
        x = sanm_layer(x, m, w, f"audio_encoder.tp_encoders.{i}", False)
    return _ln(x, w, "audio_encoder.tp_norm", 1e-5) * m.unsqueeze(-1)


def mha_block(x, m, w: W, p: str, heads: int):
    h = _ln(x, w, p + ".norm1", 1e-12)
    a = _attend(_lin(h, w, p + ".self_attn.linear_q"), _lin(h, w, p + ".self_attn.linear_k"),
                _lin(h, w, p + ".self_attn.linear_v"), heads, m)
    x = x + _lin(a, w, p + ".self_attn.linear_out")
    h = _ln(x, w, p + ".norm2", 1e-12)
    return x + _lin(F.relu(_lin(h, w, p + ".feed_forward.w_1")), w, p + ".feed_forward.w_2")


def projector(x, m, w: W, p: str, n_blocks: int, heads: int):
    x = _lin(F.relu(_lin(x, w, p + ".linear1")), w, p + ".linear2")
    for i in range(n_blocks):
        x = mha_block(x, m, w, f"{p}.blocks.{i}", heads)
    return x


def target_len(n_valid: int) -> int:
    t = (n_valid // HOP + 1 + LFR_N - 1) // LFR_N
    o1 = 1 + (t - 3 + 2) // 2
    return (1 + (o1 - 3 + 2) // 2 - 1) // 2 + 1


# ----------------------------------------------------------------------------- graphs

@torch.no_grad()
def encode_one(audio: torch.Tensor, n_valid: int, w: W, consts: W, dtype=torch.float32,
               taps: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Encoder graph for one segment: audio (S_phys,) -> enc (T,512), adaptor_output (T,1024)."""
    w, consts = _cast(w, dtype), _cast(consts, dtype)
    x, m = front_end(audio.to(dtype), int(n_valid), consts, taps)
    enc = sanm_encoder(x, m, w, consts["const.pos_enc"], taps)
    ad = projector(enc, m, w, "audio_adaptor", 2, 8)
    keep = (torch.arange(ad.shape[0]) < target_len(int(n_valid))).to(ad.dtype).unsqueeze(-1)
    return enc, ad * keep


@torch.no_grad()
def ctc_logits_one(enc: torch.Tensor, w: W, dtype=torch.float32, taps: Optional[dict] = None) -> torch.Tensor:
    w = _cast(w, dtype)
    h = projector(enc.to(dtype), None, w, "ctc_decoder", 5, 8)
    if taps is not None:
        taps["ctc_h"] = h.clone()
    return _lin(h, w, "ctc_proj.ctc_lo")


@torch.no_grad()
def ctc_ids_one(enc: torch.Tensor, w: W, dtype=torch.float32) -> torch.Tensor:
    """CTC graph for one segment: enc (T,512) -> ids (T,) int32 (first index on ties)."""
    return torch.argmax(ctc_logits_one(enc, w, dtype), dim=-1).to(torch.int32)


def encode_batch(audio: torch.Tensor, ilens, w: W, consts: W, dtype=torch.float32):
    """audio (B, S_phys): each row run on its own at the batch's physical length (F7, F8)."""
    encs, ads = zip(*(encode_one(audio[b], int(ilens[b]), w, consts, dtype) for b in range(audio.shape[0])))
    return torch.stack(encs), torch.stack(ads)


def ctc_ids_batch(enc: torch.Tensor, w: W, dtype=torch.float32) -> torch.Tensor:
    return torch.stack([ctc_ids_one(enc[b], w, dtype) for b in range(enc.shape[0])])


# ----------------------------------------------------------------------------- greedy collapse

def greedy_collapse(ids, blank_id: int) -> List[Tuple[int, int, float]]:
    """ids (T,) -> [(token_id, start_frame, start_seconds)] with repeats merged and blanks dropped.

    start_seconds = max((frame*60 - 240)/1000, 0) as in nano_ctc.py:67-68,99.  (The reference
    also drops ids missing from its vocabulary file; that filter is host text work.)
    """
    out: List[Tuple[int, int, float]] = []
    prev = None
    for i, tok in enumerate(int(v) for v in ids):
        if tok != prev:
            if tok != blank_id:
                out.append((tok, i, max((i * 60 - 240) / 1000.0, 0.0)))
            prev = tok
    return out


# ----------------------------------------------------------------------------- work accounting

def flops(t_valid: int, t_phys: Optional[int] = None) -> float:
    """Matmul/conv FLOPs of both graphs for one segment:
    542.39e6*T + 161792*T^2, encoder/adaptor terms on valid frames, CTC terms on physical."""
    t_phys = t_valid if t_phys is None else t_phys
    ctc_lin, ctc_att = 15.99e6 + 61.97e6, 4 * 5 * 512
    lin, att = 542.39e6 - ctc_lin, 161792 - ctc_att
    return lin * t_valid + att * t_valid ** 2 + ctc_lin * t_phys + ctc_att * t_phys ** 2
