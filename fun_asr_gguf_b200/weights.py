"""Weight inventory, seeded initialisation and host-side constants for the audio front half.

The tensor names are the keys of the reference's ``HybridSenseVoice.state_dict()``
(/root/reference/fun_asr_gguf/model_definition.py:223-238), so a real ``model.pt`` can be
loaded with the same prefix mapping ``load_weights`` applies (``ctc.ctc_lo.*`` ->
``ctc_proj.ctc_lo.*``).  No checkpoint ships with the reference, so the default is a
seeded random initialisation of the named architecture.  Every tensor is drawn from its own
generator seeded by (seed, crc32(name)): the same values come out on any machine with the
same torch build, independent of the order tensors are requested in.

Non-parameter constants (DFT kernels, mel filterbank, positional table) are produced here
with the same fp32 torch expressions the reference's export recipe uses
(model_definition.py:244-253, 01-Export-Encoder-Adaptor-CTC.py:101-102,
model_definition.py:13-21) and are handed to the C library as named tensors; the CUDA
side never regenerates them with device transcendentals.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Iterable, Tuple

import torch

# Geometry of the path (model_definition.py:190-200, 223-229; 01-Export...py:41-45).
SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
N_BINS = N_FFT // 2 + 1          # 201
N_MELS = 80
LFR_M = 7
LFR_N = 6
D_IN = N_MELS * LFR_M            # 560
D_ENC = 512
D_FFN = 2048
D_LLM = 1024
FSMN_K = 11
ENC_HEADS = 4
N_ENC0, N_ENC, N_TP = 1, 49, 20
ADAPTOR_BLOCKS, ADAPTOR_HEADS = 2, 8
CTC_BLOCKS, CTC_HEADS = 5, 8
VOCAB = 60515
BLANK_ID = VOCAB - 1             # "<blk>" is the last token (01-Export...py:78)
PRE_EMPHASIS = 0.97


def mel_frames(n_samples: int) -> int:
    """Frames of the centre-padded STFT: S//160 + 1 (model_definition.py:254-256)."""
    return n_samples // HOP + 1


def lfr_frames(n_samples: int) -> int:
    """LFR frames ceil(T_mel / 6) (model_definition.py:290-291)."""
    return (mel_frames(n_samples) + LFR_N - 1) // LFR_N


def adaptor_target_len(n_valid_samples: int) -> int:
    """Rows of adaptor_output that survive length control (model_definition.py:317-318)."""
    t = lfr_frames(n_valid_samples)
    o1 = 1 + (t - 3 + 2) // 2
    return (1 + (o1 - 3 + 2) // 2 - 1) // 2 + 1


# --------------------------------------------------------------------------------------
# Tensor inventory
# --------------------------------------------------------------------------------------

def _sanm_layer(prefix: str, d_in: int) -> Iterable[Tuple[str, Tuple[int, ...]]]:
    yield f"{prefix}.self_attn.linear_out.weight", (D_ENC, D_ENC)
    yield f"{prefix}.self_attn.linear_out.bias", (D_ENC,)
    yield f"{prefix}.self_attn.linear_q_k_v.weight", (3 * D_ENC, d_in)
    yield f"{prefix}.self_attn.linear_q_k_v.bias", (3 * D_ENC,)
    yield f"{prefix}.self_attn.fsmn_block.weight", (D_ENC, 1, FSMN_K)
    yield f"{prefix}.feed_forward.w_1.weight", (D_FFN, D_ENC)
    yield f"{prefix}.feed_forward.w_1.bias", (D_FFN,)
    yield f"{prefix}.feed_forward.w_2.weight", (D_ENC, D_FFN)
    yield f"{prefix}.feed_forward.w_2.bias", (D_ENC,)
    yield f"{prefix}.norm1.weight", (d_in,)
    yield f"{prefix}.norm1.bias", (d_in,)
    yield f"{prefix}.norm2.weight", (D_ENC,)
    yield f"{prefix}.norm2.bias", (D_ENC,)


def _mha_block(prefix: str, d: int) -> Iterable[Tuple[str, Tuple[int, ...]]]:
    for p in ("linear_q", "linear_k", "linear_v", "linear_out"):
        yield f"{prefix}.self_attn.{p}.weight", (d, d)
        yield f"{prefix}.self_attn.{p}.bias", (d,)
    yield f"{prefix}.feed_forward.w_1.weight", (d // 4, d)
    yield f"{prefix}.feed_forward.w_1.bias", (d // 4,)
    yield f"{prefix}.feed_forward.w_2.weight", (d, d // 4)
    yield f"{prefix}.feed_forward.w_2.bias", (d,)
    for n in ("norm1", "norm2"):
        yield f"{prefix}.{n}.weight", (d,)
        yield f"{prefix}.{n}.bias", (d,)


def _projector(prefix: str, d_out: int, n_blocks: int) -> Iterable[Tuple[str, Tuple[int, ...]]]:
    yield f"{prefix}.linear1.weight", (D_FFN, D_ENC)
    yield f"{prefix}.linear1.bias", (D_FFN,)
    yield f"{prefix}.linear2.weight", (d_out, D_FFN)
    yield f"{prefix}.linear2.bias", (d_out,)
    for i in range(n_blocks):
        yield from _mha_block(f"{prefix}.blocks.{i}", d_out)


def tensor_spec(vocab: int = VOCAB) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape for every parameter on the path, in state_dict order."""
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for i in range(N_ENC0):
        spec.update(_sanm_layer(f"audio_encoder.encoders0.{i}", D_IN))
    for i in range(N_ENC):
        spec.update(_sanm_layer(f"audio_encoder.encoders.{i}", D_ENC))
    for i in range(N_TP):
        spec.update(_sanm_layer(f"audio_encoder.tp_encoders.{i}", D_ENC))
    for n in ("after_norm", "tp_norm"):
        spec[f"audio_encoder.{n}.weight"] = (D_ENC,)
        spec[f"audio_encoder.{n}.bias"] = (D_ENC,)
    spec.update(_projector("audio_adaptor", D_LLM, ADAPTOR_BLOCKS))
    spec.update(_projector("ctc_decoder", D_ENC, CTC_BLOCKS))
    spec["ctc_proj.ctc_lo.weight"] = (vocab, D_ENC)
    spec["ctc_proj.ctc_lo.bias"] = (vocab,)
    return spec


# --------------------------------------------------------------------------------------
# Seeded initialisation
# --------------------------------------------------------------------------------------

def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFFFFFFFFFF)
    return g


def _uniform(shape, bound: float, g: torch.Generator) -> torch.Tensor:
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def random_weights(seed: int = 0, vocab: int = VOCAB) -> Dict[str, torch.Tensor]:
    """Random-init weights of the named architecture.

    Linear / conv tensors follow torch's default bound 1/sqrt(fan_in); LayerNorm affine
    parameters are jittered around (1, 0) so that a kernel ignoring them is caught.
    """
    out: Dict[str, torch.Tensor] = {}
    for name, shape in tensor_spec(vocab).items():
        g = _gen(seed, name)
        leaf = name.rsplit(".", 1)[-1]
        is_norm = ".norm" in name or name.endswith("_norm.weight") or name.endswith("_norm.bias")
        if is_norm:
            t = _uniform(shape, 0.1, g) + (1.0 if leaf == "weight" else 0.0)
        elif "fsmn_block" in name:
            t = _uniform(shape, 1.0 / math.sqrt(FSMN_K), g)
        else:
            w_shape = shape if leaf == "weight" else tensor_spec_cache(vocab)[name[: -len("bias")] + "weight"]
            t = _uniform(shape, 1.0 / math.sqrt(w_shape[1]), g)
        out[name] = t.contiguous()
    return out


_SPEC_CACHE: Dict[int, "OrderedDict[str, Tuple[int, ...]]"] = {}


def tensor_spec_cache(vocab: int) -> "OrderedDict[str, Tuple[int, ...]]":
    if vocab not in _SPEC_CACHE:
        _SPEC_CACHE[vocab] = tensor_spec(vocab)
    return _SPEC_CACHE[vocab]


def load_checkpoint(path: str, vocab: int = VOCAB) -> Dict[str, torch.Tensor]:
    """Load a FunASR ``model.pt`` with the reference's key mapping (model_definition.py:231-238)."""
    sd = torch.load(path, map_location="cpu")
    if "state_dict" in sd:
        sd = sd["state_dict"]
    spec = tensor_spec_cache(vocab)
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.startswith(("audio_encoder.", "audio_adaptor.", "ctc_decoder.")):
            nk = k
        elif k.startswith("ctc.ctc_lo."):
            nk = k.replace("ctc.ctc_lo", "ctc_proj.ctc_lo")
        else:
            continue
        if nk in spec:
            out[nk] = v.to(torch.float32).contiguous()
    missing = [k for k in spec if k not in out]
    if missing:
        raise KeyError(f"checkpoint {path} lacks {len(missing)} tensors, first: {missing[:3]}")
    for k, shape in spec.items():
        if tuple(out[k].shape) != tuple(shape):
            raise ValueError(f"{k}: checkpoint shape {tuple(out[k].shape)} != expected {shape}")
    return out


# --------------------------------------------------------------------------------------
# Front-end constants
# --------------------------------------------------------------------------------------

def dft_kernels() -> Tuple[torch.Tensor, torch.Tensor]:
    """Windowed cos / -sin DFT kernels, each (201, 400) fp32.

    Same fp32 expression order as STFT_Process.__init__ (model_definition.py:247-253): the
    phase 2*pi*f*t/400 is rounded to fp32 before cos/sin, so the tables are *not* exact
    twiddles — which is why the CUDA path multiplies by these tables instead of running an FFT.
    """
    window = torch.hamming_window(N_FFT, periodic=True)
    t = torch.arange(N_FFT).unsqueeze(0)
    f = torch.arange(N_BINS).unsqueeze(1)
    omega = 2 * torch.pi * f * t / N_FFT
    cos_k = (torch.cos(omega) * window.unsqueeze(0)).to(torch.float32).contiguous()
    sin_k = (-torch.sin(omega) * window.unsqueeze(0)).to(torch.float32).contiguous()
    return cos_k, sin_k


def _hz_to_mel_htk(f: float) -> float:
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank() -> torch.Tensor:
    """(80, 201) HTK triangular filterbank, unnormalised, 20..8000 Hz over 0..8000 Hz bins.

    Restates torchaudio.functional.melscale_fbanks(201, 20, 8000, 80, 16000, None, 'htk')
    as called at 01-Export-Encoder-Adaptor-CTC.py:102 (then transposed to mel-major).
    """
    all_freqs = torch.linspace(0, SAMPLE_RATE // 2, N_BINS)
    m_min, m_max = _hz_to_mel_htk(20.0), _hz_to_mel_htk(float(SAMPLE_RATE // 2))
    m_pts = torch.linspace(m_min, m_max, N_MELS + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)          # (201, 82)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))           # (201, 80)
    return fb.transpose(0, 1).to(torch.float32).contiguous()


def position_table(n_frames: int, depth: int = D_IN) -> torch.Tensor:
    """(n_frames, depth) sinusoidal table for positions 1..n_frames (model_definition.py:13-28)."""
    dtype = torch.float32
    positions = torch.arange(1, n_frames + 1, dtype=torch.long).unsqueeze(0).type(dtype)
    inc = torch.log(torch.tensor([10000], dtype=dtype)) / (depth / 2 - 1)
    inv = torch.exp(torch.arange(depth / 2).type(dtype) * (-inc)).unsqueeze(0)
    scaled = positions.unsqueeze(-1) * inv.unsqueeze(1)
    return torch.cat([torch.sin(scaled), torch.cos(scaled)], dim=2)[0].to(dtype).contiguous()


def front_end_constants(max_frames: int) -> Dict[str, torch.Tensor]:
    cos_k, sin_k = dft_kernels()
    return {
        "const.dft_cos": cos_k,
        "const.dft_sin": sin_k,
        "const.mel_fbank": mel_filterbank(),
        "const.pos_enc": position_table(max_frames),
    }
