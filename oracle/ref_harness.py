"""ORACLE support — test infrastructure only.

Runs the reference's *own* modules (``/root/reference/fun_asr_gguf/model_definition.py``,
loaded by file path because importing the package needs onnxruntime and libllama — SURVEY F3)
built exactly as the export recipe builds them (01-Export-Encoder-Adaptor-CTC.py:97-107,127),
in PyTorch eager FP32 on CPU.  Only usable where /root/reference is mounted (the build
container); the GPU box never has it.  Used by tests/golden/make_golden.py to produce the
committed pins and by tests/test_oracle_golden.py::test_oracle_vs_live_reference.
"""
from __future__ import annotations

import importlib.util
import os
from typing import Dict, Tuple

import torch

REF_ROOT = os.environ.get("FUNASR_REFERENCE_ROOT", "/root/reference")
REF_MODEL_DEF = os.path.join(REF_ROOT, "fun_asr_gguf", "model_definition.py")


def available() -> bool:
    return os.path.isfile(REF_MODEL_DEF)


def _load_module(path: str = REF_MODEL_DEF):
    spec = importlib.util.spec_from_file_location("_ref_model_definition", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Reference:
    """The two export wrappers of the reference, holding the given weights."""

    def __init__(self, weights: Dict[str, torch.Tensor], vocab: int = 60515, module_path: str = REF_MODEL_DEF):
        import torchaudio

        md = _load_module(module_path)
        hybrid = md.HybridSenseVoice(vocab_size=vocab)
        missing, unexpected = hybrid.load_state_dict(weights, strict=True), None
        hybrid.eval()
        stft = md.STFT_Process(n_fft=400, win_length=400, hop_len=160).eval()
        fbank = torchaudio.functional.melscale_fbanks(201, 20, 8000, 80, 16000, None, "htk").transpose(0, 1).unsqueeze(0)
        self.fbank, self.stft = fbank, stft
        self.enc = md.EncoderExportWrapperPaddable(hybrid, stft, fbank).eval()
        self.ctc = md.CTCHeadExportWrapper(hybrid).eval()
        self.hybrid = hybrid

    @torch.no_grad()
    def encode(self, audio: torch.Tensor, n_valid: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """audio (S_phys,) -> enc (T,512), adaptor_output (T,1024); batch 1 as exported."""
        enc, ad = self.enc(audio.view(1, 1, -1).float(), torch.tensor([int(n_valid)], dtype=torch.long))
        return enc[0], ad[0]

    @torch.no_grad()
    def ctc_ids(self, enc: torch.Tensor) -> torch.Tensor:
        return self.ctc(enc.unsqueeze(0))[0]

    @torch.no_grad()
    def ctc_logits(self, enc: torch.Tensor) -> torch.Tensor:
        h, _ = self.ctc.ctc_decoder(enc.unsqueeze(0), None)
        return self.ctc.ctc_proj.ctc_lo(h)[0]
