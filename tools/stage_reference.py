"""Stage the reference's own host-side call sites where the GPU box can see them (test infrastructure).

    python tools/stage_reference.py

Copies the Python files of /root/reference/fun_asr_gguf that sit on either side of the replaced sessions —
nano_onnx.py, nano_ctc.py, core/decoder.py (CTCDecoder) and the small modules they import — UNMODIFIED into
baseline/_ref/fun_asr_gguf/.  baseline/_ref/ is git-ignored (the reference's sources never enter this repo's
history) but not gpurun-ignored, so the files travel with the snapshot and tests/test_gpu_dropin.py can run the
reference's byte-for-byte call sites over the CUDA engine.  /root/reference itself does not exist on the GPU box.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("FUNASR_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref", "fun_asr_gguf")
FILES = ["nano_onnx.py", "nano_ctc.py", "nano_dataclass.py", "nano_audio.py", "utils.py", "display.py", "text_merge.py",
         "core/decoder.py", "core/orchestrator.py"]


def stage() -> bool:
    src = os.path.join(SRC, "fun_asr_gguf")
    if not os.path.isdir(src):
        return False
    for rel in FILES:
        d = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), d)
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged into " + DST if ok else f"{SRC} is not mounted: nothing staged")
    sys.exit(0 if ok else 1)
