"""Where a projection's epilogue spends its cycles (tuning aid).  Needs the timing build of the library:

    cd fun_asr_gguf_b200/csrc && for f in *.cu; do nvcc <build flags> -DFA_GEMM_TIMING -c $f -o /tmp/ab_${f%.cu}.o; done
    nvcc -shared -o tools/_ab/libfunasr_timing.so /tmp/ab_*.o
    FUNASR_B200_LIB=tools/_ab/libfunasr_timing.so FUNASR_B200_GEMM_TIMING=1 python tools/gemm_timing.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.gpu_util import RawContext  # noqa: E402

M = int(os.environ.get("M", 32032))
raw = RawContext()
rng = np.random.default_rng(0)
for prec in ("bf16x3", "bf16", "fp8"):
    for n, k, resid, planes in ((512, 512, True, False), (512, 2048, True, False), (2048, 512, False, True), (1536, 512, False, True)):
        a = rng.standard_normal((M, k), dtype=np.float32)
        w = rng.standard_normal((n, k), dtype=np.float32) * k ** -0.5
        b = rng.standard_normal((n,), dtype=np.float32)
        r = rng.standard_normal((M, n), dtype=np.float32) if resid else None
        if planes:
            os.environ["FUNASR_B200_TEST_NO_F32"] = "1"
        else:
            os.environ.pop("FUNASR_B200_TEST_NO_F32", None)
        sys.stderr.write(f"--- {prec} n={n} k={k} resid={resid} planes_only={planes}\n")
        sys.stderr.flush()
        raw.linear(a, w, b, resid=r, relu=planes, precision=prec, planes=planes)
raw.close()
