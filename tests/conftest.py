import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the C-ABI library is a build artefact (git-ignored): make sure it exists and matches the sources before any
    # test dlopens it.  A no-op when the stamp is current; nvcc cross-compiles sm_100a without a GPU.
    from fun_asr_gguf_b200 import build as _build
    _build.build()


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(d, "reference_outputs.json")) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(d, "reference_outputs.npz"))


@pytest.fixture(scope="session")
def weights():
    from fun_asr_gguf_b200 import weights as Wm
    return Wm.random_weights(0)


@pytest.fixture(scope="session")
def consts():
    from fun_asr_gguf_b200 import weights as Wm
    return Wm.front_end_constants(1100)


@pytest.fixture(scope="session")
def planted_weights(weights, consts):
    from tests import planted
    return planted.plant(weights, consts)


@pytest.fixture(scope="session", autouse=True)
def _threads():
    torch.set_num_threads(os.cpu_count() or 1)
