"""The seeded test signals live with the package (bench.py measures on the same ones): fun_asr_gguf_b200/synth.py."""
from fun_asr_gguf_b200.synth import padded, structured, white  # noqa: F401
