"""Generate tests/golden/*.npz by running the REFERENCE's own modules (build container only).

    python tests/golden/make_golden.py            # the parity cases (reference_outputs.*)
    python tests/golden/make_golden.py --bench32  # the 32 x 60 s batch bench.py measures on (bench32_outputs.*)

Needs /root/reference (read-only mount).  The reference's model_definition.py is executed in
PyTorch eager FP32 with the seeded weights of fun_asr_gguf_b200.weights.random_weights(0)
(and their planted-CTC variant), built as 01-Export-Encoder-Adaptor-CTC.py:97-107,127 builds
them.  The outputs are the pins the oracle is checked against (tests/test_oracle_golden.py).
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from fun_asr_gguf_b200 import weights as Wm  # noqa: E402
from oracle import ref_harness  # noqa: E402
from tests import cases, planted  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()[:16]


def main():
    torch.set_num_threads(os.cpu_count())
    w = Wm.random_weights(0)
    consts = Wm.front_end_constants(1100)
    wp = planted.plant(w, consts)
    ref = ref_harness.Reference(w)
    ref_p = ref_harness.Reference(wp)
    meta = {
        "oracle": "reference model_definition.py, PyTorch eager FP32, CPU (not ONNX Runtime)",
        "torch": torch.__version__,
        "weights_seed": 0,
        "weight_sha": {k: sha(w[k]) for k in (
            "audio_encoder.encoders0.0.self_attn.linear_q_k_v.weight",
            "audio_encoder.tp_encoders.19.feed_forward.w_2.weight",
            "audio_adaptor.blocks.1.norm2.bias", "ctc_proj.ctc_lo.weight")},
        "planted_sha": {k: sha(wp[k]) for k in ("ctc_proj.ctc_lo.weight", "ctc_proj.ctc_lo.bias")},
        "const_sha": {k: sha(v) for k, v in consts.items()},
        "cases": {},
    }
    blobs = {}
    for name in cases.CASES:
        audio, n_valid = cases.build(name)
        enc, ad = ref.encode(audio, n_valid)
        ids = ref.ctc_ids(enc)
        ids_p = ref_p.ctc_ids(enc)
        lg = ref.ctc_logits(enc).topk(2, -1).values
        tl = Wm.adaptor_target_len(n_valid)
        assert float(ad[tl:].abs().max() if ad.shape[0] > tl else 0.0) == 0.0
        blobs[f"{name}.audio"] = audio.numpy()
        blobs[f"{name}.enc"] = enc.numpy()
        blobs[f"{name}.adaptor"] = ad[:tl].numpy()
        blobs[f"{name}.ids"] = ids.numpy()
        blobs[f"{name}.ids_planted"] = ids_p.numpy()
        blobs[f"{name}.margin"] = (lg[:, 0] - lg[:, 1]).numpy()
        meta["cases"][name] = {"n_valid": n_valid, "n_phys": int(audio.shape[0]), "frames": int(enc.shape[0]),
                               "target_len": tl, "audio_sha": sha(audio)}
        print(name, enc.shape, tl, len(ids.unique()), len(ids_p.unique()))
    # 60 s case: ids in full, activations subsampled
    name, fn, n_valid, n_phys = cases.SIXTY
    audio = fn()
    enc, ad = ref.encode(audio, n_valid)
    tl = Wm.adaptor_target_len(n_valid)
    lg = ref_p.ctc_logits(enc).topk(2, -1).values
    blobs[f"{name}.enc_rows"] = enc[::25].numpy()
    blobs[f"{name}.adaptor_rows"] = ad[:tl][::6].numpy()
    blobs[f"{name}.ids"] = ref.ctc_ids(enc).numpy()
    blobs[f"{name}.ids_planted"] = ref_p.ctc_ids(enc).numpy()
    blobs[f"{name}.margin_planted"] = (lg[:, 0] - lg[:, 1]).numpy()
    meta["cases"][name] = {"n_valid": n_valid, "n_phys": n_phys, "frames": int(enc.shape[0]), "target_len": tl,
                           "audio_sha": sha(audio), "enc_sha": sha(enc), "enc_row_stride": 25, "adaptor_row_stride": 6}
    print(name, enc.shape, tl, len(np.unique(blobs[f"{name}.ids"])), len(np.unique(blobs[f"{name}.ids_planted"])))
    np.savez_compressed(os.path.join(OUT, "reference_outputs.npz"), **blobs)
    with open(os.path.join(OUT, "reference_outputs.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


def bench32():
    """The batch bench.py's default workload times (BASELINE configs[1]): 32 x 60 s of white noise, seeds 1234 + i.
    Pins per row: ids (random-init and planted CTC projection), the top-2 margins, and every 200th enc row."""
    torch.set_num_threads(os.cpu_count())
    from fun_asr_gguf_b200 import synth
    w = Wm.random_weights(0)
    consts = Wm.front_end_constants(1100)
    wp = planted.plant(w, consts)
    # the structured-audio plant leaves white noise on one id; a plant calibrated on white noise makes the ids of the
    # benchmark's own signals change almost every frame
    wpw = planted.plant(w, consts, cal=synth.white(6 * cases.SR, 9999))
    ref, ref_p, ref_pw = ref_harness.Reference(w), ref_harness.Reference(wp), ref_harness.Reference(wpw)
    n = 60 * cases.SR
    blobs = {k: [] for k in ("ids", "ids_planted", "ids_planted_white", "margin", "margin_planted", "margin_planted_white",
                             "enc_rows", "adaptor_rows")}
    for i in range(32):
        audio = synth.white(n, i)
        enc, ad = ref.encode(audio, n)
        lg, lgp, lgw = ref.ctc_logits(enc), ref_p.ctc_logits(enc), ref_pw.ctc_logits(enc)
        t2, t2p, t2w = lg.topk(2, -1).values, lgp.topk(2, -1).values, lgw.topk(2, -1).values
        blobs["ids_planted_white"].append(lgw.argmax(-1).to(torch.int32).numpy())
        blobs["margin_planted_white"].append((t2w[:, 0] - t2w[:, 1]).numpy())
        blobs["ids"].append(lg.argmax(-1).to(torch.int32).numpy())
        blobs["ids_planted"].append(lgp.argmax(-1).to(torch.int32).numpy())
        blobs["margin"].append((t2[:, 0] - t2[:, 1]).numpy())
        blobs["margin_planted"].append((t2p[:, 0] - t2p[:, 1]).numpy())
        blobs["enc_rows"].append(enc[::200].numpy())
        blobs["adaptor_rows"].append(ad[:126:25].numpy())
        assert np.array_equal(blobs["ids"][-1], ref.ctc_ids(enc).numpy())
        print(i, len(np.unique(blobs["ids"][-1])), len(np.unique(blobs["ids_planted"][-1])), len(np.unique(blobs["ids_planted_white"][-1])),
              float(blobs["margin_planted_white"][-1].min()), flush=True)
    np.savez_compressed(os.path.join(OUT, "bench32_outputs.npz"), **{k: np.stack(v) for k, v in blobs.items()})
    with open(os.path.join(OUT, "bench32_outputs.json"), "w") as f:
        json.dump({"oracle": "reference model_definition.py, PyTorch eager FP32, CPU (not ONNX Runtime)", "torch": torch.__version__,
                   "signal": "fun_asr_gguf_b200.synth.white(960000, i), i = 0..31 (bench.py rank 0, first input set)",
                   "enc_row_stride": 200, "adaptor_row_stride": 25, "weights_seed": 0}, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    if "--bench32" in sys.argv:
        bench32()
    else:
        main()
