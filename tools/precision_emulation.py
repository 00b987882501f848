"""CPU emulation of the reduced-precision modes' error budget (no GPU needed).

Rounds the operands of every Linear / Conv1d / ConvTranspose1d from the decoder stack onward (decoder, to_mel,
vocoder -- the part `precision = "tf32" | "bf16"` touches; the encoder/predictor prefix stays fp32 in every mode) to
the mode's operand format and keeps fp32 accumulation, using the oracle as the carrier.  The result is the error the
tolerance in tests/test_tc_gpu.py should be a small multiple of.  Test infrastructure: uses oracle/, never the engine.

    python tools/precision_emulation.py > profiles/r01_precision_emulation_cpu.json
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emotivoice_b200 import packing, synth                      # noqa: E402
from emotivoice_b200.config import default_config               # noqa: E402
from oracle import jets_oracle as O                              # noqa: E402

KEYS = ("inputs_ling", "input_lengths", "inputs_speaker", "inputs_style_embedding", "inputs_content_embedding")


def emulate(sd, conf, rnd, name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    batch = {k: torch.from_numpy(z[k]) for k in KEYS}
    lin, conv, convt, stack = F.linear, F.conv1d, F.conv_transpose1d, O.encoder_stack

    def patched_stack(sd_, prefix, *a, **kw):
        if prefix == "am.decoder":                                # from here on every GEMM-shaped op is rounded
            O.F.linear = lambda x, w, b=None: lin(rnd(x), rnd(w), b)
            O.F.conv1d = lambda x, w, b=None, **k: conv(rnd(x), rnd(w), b, **k)
            O.F.conv_transpose1d = lambda x, w, b=None, **k: convt(rnd(x), rnd(w), b, **k)
        return stack(sd_, prefix, *a, **kw)

    O.encoder_stack = patched_stack
    try:
        out = O.jets_forward(sd, conf, **batch)
    finally:
        O.F.linear, O.F.conv1d, O.F.conv_transpose1d, O.encoder_stack = lin, conv, convt, stack
    mel, wav = torch.from_numpy(z["mel"]), torch.from_numpy(z["wav"])
    return {
        "durations_identical": bool(torch.equal(out["log_duration_predictions"], torch.from_numpy(z["durations"]))),
        "mel_rel_max": float((out["dec_outputs"] - mel).abs().max() / mel.abs().max()),
        "wav_rms_rel": float(((out["wav_predictions"] - wav).double().pow(2).mean().sqrt()
                              / wav.double().pow(2).mean().sqrt())),
    }


def main():
    conf = default_config()
    sd = synth.make_state_dict(conf)
    modes = {"tf32": packing.round_tf32, "bf16": lambda t: t.to(torch.bfloat16).float()}
    res = {"what": "operand rounding from the decoder on, fp32 accumulation, CPU oracle as carrier",
           "cases": {n: {m: emulate(sd, conf, f, n) for m, f in modes.items()} for n in ("b1_t12", "b1_t100")}}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
