"""Why the duration-critical prefix is fp32-accurate in EVERY precision mode: emulate, on the CPU
oracle, a prefix whose linear / conv operands are rounded to tf32 (one tf32 MMA per K step) or to bf16, and count how
many utterances get a different duration vector (=> a different number of frames, every later frame shifted) than
the fp32 reference.  Test infrastructure (uses oracle/); prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from emotivoice_b200 import synth, packing
from emotivoice_b200.config import default_config
from oracle import jets_oracle as O

conf = default_config()
sd = synth.make_state_dict(conf)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40


def run(round_fn):
    lin, conv = F.linear, F.conv1d
    if round_fn is not None:
        O.F.linear = lambda x, w, b=None: lin(round_fn(x), round_fn(w), b)
        O.F.conv1d = lambda x, w, b=None, **k: conv(round_fn(x), round_fn(w), b, **k)
    try:
        outs = []
        for i in range(N):
            bt = synth.make_batch([100], seed=5000 + i)
            r = O.acoustic_model(sd, conf, bt["inputs_ling"], bt["input_lengths"], bt["inputs_speaker"],
                                 bt["inputs_style_embedding"], bt["inputs_content_embedding"])
            outs.append(r["log_duration_predictions"][0].clone())
        return outs
    finally:
        O.F.linear, O.F.conv1d = lin, conv


ref = run(None)
res = {"utterances": N, "phonemes_each": 100}
for name, fn in (("tf32_1x", packing.round_tf32), ("bf16", lambda t: t.to(torch.bfloat16).float())):
    got = run(fn)
    diff_utts = sum(int(not torch.equal(a, b)) for a, b in zip(ref, got))
    diff_tokens = sum(int((a != b).sum()) for a, b in zip(ref, got))
    diff_frames = sum(int(a.sum() != b.sum()) for a, b in zip(ref, got))
    res[name] = {"utterances_with_a_changed_duration": diff_utts, "changed_tokens": diff_tokens, "utterances_with_changed_length": diff_frames}
print(json.dumps(res))
