"""One cfg3 batch (32 mixed EN/ZH utterances, 20-200 phonemes) forward, repeated (for launch lists / timing):
   python tools/quick_b32.py [precision] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emotivoice_b200 import synth
from emotivoice_b200.config import default_config
from emotivoice_b200.modules import JETSGenerator
conf = default_config(); dev = torch.device("cuda:0")
m = JETSGenerator(conf).to(dev); m.load_state_dict(synth.make_state_dict(conf)); m.eval()
m.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
utts = sorted([synth.corpus_utterance(10_000 + i) for i in range(32)], key=lambda u: -len(u["ids"]))
b = {k: v.to(dev) for k, v in synth.collate_utterances(utts).items()}
for i in range(2):
    o = m(**b); torch.cuda.synchronize()
ts = []
for i in range(reps):
    t0 = time.perf_counter(); o = m(**b); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("%s B=32: median %.2f ms, valid frames %d, padded %d" % (m.precision, 1e3 * sorted(ts)[len(ts) // 2], int(o["mel_lengths_host"].sum()), 32 * o["dec_outputs"].shape[1]), flush=True)
