"""Latency of the style encoder on the GPU engine vs. the CPU path the reference uses.

    python tools/style_bench.py [--tokens 24] [--batch 1] [--steps 50] > profiles/rNN_style_encoder.json

GPU arm: emotivoice_b200.style.StyleEncoder (BERT-base dimensions, seeded weights), CUDA events around `steps` forwards of
one tokenised prompt, inputs resident.  CPU arm: transformers' BertModel (the library the reference calls, simbert.py:37)
with the same weights on the host cores -- what every reference caller pays twice per utterance
(inference_am_vocoder_joint.py:106-107).  Also prints the max deviation between the two pooled outputs.
Not part of bench.py's contract (that stays on JETSGenerator.forward); run it under gpurun.
"""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emotivoice_b200 import synth                                  # noqa: E402
from emotivoice_b200.style import StyleEncoder                     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=24)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32"])
    ap.add_argument("--cpu-threads", type=int, default=min(16, len(os.sched_getaffinity(0))))
    a = ap.parse_args()
    sc = synth.style_config(False)
    sd = synth.make_style_state_dict(sc)
    conf = SimpleNamespace(bert_path="(offline)", bert_hidden_size=sc.hidden_size, style_dim=sc.style_dim,
                           pitch_n_labels=sc.pitch_n_labels, speed_n_labels=sc.speed_n_labels,
                           energy_n_labels=sc.energy_n_labels, emotion_n_labels=sc.emotion_n_labels)
    batch = synth.make_style_batch(sc, [a.tokens] * a.batch)
    dev = torch.device("cuda:0")
    m = StyleEncoder(conf, bert_config=dict(sc), _init=sd).to(dev).eval()
    m.precision = a.precision
    gb = {k: v.to(dev) for k, v in batch.items()}
    for _ in range(5):
        out = m(**gb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = m(**gb)
    e1.record()
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / a.steps
    res = dict(workload="style encoder, BERT-base 12x768, B=%d, %d tokens" % (a.batch, a.tokens), precision=a.precision,
               gpu_ms_per_forward=gpu_ms, steps=a.steps)
    try:
        from transformers import BertConfig, BertModel
        torch.set_num_threads(a.cpu_threads)
        cfg = BertConfig(vocab_size=sc.vocab_size, hidden_size=sc.hidden_size, num_hidden_layers=sc.num_hidden_layers,
                         num_attention_heads=sc.num_attention_heads, intermediate_size=sc.intermediate_size,
                         max_position_embeddings=sc.max_position_embeddings, type_vocab_size=sc.type_vocab_size)
        ref = BertModel(cfg).eval()
        ref.load_state_dict({k[len("bert."):]: v for k, v in sd.items() if k.startswith("bert.")}, strict=False)
        with torch.no_grad():
            ref(**batch)
            t = time.perf_counter()
            n = max(3, a.steps // 10)
            for _ in range(n):
                r = ref(**batch)
            cpu_ms = (time.perf_counter() - t) * 1e3 / n
        res.update(cpu_ms_per_forward=cpu_ms, cpu_threads=a.cpu_threads, cpu_kind="transformers BertModel (the reference's library)",
                   speedup=cpu_ms / gpu_ms,
                   pooled_max_abs_diff=float((out["pooled_output"].cpu() - r["pooler_output"]).abs().max()))
    except ImportError as e:
        res["cpu_arm_unavailable"] = str(e)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
