"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/style_*.npz from the reference's own ``StyleEncoder`` class
(simbert.py:33-72) and pins oracle/style_oracle.py against it.

Run in the build container (needs /root/reference and the ``transformers`` package):

    python oracle/make_golden_style.py

The reference builds its BERT with ``AutoModel.from_pretrained('WangZeJun/simbert-base-chinese')``; the pretrained files
are not available offline, so ``from_pretrained`` is replaced by a constructor of the same architecture
(``BertModel(BertConfig(...))``) and the seeded synthetic state dict of ``emotivoice_b200.synth.make_style_state_dict`` is
loaded into the reference module (strict).  Everything downstream -- StyleEncoder.forward and transformers' BertModel --
is the unmodified code the reference runs, called like inference_am_vocoder_joint.py:25-38 calls it.
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from emotivoice_b200 import synth                           # noqa: E402
from oracle import refshim                                   # noqa: E402
from oracle import style_oracle as SO                        # noqa: E402

CASES = {
    # name: (small?, token counts, input seed)
    "style_small_b3": (True, [7, 16, 11], 5001),      # ragged batch: key-padding mask
    "style_small_b1_n40": (True, [40], 5002),
    "style_base_b2": (False, [12, 30], 5003),         # BERT-base dimensions (simbert-base-chinese), short prompts like the callers'
}


def reference_style_encoder(sc):
    import transformers
    from transformers import BertConfig, BertModel
    if refshim.REF_ROOT not in sys.path:
        sys.path.insert(0, refshim.REF_ROOT)
    cfg = BertConfig(vocab_size=sc.vocab_size, hidden_size=sc.hidden_size, num_hidden_layers=sc.num_hidden_layers,
                     num_attention_heads=sc.num_attention_heads, intermediate_size=sc.intermediate_size,
                     max_position_embeddings=sc.max_position_embeddings, type_vocab_size=sc.type_vocab_size)
    orig = transformers.AutoModel.from_pretrained
    transformers.AutoModel.from_pretrained = staticmethod(lambda *a, **k: BertModel(cfg))
    try:
        from models.prompt_tts_modified.simbert import StyleEncoder
        ref_conf = types.SimpleNamespace(bert_path="(offline)", bert_hidden_size=sc.hidden_size, style_dim=sc.style_dim,
                                         pitch_n_labels=sc.pitch_n_labels, speed_n_labels=sc.speed_n_labels,
                                         energy_n_labels=sc.energy_n_labels, emotion_n_labels=sc.emotion_n_labels)
        return StyleEncoder(ref_conf).eval()
    finally:
        transformers.AutoModel.from_pretrained = orig


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    meta = {"torch": torch.__version__, "cases": {}}
    import transformers
    meta["transformers"] = transformers.__version__
    models = {}
    for name, (small, lens, seed) in CASES.items():
        sc = synth.style_config(small)
        if small not in models:
            sd = synth.make_style_state_dict(sc)
            ref = reference_style_encoder(sc)
            missing, unexpected = ref.load_state_dict(sd, strict=False)
            assert not unexpected and all("position_ids" in m for m in missing), (missing, unexpected)
            models[small] = (sd, ref)
            meta["state_dict_digest_" + ("small" if small else "base")] = synth.state_dict_digest(sd)
            if small:      # the reference module's own state-dict contract (names + shapes) for the host-mirror test
                meta["reference_state_dict_small"] = {k: list(v.shape) for k, v in ref.state_dict().items()}
        sd, ref = models[small]
        batch = synth.make_style_batch(sc, lens, seed)
        with torch.no_grad():
            r = ref(**batch)
        o = SO.style_forward(sd, sc.num_attention_heads, **batch)
        o64 = SO.style_forward(sd, sc.num_attention_heads, dtype=torch.float64, **batch)
        errs = {}
        for k in r:
            errs[k] = (r[k] - o[k]).abs().max().item() / max(r[k].abs().max().item(), 1e-30)
            assert errs[k] <= 1e-5, (name, k, errs[k])     # two fp32 evaluation orders (the library uses fused SDPA); both sit ~1e-6 from fp64
        err64 = (r["pooled_output"].double() - o64["pooled_output"]).abs().max().item()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"),
                            **{k: v.numpy() for k, v in batch.items()}, **{k: v.numpy() for k, v in r.items()})
        meta["cases"][name] = dict(small=small, lengths=lens, seed=seed, oracle_rel_err=errs, reference_abs_err_vs_fp64=err64)
        print(name, errs, "ref vs fp64 oracle: %.2e" % err64)
    with open(os.path.join(out_dir, "style_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
