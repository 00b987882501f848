"""Pins oracle/style_oracle.py (the CPU restatement of the style encoder) against the committed
fixtures that oracle/make_golden_style.py generated from the reference's own StyleEncoder class (simbert.py:33-72)
driving transformers' BertModel.  Tolerance 1e-5 of max|ref|: two fp32 evaluation orders of a 12-layer network (the library
uses a fused SDPA); the fixtures' meta file records both at ~1e-6 of an fp64 run."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, rel_max
from emotivoice_b200 import synth
from oracle import style_oracle as SO

KEYS = ("input_ids", "token_type_ids", "attention_mask")
OUTS = ("pooled_output", "pitch_outputs", "speed_outputs", "energy_outputs", "emotion_outputs")


@pytest.fixture(scope="module")
def style_meta():
    with open(os.path.join(GOLDEN, "style_meta.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name,small", [("style_small_b3", True), ("style_small_b1_n40", True), ("style_base_b2", False)])
def test_style_oracle_matches_reference_fixture(style_meta, name, small):
    sc = synth.style_config(small)
    sd = synth.make_style_state_dict(sc)
    assert synth.state_dict_digest(sd) == style_meta["state_dict_digest_" + ("small" if small else "base")]
    g = load_golden(name)
    out = SO.style_forward(sd, sc.num_attention_heads, **{k: g[k] for k in KEYS})
    for k in OUTS:
        assert out[k].shape == g[k].shape
        assert rel_max(out[k], g[k]) <= 1e-5, k
    assert 0.05 < g["pooled_output"].abs().mean() < 0.95          # a non-degenerate fixture (tanh neither dead nor saturated)


def test_style_inputs_are_reproducible_and_padding_is_ignored(style_meta):
    """The fixture inputs come from synth.make_style_batch; a padded item equals the same item alone (key mask)."""
    sc = synth.style_config(True)
    sd = synth.make_style_state_dict(sc)
    case = style_meta["cases"]["style_small_b3"]
    batch = synth.make_style_batch(sc, case["lengths"], case["seed"])
    g = load_golden("style_small_b3")
    for k in KEYS:
        assert torch.equal(batch[k], g[k])
    n = case["lengths"][0]
    alone = SO.style_forward(sd, sc.num_attention_heads, **{k: g[k][:1, :n] for k in KEYS})
    assert rel_max(alone["pooled_output"], g["pooled_output"][:1]) <= 1e-5
