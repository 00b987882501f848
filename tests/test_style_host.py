"""Host logic of the style encoder, no GPU: the mirror's state-dict contract against the reference
module's own, and the packed engine layout ("sty.*") against the oracle through a CPU walk of the exact launch sequence
of csrc/style_engine.cu (fused q|k|v GEMM, residual epilogues, [CLS] GEMV, heads side by side)."""
import json
import math
import os
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, load_golden, rel_max
from emotivoice_b200 import packing, synth

KEYS = ("input_ids", "token_type_ids", "attention_mask")


def _ref_config(sc):
    return SimpleNamespace(bert_path="(offline)", bert_hidden_size=sc.hidden_size, style_dim=sc.style_dim,
                           pitch_n_labels=sc.pitch_n_labels, speed_n_labels=sc.speed_n_labels,
                           energy_n_labels=sc.energy_n_labels, emotion_n_labels=sc.emotion_n_labels)


def from_tc_layout(w_tc, cin, cout):
    """inverse of packing.to_tc_layout for K = 1: (2, NT, 1, Cin/4, BNp, 4) -> (Cin, Cout) hi + lo."""
    bnp = min(cout, packing.TC_BN)
    g = (w_tc[0] + w_tc[1]).reshape(cout // bnp, 1, cin // 4, bnp, 4)
    return g.permute(1, 2, 4, 0, 3).reshape(cin, cout)


def test_mirror_has_the_reference_state_dict_contract():
    from emotivoice_b200.style import StyleEncoder
    with open(os.path.join(GOLDEN, "style_meta.json")) as f:
        want = json.load(f)["reference_state_dict_small"]
    sc = synth.style_config(True)
    m = StyleEncoder(_ref_config(sc), bert_config=dict(sc))
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    sd = synth.make_style_state_dict(sc)
    legacy = dict(sd, **{"bert.embeddings.position_ids": torch.arange(sc.max_position_embeddings)[None]})
    m.load_state_dict(legacy, strict=True)                  # checkpoints of transformers < 4.31 carry the buffer
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())
    with pytest.raises(RuntimeError, match="CUDA"):         # no CPU path
        m(*(torch.ones(1, 4, dtype=torch.long),) * 3)
    with pytest.raises(ValueError):
        StyleEncoder(SimpleNamespace(bert_hidden_size=512), bert_config=dict(sc))


def test_full_size_defaults_are_bert_base():
    sc = synth.style_config(False)
    shapes = dict((n, s) for n, s, _ in synth.style_param_shapes(sc))
    assert shapes["bert.embeddings.word_embeddings.weight"] == (13685, 768)
    assert shapes["bert.encoder.layer.11.intermediate.dense.weight"] == (3072, 768)
    assert sum(math.prod(s) for s in shapes.values()) > 95e6
    slices, width = packing.style_head_slices(sc)
    assert width % 8 == 0 and slices["emotion"] == (9, 7)


@pytest.mark.parametrize("name", ["style_small_b3", "style_small_b1_n40"])
def test_packed_layout_walked_like_the_engine_reproduces_the_reference(name):
    sc = synth.style_config(True)
    sd = synth.make_style_state_dict(sc)
    p = packing.pack_style_state_dict(sd, sc)
    assert all(len(k) < 56 for k in p)
    g = load_golden(name)
    ids, tts, mask = (g[k] for k in KEYS)
    B, N = ids.shape
    H, I, nh = sc.hidden_size, sc.intermediate_size, sc.num_attention_heads
    dk = H // nh
    lens = mask.sum(1)
    ln = lambda x, pre: F.layer_norm(x, (H,), p[pre + ".w"], p[pre + ".b"], 1e-12)
    x = ln((p["sty.emb.word"][ids] + p["sty.emb.type"][tts]) + p["sty.emb.pos"][:N][None], "sty.emb.ln")
    valid = (torch.arange(N)[None, :] < lens[:, None])
    for i in range(sc.num_hidden_layers):
        o = "sty.%d" % i
        qkv = x @ from_tc_layout(p[o + ".wqkv.tc"], H, 3 * H) + p[o + ".bqkv"]
        q, k, v = (qkv[..., j * H:(j + 1) * H].view(B, N, nh, dk).transpose(1, 2) for j in range(3))      # head h = columns [h*dk, (h+1)*dk) of each third
        s = (q @ k.transpose(-1, -2)) / math.sqrt(dk)
        s = s.masked_fill(~valid[:, None, None, :], float("-inf"))
        ctx = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, N, H)
        x = ln(ctx @ from_tc_layout(p[o + ".wo.tc"], H, H) + p[o + ".bo"] + x, o + ".ln1")
        h = F.gelu(x @ from_tc_layout(p[o + ".w1.tc"], H, I) + p[o + ".b1"])
        x = ln(h @ from_tc_layout(p[o + ".w2.tc"], I, H) + p[o + ".b2"] + x, o + ".ln2")
    pooled = torch.tanh(x[:, 0] @ p["sty.pool.w"] + p["sty.pool.b"])
    heads = pooled @ p["sty.heads.w"] + p["sty.heads.b"]
    assert rel_max(pooled, g["pooled_output"]) <= 1e-5
    slices, _ = packing.style_head_slices(sc)
    for n, (c0, kk) in slices.items():
        assert rel_max(heads[:, c0:c0 + kk], g[n + "_outputs"]) <= 1e-5
