"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's style encoder.

The reference (simbert.py:33-72) wraps ``transformers.AutoModel.from_pretrained('WangZeJun/simbert-base-chinese')`` -- a
BertModel -- and reads its ``pooler_output``; the callers run it twice per utterance on the CPU
(inference_am_vocoder_joint.py:25-38,106-107).  The arithmetic lives in the third-party ``transformers`` package
(unpinned in the reference's requirements.txt; 5.5 in this image), absent from /root/reference, so this file restates the
published BERT algorithm (Devlin et al. 2018; transformers' modeling_bert.py: BertEmbeddings, BertSelfAttention,
BertSelfOutput, BertIntermediate, BertOutput, BertPooler) in plain functional PyTorch and is pinned by
``oracle/make_golden_style.py`` against the reference's own ``StyleEncoder`` class driving transformers' BertModel.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-12      # BertConfig.layer_norm_eps


def _ln(sd, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"].to(x.dtype), sd[prefix + ".bias"].to(x.dtype), LN_EPS)


def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"].to(x.dtype), sd[prefix + ".bias"].to(x.dtype))


def bert_forward(sd, n_heads, input_ids, token_type_ids, attention_mask, dtype=torch.float32, prefix="bert."):
    """BertModel.forward -> (last_hidden_state, pooler_output).  Absolute position embeddings, post-LN blocks, exact-erf
    GELU, additive key mask of finfo.min on padded keys (padded *query* rows are computed like the library does)."""
    B, N = input_ids.shape
    e = prefix + "embeddings."
    x = sd[e + "word_embeddings.weight"].to(dtype)[input_ids] + sd[e + "token_type_embeddings.weight"].to(dtype)[token_type_ids]
    x = x + sd[e + "position_embeddings.weight"].to(dtype)[:N][None]
    x = _ln(sd, e + "LayerNorm", x)
    H = x.shape[-1]
    dk = H // n_heads
    bias = (1.0 - attention_mask[:, None, None, :].to(dtype)) * torch.finfo(dtype).min      # (B,1,1,N)
    i = 0
    while (prefix + "encoder.layer.%d.attention.self.query.weight" % i) in sd:
        p = prefix + "encoder.layer.%d." % i
        split = lambda t: t.view(B, N, n_heads, dk).transpose(1, 2)
        q, k, v = (split(_lin(sd, p + "attention.self." + n, x)) for n in ("query", "key", "value"))
        s = q @ k.transpose(-1, -2) / math.sqrt(dk) + bias
        ctx = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, N, H)
        x = _ln(sd, p + "attention.output.LayerNorm", _lin(sd, p + "attention.output.dense", ctx) + x)
        h = F.gelu(_lin(sd, p + "intermediate.dense", x))
        x = _ln(sd, p + "output.LayerNorm", _lin(sd, p + "output.dense", h) + x)
        i += 1
    pooled = torch.tanh(_lin(sd, prefix + "pooler.dense", x[:, 0]))
    return x, pooled


def style_forward(sd, n_heads, input_ids, token_type_ids, attention_mask, dtype=torch.float32):
    """StyleEncoder.forward (simbert.py:48-72): pooled output + the four classification heads (dropout off)."""
    _, pooled = bert_forward(sd, n_heads, input_ids, token_type_ids, attention_mask, dtype)
    out = {"pooled_output": pooled}
    for n in ("pitch", "speed", "energy", "emotion"):
        out[n + "_outputs"] = _lin(sd, n + "_clf.classifier", pooled)
    return out
