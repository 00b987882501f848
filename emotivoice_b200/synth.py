"""Deterministic synthetic weights and inputs (no checkpoint, no network).

The released generator checkpoint (``g_00140000``) and the simbert style encoder
are not available offline, so parity tests and benchmarks run on seeded random
weights with the *reference's* state-dict names and shapes and on
synthetic phoneme/style inputs of the shapes BASELINE.json names.

numpy's ``default_rng`` (PCG64) is used instead of torch's RNG so that the very
same tensors are reproduced on the GPU box; ``state_dict_digest`` lets a test
verify that.
"""
import hashlib
import math

import numpy as np
import torch

SEED = 1234  # Config.seed, config/joint/config.py:83


def _xavier(rng, shape):
    # torch.nn.init.xavier_uniform_ as applied by initialize.py:11-34
    rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    fan_in, fan_out = shape[1] * rf, shape[0] * rf
    a = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-a, a, size=shape).astype(np.float32)


def am_param_shapes(conf):
    """(name, shape, kind) for every ``am.*`` tensor, in the reference's
    registration order (model_open_source.py:15-100)."""
    m = conf.model
    out = []
    for pre, H, nl, ks in (("am.encoder", m.encoder_n_hidden, m.encoder_n_layers, m.encoder_kernel_size_conv_mod),
                           ("am.decoder", m.decoder_n_hidden, m.decoder_n_layers, m.decoder_kernel_size_conv_mod)):
        out.append((pre + ".embed.0.alpha", (), "alpha"))
        for i in range(nl):
            p = "%s.encoders.%d" % (pre, i)
            for n in ("q", "k", "v", "out"):
                out.append(("%s.self_attn.linear_%s.weight" % (p, n), (H, H), "w"))
                out.append(("%s.self_attn.linear_%s.bias" % (p, n), (H,), "b"))
            out.append((p + ".feed_forward.w_1.weight", (4 * H, H, ks), "w"))
            out.append((p + ".feed_forward.w_1.bias", (4 * H,), "b"))
            out.append((p + ".feed_forward.w_2.weight", (H, 4 * H, ks), "w"))
            out.append((p + ".feed_forward.w_2.bias", (H,), "b"))
            for n in ("norm1", "norm2"):
                out.append(("%s.%s.weight" % (p, n), (H,), "ln_w"))
                out.append(("%s.%s.bias" % (p, n), (H,), "ln_b"))
        out.append((pre + ".after_norm.weight", (H,), "ln_w"))
        out.append((pre + ".after_norm.bias", (H,), "ln_b"))
    H, C = m.encoder_n_hidden, m.variance_n_hidden

    def predictor(pre, n_layers, ks):
        for i in range(n_layers):
            cin = H if i == 0 else C
            out.append(("%s.conv.%d.0.weight" % (pre, i), (C, cin, ks), "w"))
            out.append(("%s.conv.%d.0.bias" % (pre, i), (C,), "b"))
            out.append(("%s.conv.%d.2.weight" % (pre, i), (C,), "ln_w"))
            out.append(("%s.conv.%d.2.bias" % (pre, i), (C,), "ln_b"))
        out.append((pre + ".linear.weight", (1, C), "w"))
        out.append((pre + ".linear.bias", (1,), "b"))

    predictor("am.duration_predictor", m.duration_n_layers, m.duration_kernel_size)
    predictor("am.pitch_predictor", m.variance_n_layers, m.variance_kernel_size)
    out.append(("am.pitch_embed.0.weight", (H, 1, m.variance_embed_kernel_size), "w"))
    out.append(("am.pitch_embed.0.bias", (H,), "b"))
    predictor("am.energy_predictor", 2, 3)  # model_open_source.py:70-76 (hard-coded)
    out.append(("am.energy_embed.0.weight", (H, 1, m.variance_embed_kernel_size), "w"))
    out.append(("am.energy_embed.0.bias", (H,), "b"))
    # AlignmentModule(adim, odim) alignment.py:13-32: training only, must load
    nm = conf.n_mels
    for n, s in (("t_conv1", (H, H, 3)), ("t_conv2", (H, H, 1)), ("f_conv1", (H, nm, 3)),
                 ("f_conv2", (H, H, 3)), ("f_conv3", (H, H, 1))):
        out.append(("am.alignment_module.%s.weight" % n, s, "w"))
        out.append(("am.alignment_module.%s.bias" % n, (s[0],), "b"))
    out.append(("am.to_mel.weight", (nm, m.decoder_n_hidden), "w"))
    out.append(("am.to_mel.bias", (nm,), "b"))
    out.append(("am.spk_tokenizer.weight", (conf.n_speaker, H), "emb"))
    out.append(("am.src_word_emb.weight", (conf.n_vocab, H), "emb"))
    out.append(("am.embed_projection1.weight", (H, 2 * H + 2 * m.bert_embedding), "w"))
    out.append(("am.embed_projection1.bias", (H,), "b"))
    return out


def vocoder_conv_shapes(model_conf):
    """(module name, weight shape, transposed?) for every weight-normed conv of
    the HiFi-GAN generator (hifigan/models.py:90-113)."""
    h = model_conf
    c0 = h.upsample_initial_channel
    out = [("conv_pre", (c0, h.initial_channel, 7), False)]
    for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
        out.append(("ups.%d" % i, (c0 // 2 ** i, c0 // 2 ** (i + 1), k), True))
    nk = len(h.resblock_kernel_sizes)
    ch = c0
    for i in range(len(h.upsample_rates)):
        ch = c0 // 2 ** (i + 1)
        for j, k in enumerate(h.resblock_kernel_sizes):
            for grp in ("convs1", "convs2"):
                for l in range(len(h.resblock_dilation_sizes[j])):
                    out.append(("resblocks.%d.%s.%d" % (i * nk + j, grp, l), (ch, ch, k), False))
    out.append(("conv_post", (1, ch, 7), False))
    return out


def make_state_dict(conf, seed=SEED, legacy_weight_norm=False, dur_mean_frames=5.0,
                    am_bias_std=0.02, dur_log_scale=0.15):
    """Seeded random weights with the reference's names/shapes.

    AM: xavier-uniform matrices (initialize.py:11-34) but *non-zero* biases and
    non-trivial LayerNorm affine / alpha so every term of every kernel is
    exercised; embeddings N(0,1).  The duration head is scaled and biased so the
    mean predicted duration is ``dur_mean_frames`` frames/phoneme.
    Vocoder: v ~ U(+-1/sqrt(fan_in)) (what torch>=2.1 leaves after the no-op
    ``init_weights`` on parametrised convs), g = ||v|| * U(0.6, 1.4) so that the
    weight-norm fold is exercised.
    """
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape, kind in am_param_shapes(conf):
        if kind == "w":
            a = _xavier(rng, shape)
        elif kind == "b":
            a = rng.normal(0.0, am_bias_std, size=shape).astype(np.float32)
        elif kind == "ln_w":
            a = rng.uniform(0.8, 1.2, size=shape).astype(np.float32)
        elif kind == "ln_b":
            a = rng.normal(0.0, 0.05, size=shape).astype(np.float32)
        elif kind == "emb":
            a = rng.normal(0.0, 1.0, size=shape).astype(np.float32)
        elif kind == "alpha":
            a = np.float32(1.0 + 0.25 * rng.uniform(-1, 1))
        sd[name] = torch.from_numpy(np.asarray(a, dtype=np.float32).copy())
    sd["am.duration_predictor.linear.weight"] = sd["am.duration_predictor.linear.weight"] * dur_log_scale
    sd["am.duration_predictor.linear.bias"] = torch.full((1,), math.log(1.0 + dur_mean_frames))
    g_key, v_key = ("weight_g", "weight_v") if legacy_weight_norm else \
        ("parametrizations.weight.original0", "parametrizations.weight.original1")
    for mod, shape, transposed in vocoder_conv_shapes(conf.model):
        rf = shape[2]
        fan_in = (shape[0] if transposed else shape[1]) * rf
        bound = 1.0 / math.sqrt(fan_in)
        v = rng.uniform(-bound, bound, size=shape).astype(np.float32)
        norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
        g = (norm * rng.uniform(0.6, 1.4, size=norm.shape)).astype(np.float32)
        n_bias = shape[1] if transposed else shape[0]
        b = rng.uniform(-bound, bound, size=(n_bias,)).astype(np.float32)
        sd["generator.%s.bias" % mod] = torch.from_numpy(b)
        sd["generator.%s.%s" % (mod, g_key)] = torch.from_numpy(g)
        sd["generator.%s.%s" % (mod, v_key)] = torch.from_numpy(v)
    return sd


def state_dict_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


# ----------------------------------------------------------------------------
# inputs
# ----------------------------------------------------------------------------
USED_ID_LO, USED_ID_HI = 2, 416   # tokenlist ids 417-501 are unused placeholders
SOS_EOS_ID, PAD_ID = 1, 0         # tokenlist rows 1 / 0; collate pads with 0


def make_utterance(rng, n_phonemes, n_speaker=2014, bert_dim=768):
    ids = rng.integers(USED_ID_LO, USED_ID_HI + 1, size=n_phonemes).astype(np.int64)
    ids[0] = SOS_EOS_ID
    ids[-1] = SOS_EOS_ID
    spk = np.int64(rng.integers(0, n_speaker))
    style = np.tanh(rng.normal(size=bert_dim)).astype(np.float32)
    content = np.tanh(rng.normal(size=bert_dim)).astype(np.float32)
    return dict(ids=ids, speaker=spk, style=style, content=content)


def make_batch(lengths, seed=SEED, n_speaker=2014, bert_dim=768):
    """Padded batch in the reference's calling convention
    (inference_am_vocoder_joint.py:113-129): returns the keyword arguments of
    ``JETSGenerator.forward`` as CPU tensors."""
    rng = np.random.default_rng(seed)
    utts = [make_utterance(rng, int(n), n_speaker, bert_dim) for n in lengths]
    B, T = len(utts), int(max(lengths))
    ling = np.full((B, T), PAD_ID, dtype=np.int64)
    for b, u in enumerate(utts):
        ling[b, :len(u["ids"])] = u["ids"]
    return dict(
        inputs_ling=torch.from_numpy(ling),
        input_lengths=torch.tensor([int(n) for n in lengths], dtype=torch.int64),
        inputs_speaker=torch.tensor([int(u["speaker"]) for u in utts], dtype=torch.int64),
        inputs_style_embedding=torch.from_numpy(np.stack([u["style"] for u in utts])),
        inputs_content_embedding=torch.from_numpy(np.stack([u["content"] for u in utts])),
    )


def slice_batch(batch, b):
    """The B=1 call the reference CLI would make for item ``b`` of ``batch``."""
    n = int(batch["input_lengths"][b])
    return dict(
        inputs_ling=batch["inputs_ling"][b:b + 1, :n].contiguous(),
        input_lengths=batch["input_lengths"][b:b + 1].clone(),
        inputs_speaker=batch["inputs_speaker"][b:b + 1].clone(),
        inputs_style_embedding=batch["inputs_style_embedding"][b:b + 1].clone(),
        inputs_content_embedding=batch["inputs_content_embedding"][b:b + 1].clone(),
    )


# id ranges of the two symbol families in data/youdao/text/tokenlist (502 lines): ARPABET "[..]" + engsp* (EN) and pinyin
# initials/finals + sp* (ZH); 417-501 are unused "uncasedNN" placeholders, 0 = pad "_", 1 = <sos/eos>
EN_ID_RANGES = ((2, 70), (148, 150), (402, 402), (409, 409))
ZH_ID_RANGES = ((71, 147), (151, 401), (403, 408), (410, 416))


def _ids_from(ranges):
    return np.concatenate([np.arange(lo, hi + 1) for lo, hi in ranges]).astype(np.int64)


def corpus_utterance(index, seed=SEED, lo=20, hi=200, n_phonemes=None, n_speaker=2014, bert_dim=768):
    """Utterance ``index`` of the seeded synthetic corpus (BASELINE.json configs[2] / [4]: 20-200 phonemes, even indices
    drawn from the EN symbols, odd ones from the ZH symbols).  Each utterance has its own generator seeded by
    ``(seed, index)``, so it is the same whatever the corpus size, shard or world size -- what lets a multi-GPU run check
    that its outputs are bit-identical to a single-GPU run's."""
    rng = np.random.default_rng([int(seed), int(index)])
    n = int(n_phonemes) if n_phonemes is not None else int(rng.integers(lo, hi + 1))
    pool = _ids_from(EN_ID_RANGES if index % 2 == 0 else ZH_ID_RANGES)
    ids = pool[rng.integers(0, len(pool), size=n)]
    ids[0] = SOS_EOS_ID
    ids[-1] = SOS_EOS_ID
    return dict(ids=ids, speaker=np.int64(rng.integers(0, n_speaker)),
                style=np.tanh(rng.normal(size=bert_dim)).astype(np.float32),
                content=np.tanh(rng.normal(size=bert_dim)).astype(np.float32), index=int(index))


def corpus_lengths(n, seed=SEED, lo=20, hi=200, n_phonemes=None):
    """Phoneme counts of utterances 0..n-1 without building them (the sharding plan needs only these)."""
    if n_phonemes is not None:
        return [int(n_phonemes)] * n
    return [int(np.random.default_rng([int(seed), i]).integers(lo, hi + 1)) for i in range(n)]


def collate_utterances(utts, pin=False):
    """list of utterance dicts -> JETSGenerator.forward keyword arguments (CPU tensors, optionally pinned), padded with
    id 0 like the reference's collate (prompt_dataset.py:183)."""
    B, T = len(utts), max(len(u["ids"]) for u in utts)
    ling = np.full((B, T), PAD_ID, dtype=np.int64)
    for b, u in enumerate(utts):
        ling[b, :len(u["ids"])] = u["ids"]
    out = dict(
        inputs_ling=torch.from_numpy(ling),
        input_lengths=torch.tensor([len(u["ids"]) for u in utts], dtype=torch.int64),
        inputs_speaker=torch.tensor([int(u["speaker"]) for u in utts], dtype=torch.int64),
        inputs_style_embedding=torch.from_numpy(np.stack([u["style"] for u in utts])),
        inputs_content_embedding=torch.from_numpy(np.stack([u["content"] for u in utts])))
    return {k: v.pin_memory() for k, v in out.items()} if pin else out


def make_alignment_state_dict(adim=384, odim=80, seed=SEED + 7):
    """Seeded weights of the reference's AlignmentModule (alignment.py:19-25), torch's default Conv1d init range."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, cin, k in (("t_conv1", adim, 3), ("t_conv2", adim, 1), ("f_conv1", odim, 3), ("f_conv2", adim, 3), ("f_conv3", adim, 1)):
        bound = 1.0 / math.sqrt(cin * k)
        sd[name + ".weight"] = torch.from_numpy(rng.uniform(-bound, bound, size=(adim, cin, k)).astype(np.float32))
        sd[name + ".bias"] = torch.from_numpy(rng.uniform(-bound, bound, size=(adim,)).astype(np.float32))
    return sd


def make_mel(batch, frames, seed=SEED, n_mels=80):
    """cfg4 vocoder-only input: N(0,1)*1.2, shape (B, 80, F)."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy((rng.normal(size=(batch, n_mels, frames)) * 1.2).astype(np.float32))


# ---- style encoder: seeded BERT weights with the reference's state-dict names ----------------

STYLE_SEED = 4321


def style_config(small=False):
    """Architecture integers of the style encoder.  Full size = ``WangZeJun/simbert-base-chinese`` (config/joint/config.py:44:
    BERT-base, 12 x 768, 12 heads, FFN 3072, vocab 13685, 512 positions, 2 token types) + the four classification heads of
    simbert.py:39-42 with the label counts of data/youdao/text/{pitch,speed,energy,emotion}.  ``small`` is a 2-layer,
    256-wide model of the same structure for fast tests."""
    from .config import AttrDict
    if small:
        return AttrDict(vocab_size=1000, hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=1024,
                        max_position_embeddings=128, type_vocab_size=2, pitch_n_labels=3, speed_n_labels=3,
                        energy_n_labels=3, emotion_n_labels=7, style_dim=128)
    return AttrDict(vocab_size=13685, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                    max_position_embeddings=512, type_vocab_size=2, pitch_n_labels=3, speed_n_labels=3,
                    energy_n_labels=3, emotion_n_labels=7, style_dim=128)


def style_param_shapes(sc):
    """(name, shape, kind) of every tensor of the reference StyleEncoder's state dict (simbert.py:33-44: ``bert.*`` is
    transformers' BertModel, then the heads)."""
    H, I = sc.hidden_size, sc.intermediate_size
    out = [("bert.embeddings.word_embeddings.weight", (sc.vocab_size, H), "emb"),
           ("bert.embeddings.position_embeddings.weight", (sc.max_position_embeddings, H), "emb"),
           ("bert.embeddings.token_type_embeddings.weight", (sc.type_vocab_size, H), "emb"),
           ("bert.embeddings.LayerNorm.weight", (H,), "ln_w"), ("bert.embeddings.LayerNorm.bias", (H,), "ln_b")]
    for i in range(sc.num_hidden_layers):
        p = "bert.encoder.layer.%d." % i
        for n in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
            out += [(p + n + ".weight", (H, H), "w"), (p + n + ".bias", (H,), "b")]
        out += [(p + "attention.output.LayerNorm.weight", (H,), "ln_w"), (p + "attention.output.LayerNorm.bias", (H,), "ln_b"),
                (p + "intermediate.dense.weight", (I, H), "w"), (p + "intermediate.dense.bias", (I,), "b"),
                (p + "output.dense.weight", (H, I), "w"), (p + "output.dense.bias", (H,), "b"),
                (p + "output.LayerNorm.weight", (H,), "ln_w"), (p + "output.LayerNorm.bias", (H,), "ln_b")]
    out += [("bert.pooler.dense.weight", (H, H), "w"), ("bert.pooler.dense.bias", (H,), "b")]
    for n, k in (("pitch_clf", sc.pitch_n_labels), ("speed_clf", sc.speed_n_labels), ("energy_clf", sc.energy_n_labels),
                 ("emotion_clf", sc.emotion_n_labels)):
        out += [(n + ".classifier.weight", (k, H), "w"), (n + ".classifier.bias", (k,), "b")]
    out += [("style_embed_proj.weight", (sc.style_dim, H), "w"), ("style_embed_proj.bias", (sc.style_dim,), "b")]
    return out


def make_style_state_dict(sc, seed=STYLE_SEED):
    """Seeded weights in the range of a trained BERT (matrices N(0, 0.04) -- twice the HF init so attention is not
    uniform --, embeddings N(0, 0.05), non-trivial LayerNorm affine and biases so every term is exercised)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape, kind in style_param_shapes(sc):
        if kind == "w":
            a = rng.normal(0.0, 0.04, size=shape)
        elif kind == "b":
            a = rng.normal(0.0, 0.02, size=shape)
        elif kind == "ln_w":
            a = rng.uniform(0.8, 1.2, size=shape)
        elif kind == "ln_b":
            a = rng.normal(0.0, 0.05, size=shape)
        else:
            a = rng.normal(0.0, 0.05, size=shape)
        sd[name] = torch.from_numpy(a.astype(np.float32))
    return sd


def make_style_batch(sc, lengths, seed=STYLE_SEED + 1):
    """Tokenizer-shaped inputs (inference_am_vocoder_joint.py:25-29): ids with [CLS]=101 ... [SEP]=102 when the vocabulary
    has them, right-padded with 0; token_type_ids all zero; attention_mask = 1 on the valid prefix."""
    rng = np.random.default_rng(seed)
    B, N = len(lengths), int(max(lengths))
    ids = np.zeros((B, N), np.int64)
    mask = np.zeros((B, N), np.int64)
    for b, n in enumerate(lengths):
        row = rng.integers(min(200, sc.vocab_size // 2), sc.vocab_size, size=n)
        if sc.vocab_size > 102 and n >= 2:
            row[0], row[-1] = 101, 102
        ids[b, :n], mask[b, :n] = row, 1
    return dict(input_ids=torch.from_numpy(ids), token_type_ids=torch.zeros((B, N), dtype=torch.int64),
                attention_mask=torch.from_numpy(mask))
