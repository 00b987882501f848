"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the training-mode alignment helpers:
monotonic alignment search, its duration / bin-loss post-processing and the per-token averaging of frame-level targets
(models/prompt_tts_modified/modules/alignment.py:90-177).  The reference runs them per sample in numba on the CPU, between
GPU stages of the training step.  Pinned against the reference's own numba functions by oracle/make_golden_align.py.

Arithmetic that decides the (integer) path, restated exactly:
  * log_p_attn arrives as float32; Q is float64 (np.full default); row 0 is the float32 running sum of log_prob[0, :j+1]
    (numba sums a float32 slice sequentially in float32) widened to float64; every other cell is a float64 max + float32 add;
  * cells with i > j keep -inf; the backtrack prefers the SMALLER token index on ties (`Q[i_a, j] >= Q[i_b, j]`).
"""
import numpy as np


def monotonic_alignment_search(log_p_attn):
    """(T_mel, T_inp) float32 -> (T_mel,) int64 token index per frame (alignment.py:90-121)."""
    lp = np.ascontiguousarray(log_p_attn, dtype=np.float32).T          # (T_inp, T_mel)
    T_inp, T_mel = lp.shape
    Q = np.full((T_inp, T_mel), -np.inf, dtype=np.float64)
    Q[0] = np.cumsum(lp[0], dtype=np.float32).astype(np.float64)       # sequential float32 prefix sums
    for j in range(1, T_mel):
        hi = min(j + 1, T_inp)
        if hi > 1:
            Q[1:hi, j] = np.maximum(Q[0:hi - 1, j - 1], Q[1:hi, j - 1]) + lp[1:hi, j].astype(np.float64)
    A = np.full((T_mel,), T_inp - 1, dtype=np.int64)
    for j in range(T_mel - 2, -1, -1):
        i_b = A[j + 1]
        i_a = i_b - 1
        if i_b == 0:
            A[j] = 0
        elif Q[i_a, j] >= Q[i_b, j]:
            A[j] = i_a
        else:
            A[j] = i_b
    return A


def viterbi_decode(log_p_attn, text_lengths, feats_lengths):
    """(B, T_mel, T_inp) float32 -> durations (B, T_inp) float32, bin_loss scalar (alignment.py:124-142)."""
    B, _, T_text = log_p_attn.shape
    ds = np.zeros((B, T_text), np.float32)
    bin_loss = 0.0
    for b in range(B):
        cur = log_p_attn[b, :feats_lengths[b], :text_lengths[b]]
        path = monotonic_alignment_search(cur)
        cnt = np.bincount(path)
        ds[b, :len(cnt)] = cnt
        bin_loss = bin_loss - np.float32(cur[np.arange(feats_lengths[b]), path].astype(np.float32).mean(dtype=np.float32))
    return ds, np.float32(bin_loss / B)


def average_by_duration(ds, xs, text_lengths, feats_lengths):
    """Per-token mean of a frame-level track over the token's frames; 0 for zero-length tokens (alignment.py:145-177)."""
    B = ds.shape[0]
    out = np.zeros_like(ds, dtype=np.float32)
    d = ds.astype(np.int32)
    for b in range(B):
        cs = np.concatenate([[0], np.cumsum(d[b, :text_lengths[b]])])
        x = xs[b, :feats_lengths[b]]
        for n, (s, e) in enumerate(zip(cs[:-1], cs[1:])):
            seg = x[s:e]
            out[b, n] = seg.mean(dtype=np.float32) if len(seg) else 0.0
    return out


def alignment_module_forward(sd, text, feats, text_lengths, feats_lengths, x_masks=None, prefix="", prior_fn=None):
    """AlignmentModule.forward (alignment.py:33-56) as plain functional torch on the CPU.  ``sd``: the module's state dict
    (t_conv1.weight ...).  The prior comes from ``prior_fn(text_lengths, feats_lengths)`` (the product's host code builds it with
    the same scipy call as the reference; passing the reference's own here keeps this function a pure restatement)."""
    import torch
    import torch.nn.functional as F
    g = lambda n: sd[prefix + n]
    t = text.transpose(1, 2)
    t = F.relu(F.conv1d(t, g("t_conv1.weight"), g("t_conv1.bias"), padding=1))
    t = F.conv1d(t, g("t_conv2.weight"), g("t_conv2.bias")).transpose(1, 2)
    f = feats.transpose(1, 2)
    f = F.relu(F.conv1d(f, g("f_conv1.weight"), g("f_conv1.bias"), padding=1))
    f = F.relu(F.conv1d(f, g("f_conv2.weight"), g("f_conv2.bias"), padding=1))
    f = F.conv1d(f, g("f_conv3.weight"), g("f_conv3.bias")).transpose(1, 2)
    score = -torch.norm(f.unsqueeze(2) - t.unsqueeze(1), p=2, dim=3)
    if x_masks is not None:
        score = score.masked_fill(x_masks.unsqueeze(-2), -np.inf)
    lp = F.log_softmax(score, dim=-1)
    return lp + prior_fn(text_lengths, feats_lengths).to(lp.dtype) if prior_fn is not None else lp


def get_segments(x, start_idxs, segment_size):
    """models/hifigan/get_random_segments.py:19-27 in numpy: (B, C, T) -> (B, C, segment_size), zero padded."""
    B, C, T = x.shape
    out = np.zeros((B, C, segment_size), x.dtype)
    for b in range(B):
        s = int(start_idxs[b])
        seg = x[b, :, s:s + segment_size]
        out[b, :, :seg.shape[1]] = seg
    return out
