"""Multi-GPU plumbing for offline batch synthesis.

The path shards naturally: utterances are independent end to end, so ranks exchange NOTHING
per sample.  The only collective is a one-time weight broadcast from rank 0 (NCCL over
NVLink on the GPU box, gloo in the CPU tests) and, for reporting, an all-reduce of
{frames, seconds}.  The reference's own scheme is one OS process per GPU over contiguous
line chunks (inference_tts.py:178-220); here the chunks are length-balanced instead.
"""
import torch

from . import synth


def plan_shards(n_phonemes, world_size):
    """Longest-processing-time assignment: cost of an utterance ~ its frame count ~ its
    phoneme count (vocoder-dominated, linear).  Returns a list (per rank) of utterance
    indices; deterministic, every utterance assigned exactly once."""
    order = sorted(range(len(n_phonemes)), key=lambda i: (-int(n_phonemes[i]), i))
    loads = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(n_phonemes[i])
    return shards


def bucket_batches(indices, n_phonemes, batch_size):
    """Within a rank: sort by length and cut into batches so padding is small."""
    idx = sorted(indices, key=lambda i: (-int(n_phonemes[i]), i))
    return [idx[i:i + batch_size] for i in range(0, len(idx), batch_size)]


def param_specs(conf):
    """(name, shape, dtype) of every tensor of the reference state dict, same on all ranks."""
    specs = [(n, tuple(s)) for n, s, _ in synth.am_param_shapes(conf)]
    for mod, shape, transposed in synth.vocoder_conv_shapes(conf.model):
        nb = shape[1] if transposed else shape[0]
        specs.append(("generator.%s.bias" % mod, (nb,)))
        specs.append(("generator.%s.parametrizations.weight.original0" % mod, (shape[0], 1, 1)))
        specs.append(("generator.%s.parametrizations.weight.original1" % mod, tuple(shape)))
    return specs


def broadcast_state_dict(sd, conf, device, src=0):
    """One-time weight broadcast: rank ``src`` flattens its state dict into ONE fp32 buffer
    (~213 MB), every rank receives it with a single ``dist.broadcast`` and unflattens.
    ``sd`` may be None on the other ranks.  Works with any initialised backend."""
    import torch.distributed as dist
    specs = param_specs(conf)
    total = sum(int(torch.Size(s).numel()) for _, s in specs)
    flat = torch.empty(total, dtype=torch.float32, device=device)
    if dist.get_rank() == src:
        off = 0
        for name, shape in specs:
            n = int(torch.Size(shape).numel())
            flat[off:off + n] = sd[name].reshape(-1).to(device=device, dtype=torch.float32)
            off += n
    dist.broadcast(flat, src=src)
    out, off = {}, 0
    for name, shape in specs:
        n = int(torch.Size(shape).numel())
        out[name] = flat[off:off + n].reshape(shape).clone()
        off += n
    return out


@torch.no_grad()
def synthesize_corpus(model, utterances, device, batch_size=32, indices=None):
    """Runs ``model`` (emotivoice_b200.modules.JETSGenerator) over a list of utterances
    (dicts from synth.make_utterance) in length-bucketed batches.  Returns
    {index: (pcm16 numpy, n_frames)}.  Batch composition does not change results
    (batch-invariant contract), so outputs are independent of world size."""
    import numpy as np
    indices = list(range(len(utterances))) if indices is None else list(indices)
    lens = [len(u["ids"]) for u in utterances]
    results = {}
    for group in bucket_batches(indices, lens, batch_size):
        T = max(lens[i] for i in group)
        ling = np.zeros((len(group), T), dtype=np.int64)
        for r, i in enumerate(group):
            ling[r, :lens[i]] = utterances[i]["ids"]
        batch = dict(
            inputs_ling=torch.from_numpy(ling).to(device),
            input_lengths=torch.tensor([lens[i] for i in group], dtype=torch.int64, device=device),
            inputs_speaker=torch.tensor([int(utterances[i]["speaker"]) for i in group], dtype=torch.int64, device=device),
            inputs_style_embedding=torch.from_numpy(np.stack([utterances[i]["style"] for i in group])).to(device),
            inputs_content_embedding=torch.from_numpy(np.stack([utterances[i]["content"] for i in group])).to(device))
        out = model(**batch)
        pcm = model.to_pcm16(out["wav_predictions"]).cpu().numpy()
        ml = out["mel_lengths"].cpu().tolist()
        up = model.upsample_factor
        for r, i in enumerate(group):
            results[i] = (pcm[r, 0, :ml[r] * up].copy(), ml[r])
    return results
