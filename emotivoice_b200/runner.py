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


def broadcast_engine(model, device, src=0):
    """One-time weight distribution for a multi-GPU run: rank ``src`` packs the model once (weight-norm fold, layout change,
    tensor-core hi/lo planes: ~625 MB, seconds of CPU work) and every other rank RECEIVES the packed blob with one
    ``dist.broadcast`` (NCCL over NVLink on the GPU box) instead of re-packing its own copy.  Every rank's module must
    already hold the same parameters (``broadcast_state_dict``).  Returns the blob size in bytes."""
    import torch.distributed as dist
    rank = dist.get_rank()
    if rank == src:
        eng = model._engine()
        obj = [eng.index_meta(), int(eng.blob.numel())]
    else:
        eng, obj = None, [None, None]
    dist.broadcast_object_list(obj, src=src)
    meta, n = obj
    blob = eng.blob if rank == src else torch.empty(n, dtype=torch.float32, device=device)
    dist.broadcast(blob, src=src)
    if rank != src:
        model.attach_packed(blob, meta)
    return n * 4


def utterance_digest(pcm):
    """sha1 of one utterance's int16 samples (little endian bytes)."""
    import hashlib
    return hashlib.sha1(memoryview(pcm).cast("B")).hexdigest()


def combine_digests(pairs):
    """[(utterance index, digest)] from any number of ranks -> one digest that is the same for every sharding of the same
    corpus iff every utterance's PCM is bit-identical."""
    import hashlib
    h = hashlib.sha1()
    for i, d in sorted(pairs):
        h.update(("%d:%s;" % (i, d)).encode())
    return h.hexdigest()


@torch.no_grad()
def synthesize_corpus(model, utterances, device, batch_size=32, indices=None, keep_pcm=True, stats=None):
    """Runs ``model`` (emotivoice_b200.modules.JETSGenerator) over a list of utterances (dicts from
    synth.make_utterance / synth.corpus_utterance) in length-bucketed batches, the offline-batch scheme of
    inference_tts.py:178-220 with one process per GPU.  Per bucket: collate into pinned host memory -> async H2D ->
    forward (one host sync inside: the data-dependent output length) -> int16 conversion on the GPU -> async D2H into a
    pinned buffer.  The host-side trimming / hashing of bucket i runs while the GPU computes bucket i+1.
    Returns {index: (pcm16 numpy or None, n_frames, sha1)}.  Batch composition does not change results
    (batch-invariant contract), so outputs are independent of batch size, sharding and world size.
    ``stats`` (dict, optional) receives wall-clock seconds per phase: collate_s, forward_call_s (includes waiting for the
    device at the length sync), finish_s (trim + hash), buckets, frames."""
    import time
    import numpy as np
    indices = list(range(len(utterances))) if indices is None else list(indices)
    lens = [len(u["ids"]) for u in utterances]
    results = {}
    up = model.upsample_factor
    st = dict(collate_s=0.0, forward_call_s=0.0, finish_s=0.0, buckets=0, frames=0)
    pending = None        # (group, pinned pcm, lengths list, event)

    def finish(p):
        group, host, ml, ev = p
        ev.synchronize()
        t0 = time.perf_counter()
        arr = host.numpy().reshape(len(group), -1)
        for r, i in enumerate(group):
            pcm = arr[r, :ml[r] * up]
            results[i] = (pcm.copy() if keep_pcm else None, ml[r], utterance_digest(np.ascontiguousarray(pcm)))
            st["frames"] += ml[r]
        st["finish_s"] += time.perf_counter() - t0

    for group in bucket_batches(indices, lens, batch_size):
        t0 = time.perf_counter()
        batch = synth.collate_utterances([utterances[i] for i in group], pin=True)
        batch = {k: v.to(device, non_blocking=True) for k, v in batch.items()}
        t1 = time.perf_counter()
        out = model(**batch)
        pcm = model.to_pcm16(out["wav_predictions"])
        host = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True)
        host.copy_(pcm, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        ml = [int(v) for v in out["mel_lengths_host"].tolist()]
        t2 = time.perf_counter()
        st["collate_s"] += t1 - t0
        st["forward_call_s"] += t2 - t1
        st["buckets"] += 1
        if pending is not None:
            finish(pending)          # the previous bucket's copy completed before this bucket's length sync returned
        pending = (group, host, ml, ev)
    if pending is not None:
        finish(pending)
    if stats is not None:
        stats.update(st)
    return results
