"""The caller-side plumbing either side of the hot path: the
``<speaker>|<prompt>|<phoneme>|<content>`` input contract and a micro-batching queue.

Every reference front-end (inference_am_vocoder_joint.py:76-131, demo_page.py:119-150,
openaiapi.py:110-142) does the same four things around ``generator(...)``: read the symbol tables,
split the 4-field line, map phonemes / speaker to ids, and run B=1.  This module restates that
plumbing for hosts that want it without the reference scripts, and adds what turns batch throughput
into served throughput: a queue that groups concurrent requests into one padded forward.  Because the
engine's batches are bitwise equal to B=1 runs, micro-batching is invisible to callers.

Pure host code: no CUDA, no torch ops beyond tensor construction.  The style / content vectors come
from the caller (the simbert encoder is out of scope).

Prompt side (s8f rank 1): ``PromptEmbeddingCache`` -- batched, cached ``get_style_embedding``.
Output side (s8f rank 2): ``fetch_pcm16`` (GPU int16 conversion + one pinned device->host copy + per-item trim) and
``pcm16_to_wav_bytes`` (the 16 kHz mono PCM16 RIFF image the front-ends emit).
"""
import struct
import threading
import time
from collections import namedtuple
from concurrent.futures import Future

import numpy as np
import torch

Request = namedtuple("Request", "speaker prompt phonemes content")


def load_symbol_table(path):
    """line -> index, exactly like inference_am_vocoder_joint.py:76-80 (tokenlist, speaker2)."""
    with open(path, encoding="utf-8") as f:
        return {t.strip(): idx for idx, t in enumerate(f.readlines())}


def parse_line(line):
    """``speaker|prompt|phonemes|content`` (inference_am_vocoder_joint.py:96-102).  Phonemes are
    whitespace separated; extra fields are ignored like the reference ignores them."""
    parts = line.strip().split("|")
    if len(parts) < 4:
        raise ValueError("expected <speaker>|<prompt>|<phoneme>|<content>, got %d field(s)" % len(parts))
    return Request(parts[0], parts[1], parts[2].split(), parts[3])


def encode(req, token2id, speaker2id):
    """-> (int64 ids, speaker id) or None for an unknown speaker (the reference skips such lines,
    inference_am_vocoder_joint.py:109-110).  An unknown phoneme raises KeyError like the reference (:113)."""
    if req.speaker not in speaker2id:
        return None
    ids = np.asarray([token2id[ph] for ph in req.phonemes], dtype=np.int64)
    if ids.size == 0:
        raise ValueError("empty phoneme sequence")
    return ids, int(speaker2id[req.speaker])


def collate(items, device="cpu", pad_id=0):
    """items: list of (ids, speaker_id, style_vec, content_vec) -> the keyword arguments of
    ``JETSGenerator.forward`` (inference_am_vocoder_joint.py:113-128), padded with id 0 (the collate
    convention of the reference's dataset, prompt_dataset.py:183)."""
    B, T = len(items), max(len(it[0]) for it in items)
    ling = np.full((B, T), pad_id, dtype=np.int64)
    for b, it in enumerate(items):
        ling[b, :len(it[0])] = it[0]
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)

    def vecs(col):
        # rows may be numpy arrays / lists, or tensors on any device (PromptEmbeddingCache.embed keeps them on the GPU)
        rows = [it[col] for it in items]
        if any(isinstance(r, torch.Tensor) for r in rows):
            return torch.stack([torch.as_tensor(r, dtype=torch.float32).to(device) for r in rows])
        return to(np.stack([np.asarray(r, dtype=np.float32) for r in rows]))

    return dict(
        inputs_ling=to(ling),
        input_lengths=to(np.asarray([len(it[0]) for it in items], dtype=np.int64)),
        inputs_speaker=to(np.asarray([it[1] for it in items], dtype=np.int64)),
        inputs_style_embedding=vecs(2),
        inputs_content_embedding=vecs(3),
    )


class MicroBatcher:
    """Groups concurrent synthesis requests into padded batches.

    ``forward(**kwargs) -> dict`` is the model (``emotivoice_b200.modules.JETSGenerator``); requests are
    ``(ids, speaker_id, style_vec, content_vec)``.  ``submit`` returns a Future whose result is the item's
    float32 waveform trimmed to its own length (``mel_lengths[b] * hop``).  A worker thread collects up to
    ``max_batch`` requests, waiting at most ``max_wait_s`` after the first one, and runs ONE forward.
    Errors of a batch are delivered to every future of that batch.
    """

    def __init__(self, forward, device="cpu", max_batch=32, max_wait_s=0.005, hop=256):
        self._forward, self._device, self._hop = forward, device, hop
        self._max_batch, self._max_wait = int(max_batch), float(max_wait_s)
        self._lock = threading.Condition()
        self._queue = []
        self._closed = False
        self.batches_run = 0
        self._thread = threading.Thread(target=self._loop, name="ev-microbatcher", daemon=True)
        self._thread.start()

    def submit(self, ids, speaker_id, style_vec, content_vec):
        fut = Future()
        with self._lock:
            if self._closed:
                raise RuntimeError("MicroBatcher is closed")
            self._queue.append(((np.asarray(ids, dtype=np.int64), int(speaker_id), style_vec, content_vec), fut))
            self._lock.notify()
        return fut

    def close(self):
        with self._lock:
            self._closed = True
            self._lock.notify()
        self._thread.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _take(self):
        with self._lock:
            while not self._queue and not self._closed:
                self._lock.wait()
            if not self._queue:
                return None
            deadline = time.monotonic() + self._max_wait
            while len(self._queue) < self._max_batch and not self._closed:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                self._lock.wait(left)
            batch, self._queue = self._queue[:self._max_batch], self._queue[self._max_batch:]
            return batch

    def _loop(self):
        while True:
            batch = self._take()
            if batch is None:
                return
            futs = [f for _, f in batch]
            try:
                out = self._forward(**collate([it for it, _ in batch], self._device))
                wav = out["wav_predictions"]
                lens = out.get("mel_lengths")
                lens = [int(wav.shape[-1]) // self._hop] * len(batch) if lens is None else [int(v) for v in lens.tolist()]
                wav = wav.detach().cpu()
                self.batches_run += 1
                for b, f in enumerate(futs):
                    f.set_result(wav[b, 0, :lens[b] * self._hop].clone())
            except BaseException as e:       # deliver, keep serving
                for f in futs:
                    if not f.done():
                        f.set_exception(e)


# ---- output side: the on-wire format every front-end emits -----------------------------------

def pcm16_to_wav_bytes(pcm, sample_rate=16000):
    """int16 mono samples -> a complete RIFF/WAVE file image (PCM, 16 bit, mono), i.e. what the reference writes with
    ``sf.write(path, int16_audio, samplerate=config.sampling_rate)`` (inference_am_vocoder_joint.py:132-134) and what
    openaiapi.py:139-140,172-174 sends for ``response_format="wav"``.  44-byte canonical header, little endian."""
    pcm = np.ascontiguousarray(np.asarray(pcm))
    if pcm.dtype != np.int16 or pcm.ndim != 1:
        raise ValueError("expected a 1-D int16 array, got %s %s" % (pcm.dtype, pcm.shape))
    data = pcm.astype("<i2", copy=False).tobytes()
    if len(data) > 0xFFFFFFFF - 36:
        raise ValueError("waveform too long for a RIFF container")
    header = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + len(data), b"WAVE", b"fmt ", 16,
                         1, 1, int(sample_rate), int(sample_rate) * 2, 2, 16, b"data", len(data))
    return header + data


def fetch_pcm16(model, out, hop=256):
    """Finish one forward the way the callers do (inference_am_vocoder_joint.py:130-131), without the fp32 waveform
    ever crossing PCIe: ``wav * 32768 -> int16`` on the GPU (``model.to_pcm16``), ONE device->host copy of the int16
    batch into pinned memory, then per-item trimming to ``mel_lengths[b] * hop`` on the host.
    ``out`` is the dict ``model(...)`` returned.  Returns a list of 1-D int16 numpy arrays (one per batch item)."""
    wav = out["wav_predictions"]
    pcm = model.to_pcm16(wav)                                              # (B, 1, 256 F) int16, device (saturating, see to_pcm16)
    host = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True)
    host.copy_(pcm, non_blocking=True)
    lens = out.get("mel_lengths")
    lens = None if lens is None else [int(v) for v in lens.tolist()]      # tiny D2H; also orders after the copy above
    torch.cuda.current_stream(pcm.device).synchronize()
    B, n = pcm.shape[0], int(pcm.shape[-1])
    arr = host.numpy().reshape(B, n)
    return [arr[b, :(n if lens is None else min(n, lens[b] * hop))].copy() for b in range(B)]


# ---- prompt / content embeddings ("batched, prompt-embedding cache") -----------------------------

class PromptEmbeddingCache:
    """The callers' ``get_style_embedding`` (inference_am_vocoder_joint.py:25-38) for many texts at once, with a cache.

    The reference tokenises ONE text and runs the BERT style encoder on the CPU, twice per utterance (prompt and content,
    :106-107), recomputing identical prompts ("Happy", "Sad", ... repeat for every line).  Here all texts of a call that are
    not cached go through the tokenizer as one right-padded batch and through ``style_encoder`` as ONE forward; results are
    kept (LRU, on the encoder's device) keyed by the text.  Batching is invisible: the engine's padded batches are bitwise
    equal to single-item calls.

    ``tokenizer(list_of_str, return_tensors="pt", padding=True)`` -> dict with input_ids / token_type_ids / attention_mask
    (a transformers tokenizer, as in the reference); ``style_encoder(**those)`` -> dict with "pooled_output" (B, 768)
    (``emotivoice_b200.style.StyleEncoder`` or the reference's own module).
    """

    def __init__(self, tokenizer, style_encoder, device=None, max_entries=4096):
        from collections import OrderedDict
        self._tok, self._enc, self._device = tokenizer, style_encoder, device
        self._max = int(max_entries)
        self._cache = OrderedDict()
        self._lock = threading.Lock()
        self.hits = self.misses = self.forwards = 0

    def embed(self, texts):
        """list of str -> (len(texts), D) float32 tensor, row i = pooled_output of texts[i]."""
        texts = list(texts)
        uniq = list(dict.fromkeys(texts))
        have = {}
        # rows are captured into a local dict inside the critical section that finds them, so a concurrent embed() that evicts
        # them afterwards cannot make this call's final lookup fail
        with self._lock:
            for t in uniq:
                row = self._cache.get(t)
                if row is not None:
                    self._cache.move_to_end(t)
                    have[t] = row
            n_hit = sum(1 for t in texts if t in have)
            self.hits += n_hit
            self.misses += len(texts) - n_hit
        missing = [t for t in uniq if t not in have]
        if missing:
            enc = self._tok(missing, return_tensors="pt", padding=True)
            keys = ("input_ids", "token_type_ids", "attention_mask")
            args = {k: (enc[k].to(self._device) if self._device is not None else enc[k]) for k in keys}
            with torch.no_grad():
                pooled = self._enc(**args)["pooled_output"].detach()
            fresh = {t: pooled[i].clone() for i, t in enumerate(missing)}
            have.update(fresh)
            with self._lock:
                self.forwards += 1
                self._cache.update(fresh)
                while len(self._cache) > self._max:
                    self._cache.popitem(last=False)
        return torch.stack([have[t] for t in texts])
