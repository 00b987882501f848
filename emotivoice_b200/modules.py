"""Host-side mirror of the reference's model boundary.

Same class names, constructor arguments, ``state_dict`` keys, ``forward()`` keyword
signature and returned dict as

* ``JETSGenerator``  -- models/prompt_tts_modified/jets.py:26-71
* ``PromptTTS``      -- models/prompt_tts_modified/model_open_source.py:14-163
* ``Generator``      -- models/hifigan/models.py:90-140

so the reference's callers (inference_am_vocoder_joint.py:70-74,120-129, demo_page.py:88-92,
openaiapi.py:78-82) run unchanged: ``JETSGenerator(conf).to(device)``,
``.load_state_dict(torch.load(path)['generator'])``, ``.eval()``, call under ``no_grad``.

The modules hold ``nn.Parameter`` trees only (for state-dict compatibility); every FLOP
of ``forward`` runs in libemotivoice_b200.so (hand-written sm_100a kernels) through the C
ABI in include/emotivoice_b200.h.  PyTorch provides device memory and the stream.  There
is no CPU path: calling ``forward`` on CPU tensors raises.
"""
import ctypes
import os
import threading

import numpy as np
import torch
import torch.nn as nn

from . import _abi, packing, synth


class _Holder(nn.Module):
    """Parameter container; never called."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: compute happens in libemotivoice_b200.so")


_HOST_WAIT_BLOCK = os.environ.get("EV_HOST_WAIT", "") == "block"
_TLS = threading.local()          # per-thread pinned read-back buffer + event


def _bucket(nbytes):
    """Workspace sizes are rounded up to a geometric series (x1.125 steps, 2 MiB granularity): utterances of similar length
    then request IDENTICAL sizes, so torch's caching allocator serves them from its pool instead of calling cudaMalloc
    (a device-synchronising, ~1-2 ms call) whenever a request is a little larger than anything it has cached."""
    g = 2 << 20
    n = max(int(nbytes), g)
    b = g
    while b < n:
        b = (b + (b >> 3) + g - 1) // g * g
    return b


def _register(root, dotted, tensor):
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Holder())
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


def _legacy_weight_norm_hook(module, state_dict, prefix, *args):
    """Checkpoints written by torch < 2.1 carry weight_g / weight_v instead of
    parametrizations.weight.original0/1 (checkpoint row): accept both."""
    for k in list(state_dict.keys()):
        if not k.startswith(prefix):
            continue
        if k.endswith(".weight_g"):
            state_dict[k[:-len("weight_g")] + "parametrizations.weight.original0"] = state_dict.pop(k)
        elif k.endswith(".weight_v"):
            state_dict[k[:-len("weight_v")] + "parametrizations.weight.original1"] = state_dict.pop(k)


def _dirty_post_hook(module, incompatible_keys):
    for m in module.modules():
        if isinstance(m, _EngineOwner):
            m._ev_dirty = True


class _Engine:
    """One ev_ctx + its packed weight blob and positional table on one device."""

    def __init__(self, conf, packed, device, precision="fp32", blob=None, index_meta=None):
        """``packed``: name -> tensor dict (packing.pack_state_dict + add_tc_weights), laid into one blob here; or pass a
        ready ``blob`` (fp32 tensor already on ``device``) with its ``index_meta`` [(name, offset, numel)] -- what the
        other ranks of a multi-GPU run receive from rank 0 instead of re-packing (runner.broadcast_engine)."""
        self.lib = _abi.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("emotivoice_b200 runs on CUDA (sm_100a) only; got device %s. "
                               "There is no CPU fallback." % (self.device,))
        self.index_dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.cfg = _abi.make_config(conf)
        self.hidden = int(self.cfg.hidden)
        handle = ctypes.c_void_p()
        _abi.check(self.lib.ev_create(ctypes.byref(handle), self.index_dev, ctypes.byref(self.cfg)))
        self.handle = handle
        if blob is None:
            blob, self.index = packing.make_blob(packed)
            self.blob = blob.to(self.device)
        else:
            self.index = packing.index_from_meta(index_meta)
            self.blob = blob
            assert blob.device == self.device and blob.dtype == torch.float32 and blob.is_contiguous()
        _abi.check(self.lib.ev_bind_weights(self.handle, self.blob.data_ptr(), self.blob.numel(),
                                            ctypes.cast(self.index, ctypes.c_void_p), len(self.index)))
        self.set_precision(precision)
        self.pe = None
        self._arena = {}              # (stream, kind) -> uint8 workspace, grow-only
        self.call_lock = threading.RLock()      # one forward at a time enqueues on an engine (its workspaces are reused, stream-ordered)
        self.ensure_pe(5000)          # PositionalEncoding max_len=5000 (encoder.py:206)
        self.total_up = int(np.prod([self.cfg.up_rates[i] for i in range(self.cfg.n_ups)]))

    def set_precision(self, precision):
        _abi.check(self.lib.ev_set_precision(self.handle, _abi.PRECISIONS[precision]))
        self.precision = precision

    def index_meta(self):
        return [(e.name.decode(), int(e.offset), int(e.numel)) for e in self.index]

    def ensure_pe(self, n):
        if self.pe is not None and self.pe.shape[0] >= n:
            return
        n = max(n, 2 * (self.pe.shape[0] if self.pe is not None else 0))
        self.pe = packing.build_pe_table(n, self.hidden).to(self.device)
        _abi.check(self.lib.ev_bind_pe(self.handle, self.pe.data_ptr(), n))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.ev_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- calls ------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _ws(self, kind, nbytes):
        """Grow-only workspace arena, one buffer per (stream, kind): after the largest request has been seen (or `reserve`d) no call
        allocates device memory any more -- a fresh cudaMalloc in the middle of a forward is a device-synchronising stall of
        milliseconds.  Reuse across calls is safe because every use is ordered on the stream the buffer belongs to."""
        key = (self._stream(), kind)
        t = self._arena.get(key)
        if t is None or t.numel() < nbytes:
            self._arena.pop(key, None)
            t = None
            t = torch.empty((_bucket(nbytes),), dtype=torch.uint8, device=self.device)
            self._arena[key] = t
        return t

    def reserve(self, batch, phonemes, frames):
        """Pre-size the arena of the current stream for requests up to (batch, phonemes, frames)."""
        lib = self.lib
        fs = {int(frames)}
        if batch * frames > 2400:                 # small batches carry extra grouped-launch buffers: cover the largest of those shapes too
            fs.add(max(1, 2400 // int(batch)))
        with self.call_lock:
            self.ensure_pe(max(int(frames), int(phonemes)))
            self._ws("p1", lib.ev_phase1_workspace_bytes(self.handle, int(batch), int(phonemes)))
            self._ws("p2", max(lib.ev_phase2_workspace_bytes(self.handle, int(batch), f) for f in fs))

    def _read_back(self, t):
        """Device int32 vector -> host tensor, waiting by polling (EV_HOST_WAIT=block: a blocking copy)."""
        if _HOST_WAIT_BLOCK:
            return t.cpu()
        tl = _TLS
        n = t.numel()
        if getattr(tl, "pin", None) is None or tl.pin.numel() < n:
            tl.pin = torch.empty((max(n, 1024),), dtype=torch.int32).pin_memory()
            tl.ev = torch.cuda.Event()
        tl.pin[:n].copy_(t, non_blocking=True)
        tl.ev.record(torch.cuda.current_stream(self.device))
        while not tl.ev.query():
            pass
        return tl.pin[:n].clone()

    def acoustic(self, ling, lens, spk, style, content, invariant):
        lib, dev = self.lib, self.device
        B, T = ling.shape
        self.ensure_pe(T)
        dur = torch.empty((B, T), dtype=torch.int64, device=dev)
        pitch = torch.empty((B, T), dtype=torch.float32, device=dev)
        energy = torch.empty((B, T), dtype=torch.float32, device=dev)
        meta = torch.empty((2 * B + 2,), dtype=torch.int32, device=dev)      # lens32 (B) | mel_lens (B) | max | input status
        ws1 = self._ws("p1", lib.ev_phase1_workspace_bytes(self.handle, B, T))
        n1 = ws1.numel()
        st = self._stream()
        lens32_ptr = meta.data_ptr()
        mel_lens_ptr = meta.data_ptr() + 4 * B
        _abi.check(lib.ev_am_phase1(self.handle, ling.data_ptr(), lens.data_ptr(), spk.data_ptr(), style.data_ptr(),
                                    content.data_ptr(), B, T, int(invariant), dur.data_ptr(), pitch.data_ptr(),
                                    energy.data_ptr(), lens32_ptr, mel_lens_ptr, ws1.data_ptr(), n1, st))
        # the path's single host sync: the output length is data dependent (alignment.py:194-195).  Asynchronous copy into pinned
        # memory + a polled event instead of a blocking .cpu(): a blocking wait of a few milliseconds puts the thread to sleep, and its
        # wake-up latency (measured: 5-15 ms now and then on a busy host) would sit in the middle of the forward with the GPU idle.
        mel_lens_host = self._read_back(meta[B:])
        status = int(mel_lens_host[B + 1])
        if status:      # what nn.Embedding / the mask construction raise in the reference (checked on the device, read with the lengths)
            if status & 1:
                raise IndexError("inputs_ling holds token ids outside [0, %d)" % int(self.cfg.n_vocab))
            if status & 2:
                raise IndexError("inputs_speaker holds ids outside [0, %d)" % int(self.cfg.n_speaker))
            raise RuntimeError("input_lengths must lie in [1, %d] (the padded width of inputs_ling)" % T)
        F = int(mel_lens_host[B])
        self.ensure_pe(F)
        ws2 = self._ws("p2", lib.ev_phase2_workspace_bytes(self.handle, B, F))
        n2 = ws2.numel()
        mel = torch.empty((B, F, int(self.cfg.n_mels)), dtype=torch.float32, device=dev)
        _abi.check(lib.ev_am_phase2(self.handle, ws1.data_ptr(), lens32_ptr, mel_lens_ptr, B, T, F, int(invariant),
                                    mel.data_ptr(), ws2.data_ptr(), n2, st))
        return dict(mel=mel, dur=dur, pitch=pitch, energy=energy, meta=meta, mel_lens=meta[B:2 * B],
                    mel_lens_host=mel_lens_host[:B], F=F, ws2=ws2, n2=n2)

    def vocode(self, mel, time_major, mel_lens_ptr, ws=None, n=0):
        lib, dev = self.lib, self.device
        if time_major:
            B, F, _ = mel.shape
        else:
            B, _, F = mel.shape
        if ws is None:
            ws = self._ws("p2", lib.ev_phase2_workspace_bytes(self.handle, B, F))
            n = ws.numel()
        wav = torch.empty((B, 1, F * self.total_up), dtype=torch.float32, device=dev)
        _abi.check(lib.ev_vocoder(self.handle, mel.data_ptr(), int(bool(time_major)), mel_lens_ptr, B, F,
                                  wav.data_ptr(), ws.data_ptr(), n, self._stream()))
        return wav


class _EngineOwner(nn.Module):
    """Lazy (re)packing of the parameter tree into an engine on the parameters' device."""

    def __init__(self):
        super().__init__()
        self._ev_engine = None
        self._ev_dirty = True
        self._ev_lock = threading.Lock()
        self._ev_precision = "fp32"
        self.register_load_state_dict_post_hook(_dirty_post_hook)

    @property
    def precision(self):
        """"fp32" (default): fp32-accurate 3xTF32 on the tcgen05 tensor cores (~1e-6 relative error).
        "tf32": decoder + vocoder with one tf32 MMA per K step (what the reference's eager PyTorch does for
        convolutions on a GPU); the duration-critical prefix stays fp32-accurate, so durations are
        identical in all modes.  "fp32_ffma": plain fp32 FFMA kernels, no tensor cores."""
        return self._ev_precision

    @precision.setter
    def precision(self, value):
        if value not in _abi.PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_abi.PRECISIONS))
        self._ev_precision = value
        if self._ev_engine is not None:
            self._ev_engine.set_precision(value)
        for m in self.children():            # JETSGenerator.precision also governs .am / .generator used stand-alone
            if isinstance(m, _EngineOwner):
                m.precision = value

    def _mark_dirty(self):
        self._ev_dirty = True

    def _apply(self, fn, *a, **k):          # .to() / .cuda() / .float() ...
        r = super()._apply(fn, *a, **k)
        self._ev_dirty = True
        return r

    def refresh_weights(self):
        """Call after modifying parameters in place (load_state_dict / .to() do it for you)."""
        self._ev_dirty = True

    def _pack(self):
        raise NotImplementedError

    def attach_packed(self, blob, index_meta):
        """Adopt a packed weight blob produced elsewhere (rank 0 of a multi-GPU run) instead of packing this module's own
        parameters: ``blob`` is already on the module's device.  The caller guarantees it corresponds to the parameters."""
        dev = next(self.parameters()).device
        with self._ev_lock:
            self._ev_engine = _Engine(self.config, None, dev, self._ev_precision, blob=blob, index_meta=index_meta)
            self._ev_dirty = False
        return self._ev_engine

    def _engine(self):
        eng = self._ev_engine
        dev = next(self.parameters()).device
        if eng is not None and not self._ev_dirty and eng.device == dev:
            return eng
        with self._ev_lock:
            if self._ev_engine is None or self._ev_dirty or self._ev_engine.device != dev:
                self._ev_engine = _Engine(self.config, packing.add_tc_weights(self._pack()), dev, self._ev_precision)
                self._ev_dirty = False
            return self._ev_engine


def _prep(t, dtype, device):
    if t.device != device:
        raise RuntimeError("input tensor on %s but the module is on %s" % (t.device, device))
    if t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_contiguous() else t.contiguous()


def _seeded_init(conf):
    seed = int(torch.randint(0, 2 ** 31 - 1, ()).item())
    return synth.make_state_dict(conf, seed=seed)


class Generator(_EngineOwner):
    """HiFi-GAN generator (hifigan/models.py:90-140).  ``Generator(h)`` takes the ``model``
    node of the config, like the reference.  forward: (B, 80, F) -> (B, 1, 256 F)."""

    def __init__(self, h, _init=None):
        super().__init__()
        self.h = h
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        self.upsample_factor = int(np.prod(h.upsample_rates))
        if str(h.resblock) != "1":
            raise NotImplementedError("only resblock '1' (config.yaml:87)")
        from .config import AttrDict
        self.config = AttrDict(model=h, n_mels=int(h.initial_channel), segment_size=32, n_vocab=1, n_speaker=1)
        if _init is None:
            full = AttrDict(model=h, n_mels=int(h.initial_channel), n_vocab=2, n_speaker=2)
            _init = {k[len("generator."):]: v for k, v in _seeded_init(full).items() if k.startswith("generator.")}
        for k, v in _init.items():
            _register(self, k, v.clone())
        self._register_load_state_dict_pre_hook(_legacy_weight_norm_hook, with_module=True)

    def _pack(self):
        return packing.pack_vocoder({"generator." + k: v for k, v in self.state_dict().items()}, self.h)

    @torch.no_grad()
    def forward(self, x):
        eng = self._engine()
        x = _prep(x, torch.float32, eng.device)
        with eng.call_lock:
            return eng.vocode(x, time_major=False, mel_lens_ptr=None)

    def remove_weight_norm(self):
        """hifigan/models.py:133-140 (broken on torch >= 2.1 in the reference):
        replaces every (g, v) pair by the folded ``weight``."""
        print('Removing weight norm...')
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        for mod, _, _ in synth.vocoder_conv_shapes(self.h):
            if mod + ".parametrizations.weight.original0" not in sd:
                continue
            w = packing.fold_weight_norm(sd, mod)
            holder = self
            for p in mod.split("."):
                holder = holder._modules[p]
            del holder._modules["parametrizations"]
            holder.register_parameter("weight", nn.Parameter(w.to(next(self.parameters()).device), requires_grad=False))
        self._mark_dirty()


class PromptTTS(_EngineOwner):
    """Acoustic model (model_open_source.py:14-163), inference branch only."""

    def __init__(self, config, _init=None):
        super().__init__()
        self.config = config
        if _init is None:
            _init = {k[len("am."):]: v for k, v in _seeded_init(config).items() if k.startswith("am.")}
        for k, v in _init.items():
            _register(self, k, v.clone())
        self.compat_padded_batch = False

    def _pack(self):
        sd = {"am." + k: v for k, v in self.state_dict().items()}
        packed = packing.pack_state_dict_am(sd, self.config)
        return packed

    @torch.no_grad()
    def forward(self, inputs_ling, input_lengths, inputs_speaker, inputs_style_embedding, inputs_content_embedding,
                mel_targets=None, output_lengths=None, pitch_targets=None, energy_targets=None, alpha=1.0):
        if mel_targets is not None:
            raise NotImplementedError("training-mode forward (teacher forcing) is out of scope for this engine")
        eng = self._engine()
        with eng.call_lock:
            return _am_forward(eng, inputs_ling, input_lengths, inputs_speaker, inputs_style_embedding,
                               inputs_content_embedding, not self.compat_padded_batch)[0]


def _am_forward(eng, inputs_ling, input_lengths, inputs_speaker, style, content, invariant):
    dev = eng.device
    ling = _prep(inputs_ling, torch.int64, dev)
    lens = _prep(input_lengths, torch.int64, dev)
    spk = _prep(inputs_speaker, torch.int64, dev)
    style = _prep(style, torch.float32, dev)
    content = _prep(content, torch.float32, dev)
    if ling.dim() != 2 or lens.dim() != 1 or lens.shape[0] != ling.shape[0] or spk.numel() != ling.shape[0]:
        raise RuntimeError("shape mismatch: inputs_ling %s, input_lengths %s, inputs_speaker %s"
                           % (tuple(ling.shape), tuple(lens.shape), tuple(spk.shape)))
    want = (ling.shape[0], int(eng.cfg.bert_dim))
    if tuple(style.shape) != want or tuple(content.shape) != want:
        raise RuntimeError("inputs_style_embedding %s / inputs_content_embedding %s must both be %s"
                           % (tuple(style.shape), tuple(content.shape), want))
    r = eng.acoustic(ling, lens, spk, style, content, invariant)
    out = {
        "mel_targets": None,
        "dec_outputs": r["mel"],
        "postnet_outputs": None,
        "pitch_predictions": r["pitch"].squeeze(),     # model_open_source.py:153
        "pitch_targets": None,
        "energy_predictions": r["energy"].squeeze(),   # :155
        "energy_targets": None,
        "log_duration_predictions": r["dur"],          # linear-domain integer frames despite the name (:157)
        "duration_targets": None,
        "input_lengths": input_lengths,
        "output_lengths": None,
        "log_p_attn": None,
        "bin_loss": None,
        "mel_lengths": r["mel_lens"],                  # extension: per-item frame counts (B,) int32
        "mel_lengths_host": r["mel_lens_host"],        # extension: the same on the host (read at the path's one sync)
    }
    return out, r


class JETSGenerator(_EngineOwner):
    """Joint acoustic model + vocoder (jets.py:26-71).

    ``compat_padded_batch`` (default False): with B > 1 the reference's padded forward leaks
    padding into the shorter items (decoder runs unmasked, convolutions run across pad frames;
    SURVEY.md s4 item 4).  By default every item of a batch is computed exactly like the
    reference's B=1 call for that item (what every reference caller runs); set the attribute
    to True to reproduce the literal padded-batch forward instead.  For B=1 both agree.
    """

    def __init__(self, config):
        super().__init__()
        self.upsample_factor = int(np.prod(config.model.upsample_rates))
        self.segment_size = config.segment_size
        init = _seeded_init(config)
        self.am = PromptTTS(config, _init={k[3:]: v for k, v in init.items() if k.startswith("am.")})
        self.generator = Generator(config.model, _init={k[10:]: v for k, v in init.items() if k.startswith("generator.")})
        self.config = config
        self.compat_padded_batch = False
        self._register_load_state_dict_pre_hook(_legacy_weight_norm_hook, with_module=True)

    def _pack(self):
        return packing.pack_state_dict(self.state_dict(), self.config)

    @torch.no_grad()
    def forward(self, inputs_ling, input_lengths, inputs_speaker, inputs_style_embedding, inputs_content_embedding,
                mel_targets=None, output_lengths=None, pitch_targets=None, energy_targets=None, alpha=1.0,
                cut_flag=True):
        if mel_targets is not None:
            raise NotImplementedError("training-mode forward (teacher forcing / random segments) is out of scope")
        eng = self._engine()
        invariant = not self.compat_padded_batch
        with eng.call_lock:
            return self._forward_locked(eng, invariant, inputs_ling, input_lengths, inputs_speaker, inputs_style_embedding, inputs_content_embedding)

    def reserve(self, batch=1, phonemes=256, frames=2048):
        """Serving set-up: pre-size the engine's workspace arena (current CUDA stream) for requests up to this shape, so that no
        forward allocates device memory afterwards.  Without it the arena simply grows when a larger request arrives (one allocation
        stall per new maximum)."""
        self._engine().reserve(batch, phonemes, frames)

    def _forward_locked(self, eng, invariant, inputs_ling, input_lengths, inputs_speaker, inputs_style_embedding, inputs_content_embedding):
        outputs, r = _am_forward(eng, inputs_ling, input_lengths, inputs_speaker, inputs_style_embedding,
                                 inputs_content_embedding, invariant)
        B = r["mel"].shape[0]
        mel_lens_ptr = (r["meta"].data_ptr() + 4 * B) if invariant else None
        # jets.py:62-66: z = dec_outputs.transpose(1, 2); wav = generator(z).  dec_outputs is already the
        # vocoder's time-major input layout: no transpose, no copy.
        wav = eng.vocode(r["mel"], time_major=True, mel_lens_ptr=mel_lens_ptr, ws=r["ws2"], n=r["n2"])
        outputs["wav_predictions"] = wav
        outputs["z_start_idxs"] = None
        outputs["segment_size"] = self.segment_size
        return outputs

    @torch.no_grad()
    def to_pcm16(self, wav):
        """The callers' ``wav * 32768 -> int16`` (inference_am_vocoder_joint.py:130-131) on the GPU: truncation toward zero
        like ``astype``.  Deviation: samples outside the int16 range SATURATE to [-32768, 32767] (tanh can round to exactly
        1.0 -> 32768) where numpy's cast wraps around; for |wav| < 1 the two agree bit for bit."""
        eng = self._engine()
        wav = _prep(wav, torch.float32, eng.device)
        pcm = torch.empty(wav.shape, dtype=torch.int16, device=eng.device)
        _abi.check(eng.lib.ev_wav_to_pcm16(wav.data_ptr(), pcm.data_ptr(), wav.numel(), eng._stream()))
        return pcm
