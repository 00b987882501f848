"""Training-mode alignment helpers on the GPU with the reference's function names and signatures
(models/prompt_tts_modified/modules/alignment.py:124-177): ``viterbi_decode`` and ``average_by_duration``.

The reference copies ``log_p_attn`` to the host and runs a numba loop per sample in the middle of every training step
(model_open_source.py:114-118); here one kernel launch handles the batch and nothing leaves the device.  Paths and durations are
bit-exact with the reference (the kernel restates its float64 dynamic programme); ``bin_loss`` and the averages are float32
means (1e-6).  Everything else of the training forward (AlignmentModule's convolutions, losses, autograd) is out of scope.
"""
import torch

from . import _abi


def _check(t, name):
    if t.device.type != "cuda":
        raise RuntimeError("%s must be a CUDA tensor: emotivoice_b200 has no CPU path" % name)


def viterbi_decode(log_p_attn, text_lengths, feats_lengths, return_path=False):
    """log_p_attn (B, T_feats, T_text) float32 -> (ds (B, T_text) float32, bin_loss 0-dim float32)  [+ path (B, T_feats) int32]."""
    _check(log_p_attn, "log_p_attn")
    lib = _abi.load()
    dev = log_p_attn.device
    lp = log_p_attn.detach().to(torch.float32).contiguous()
    B, F, T = lp.shape
    tl = text_lengths.to(device=dev, dtype=torch.int64).contiguous()
    fl = feats_lengths.to(device=dev, dtype=torch.int64).contiguous()
    path = torch.empty((B, F), dtype=torch.int32, device=dev)
    ds = torch.empty((B, T), dtype=torch.float32, device=dev)
    loss = torch.empty((B,), dtype=torch.float32, device=dev)
    ws = torch.empty((B * F * T,), dtype=torch.uint8, device=dev)
    _abi.check(lib.ev_op_mas(lp.data_ptr(), tl.data_ptr(), fl.data_ptr(), B, F, T, path.data_ptr(), ds.data_ptr(), loss.data_ptr(),
                             ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    bin_loss = loss.sum() / B                      # alignment.py:141: the per-item means, averaged over the batch
    return (ds, bin_loss, path) if return_path else (ds, bin_loss)


def average_by_duration(ds, xs, text_lengths, feats_lengths):
    """ds (B, T_text), xs (B, T_feats) -> (B, T_text) float32: mean of xs over each token's frames (0 for empty tokens)."""
    _check(ds, "ds")
    lib = _abi.load()
    dev = ds.device
    d = ds.detach().to(torch.float32).contiguous()
    x = xs.detach().to(device=dev, dtype=torch.float32).contiguous()
    B, T = d.shape
    F = x.shape[1]
    tl = text_lengths.to(device=dev, dtype=torch.int64).contiguous()
    fl = feats_lengths.to(device=dev, dtype=torch.int64).contiguous()
    out = torch.empty((B, T), dtype=torch.float32, device=dev)
    _abi.check(lib.ev_op_average_by_duration(d.data_ptr(), x.data_ptr(), tl.data_ptr(), fl.data_ptr(), B, F, T, out.data_ptr(),
                                             torch.cuda.current_stream(dev).cuda_stream))
    return out
