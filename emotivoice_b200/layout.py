"""Granule-planar activation layout of the vocoder (csrc/conv1d_gp.cu): a (B, L, C) tensor is stored
``[b][C / cpg][l][cpg]`` in 16-byte granules -- cpg = 4 fp32 or 8 bf16 channels.  Host-side converters for the
tests and tools; the engine converts on the device (ev_op_to_gp) and never materialises the time-major form."""
import torch


def to_gp(x_tm, bf16=False):
    """(B, L, C) fp32 time-major -> GP tensor (B, C/cpg, L, cpg), fp32 or bf16."""
    B, L, C = x_tm.shape
    cpg = 8 if bf16 else 4
    assert C % cpg == 0
    g = x_tm.reshape(B, L, C // cpg, cpg).permute(0, 2, 1, 3).contiguous()
    return g.to(torch.bfloat16) if bf16 else g


def from_gp(x_gp):
    """GP tensor (B, G, L, cpg) -> (B, L, C) fp32 time-major."""
    B, G, L, cpg = x_gp.shape
    return x_gp.float().permute(0, 2, 1, 3).reshape(B, L, G * cpg).contiguous()
