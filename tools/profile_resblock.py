"""One ResBlock layer through the C ABI: the fused kernel (ev_op_resblock_pair) next to the two launches it replaces
(for timing and for ncu).  Algorithmic traffic per layer: unfused 5 activation passes (x in, xt out, xt in, x residual, out),
fused 2 (x in -- also the residual --, out); FLOPs 2 * 2*L*C*C*K in both.

usage: python tools/profile_resblock.py MODE C K DIL L [B] [reps]      MODE in {tf32, fp32, bf16}
       e.g. the vocoder stages of the bench utterance:  fp32 32 11 5 137472 ; fp32 64 7 3 68736 ; tf32 128 11 5 34368
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emotivoice_b200 import _abi, build, packing

build.build(verbose=False)
lib = _abi.load()
mode, C, K, dil, L = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
B = int(sys.argv[6]) if len(sys.argv) > 6 else 1
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
split3 = {"fp32": 1, "tf32": 0, "bf16": 2}[mode]
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randn(B, L, C, generator=g).to(dev)
pack = packing.to_tc16_layout if mode == "bf16" else packing.to_tc_layout
w1 = pack(torch.randn(K, C, C, generator=g) / math.sqrt(C * K)).to(dev)
w2 = pack(torch.randn(K, C, C, generator=g) / math.sqrt(C * K)).to(dev)
b1, b2 = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
xt, ref, out = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
flush = torch.empty(64 * 1024 * 1024, device=dev)
st = torch.cuda.current_stream().cuda_stream


def unfused():
    _abi.check(lib.ev_op_conv1d_tc(x.data_ptr(), w1.data_ptr(), split3, b1.data_ptr(), 0, None, xt.data_ptr(), B, L, C, C, K, dil, None, 1,
                                   _abi.ACT_LRELU, 0.1, _abi.ACT_NONE, _abi.ACC_STORE, 1.0, None, 0, st))
    _abi.check(lib.ev_op_conv1d_tc(xt.data_ptr(), w2.data_ptr(), split3, b2.data_ptr(), 0, x.data_ptr(), ref.data_ptr(), B, L, C, C, K, 1, None, 1,
                                   _abi.ACT_LRELU, 0.1, _abi.ACT_NONE, _abi.ACC_STORE, 1.0, None, 0, st))


def fused():
    _abi.check(lib.ev_op_resblock_pair(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), split3, out.data_ptr(),
                                       B, L, C, K, dil, None, 1, _abi.ACC_STORE, 1.0, st))


def timed(fn):
    ts = []
    for _ in range(reps + 1):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts[1:])


flops = 2 * 2.0 * B * L * C * C * K
bytes_act = 4.0 * B * L * C
res = {"mode": mode, "C": C, "K": K, "dil": dil, "L": L, "B": B}
t_u = timed(unfused)
res.update(unfused_us=t_u, unfused_tflops=flops / t_u / 1e6, unfused_gbs_algorithmic=5 * bytes_act / t_u / 1e3)
try:
    t_f = timed(fused)
    torch.cuda.synchronize()
    res.update(fused_us=t_f, fused_tflops=flops / t_f / 1e6, fused_gbs_algorithmic=2 * bytes_act / t_f / 1e3, speedup=t_u / t_f,
               bitwise_equal=bool(torch.equal(out, ref)))
except _abi.EvError as e:
    res["fused"] = "unsupported: %s" % e
print(json.dumps(res))
