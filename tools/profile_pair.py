"""One ResBlock layer through the fused kernel (ev_op_resblock_gp) next to the two conv1d_gp launches it replaces (for ncu / A-B).
usage: python tools/profile_pair.py MODE C K DIL L [B] [reps]      MODE in {tf32, fp32, bf16, bf16x3}
Layer-granular bytes: fused = x in + out (+ residual re-read counted once: L2 hit); unfused = 5 activation passes."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emotivoice_b200 import _abi, build, layout, packing

build.build(verbose=False)
lib = _abi.load()
prec, C, K, dil, L = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
B = int(sys.argv[6]) if len(sys.argv) > 6 else 1
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
mode = {"tf32": 0, "fp32": 1, "bf16": 2, "bf16x3": 3}[prec]
bf = mode == 2
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
pack = packing.to_tc16x2_layout if mode == 3 else (packing.to_tc16_layout if bf else packing.to_tc_layout)
w1 = pack(torch.randn(K, C, C, generator=g) / math.sqrt(C * K)).to(dev)
w2 = pack(torch.randn(K, C, C, generator=g) / math.sqrt(C * K)).to(dev)
b1, b2 = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
x = layout.to_gp(torch.randn(B, L, C, generator=g), bf).to(dev)
xt, out = torch.empty_like(x), torch.empty_like(x)
flush = torch.empty(64 * 1024 * 1024, device=dev)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: t.data_ptr()


def fused():
    _abi.check(lib.ev_op_resblock_gp(p(x), p(w1), p(b1), p(w2), p(b2), mode, p(out), B, L, C, K, dil, None, 1, 0, 1.0, st))


def unfused():
    _abi.check(lib.ev_op_conv1d_gp(p(x), p(w1), mode, p(b1), None, p(xt), B, L, C, C, K, dil, 1, None, 1, 1, 0.1, 0, 1.0, st))
    _abi.check(lib.ev_op_conv1d_gp(p(xt), p(w2), mode, p(b2), p(x), p(out), B, L, C, C, K, 1, 1, None, 1, 1, 0.1, 0, 1.0, st))


res = {}
for name, fn in (("fused", fused), ("unfused", unfused)):
    ts = []
    for i in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(round(e0.elapsed_time(e1) * 1e3, 1))
    res[name] = ts
es = 2 if bf else 4
fl = 2 * 2.0 * B * L * C * C * K
bf_, bu = es * B * L * C * 2.0, es * B * L * C * 5.0
tf, tu = min(res["fused"][1:]), min(res["unfused"][1:])
print(json.dumps({"mode": prec, "C": C, "K": K, "dil": dil, "L": L, "B": B, "us": res, "fused_us": tf, "unfused_us": tu, "speedup": round(tu / tf, 3),
                  "fused_tflops": round(fl / tf / 1e6, 1), "fused_gbs_boundary": round(bf_ / tf / 1e3, 1), "unfused_gbs_layer": round(bu / tu / 1e3, 1)}))
