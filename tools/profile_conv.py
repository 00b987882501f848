"""Launch one conv shape through the C ABI a few times (for ncu / timing).

usage: python tools/profile_conv.py MODE C K DIL L [B] [reps]
  MODE in {ffma, tf32, fp32, bf16, bf16x3}    time-major kernels (conv1d_tm / conv1d_tc: what the acoustic model runs); env KSPLIT=n splits K
          {gp:tf32, gp:fp32, gp:bf16, gp:bf16x3}   granule-planar kernel (conv1d_gp), the vocoder's default path
                                              (gp:fp32 = 3xTF32; gp:bf16x3 = the fp32 mode's default emulation)
  C may be "Cin:Cout" (or "Cin:Cout:rate" for the polyphase ConvTranspose1d form, gp only).
Prints per-launch times (L2 flushed before each), algorithmic TFLOP/s and layer-granular GB/s (in + out + residual, weights once)."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emotivoice_b200 import _abi, build, layout, packing

build.build(verbose=False)
lib = _abi.load()
mode, K, dil, L = sys.argv[1], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
cs = [int(v) for v in sys.argv[2].split(":")]
Cin, Cout, rate = (cs[0], cs[0], 1) if len(cs) == 1 else ((cs[0], cs[1], 1) if len(cs) == 2 else tuple(cs))
B = int(sys.argv[6]) if len(sys.argv) > 6 else 1
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
gp = mode.startswith("gp:")
prec = mode[3:] if gp else mode
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randn(B, L, Cin, generator=g)
w = torch.randn(K, Cin, Cout, generator=g) / math.sqrt(Cin * K)
wd = (packing.to_tc16x2_layout(w) if prec == "bf16x3" else (packing.to_tc16_layout(w) if prec == "bf16" else (packing.to_tc_layout(w) if prec != "ffma" else w))).to(dev)
b = torch.randn(Cout, generator=g).to(dev)
coutR = Cout // rate
res = torch.randn(B, L * rate, coutR, generator=g) if rate == 1 else None
esize = 2 if (gp and prec == "bf16") else 4
if gp:
    xd = layout.to_gp(x, prec == "bf16").to(dev)
    rd = layout.to_gp(res, prec == "bf16").to(dev) if res is not None else None
    out = torch.empty_like(layout.to_gp(torch.zeros(B, L * rate, coutR), prec == "bf16")).to(dev)
else:
    xd, rd, out = x.to(dev), res.to(dev), torch.empty(B, L, Cout, device=dev)
ksplit = int(os.environ.get("KSPLIT", "0"))
ws = torch.empty(ksplit * B * (L + 256) * Cout, device=dev) if ksplit > 1 and not gp else None
flush = torch.empty(64 * 1024 * 1024, device=dev)
st = torch.cuda.current_stream().cuda_stream
ptr = lambda t: None if t is None else t.data_ptr()
ts = []
warm = os.environ.get("WARM", "none")      # none: everything cold (L2 flushed); w: weights re-read after the flush (L2 warm); all: no flush
for i in range(reps):
    if warm != "all":
        flush.zero_()
    if warm == "w":
        wd.view(torch.uint8).sum().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if gp:
        _abi.check(lib.ev_op_conv1d_gp(ptr(xd), ptr(wd), {"tf32": 0, "fp32": 1, "bf16": 2, "bf16x3": 3}[prec], ptr(b), ptr(rd), ptr(out), B, L, Cin, Cout, K, dil, rate,
                                       None, 1, 1, 0.1, 0, 1.0, st))
    elif mode == "ffma":
        _abi.check(lib.ev_op_conv1d(ptr(xd), ptr(wd), ptr(b), 0, ptr(rd), ptr(out), B, L, Cin, Cout, K, dil, None, 1, 1, 0.1, 0, 0, 1.0, st))
    else:
        _abi.check(lib.ev_op_conv1d_tc(ptr(xd), ptr(wd), {"fp32": 1, "tf32": 0, "bf16": 2, "bf16x3": 3}[mode], ptr(b), 0, ptr(rd), ptr(out), B, L, Cin, Cout, K, dil,
                                       None, 1, 1, 0.1, 0, 0, 1.0, ptr(ws), 0 if ws is None else ws.numel(), st))
    e1.record()
    e1.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
fl = 2.0 * B * L * Cin * Cout * K
by = esize * B * L * (Cin + Cout * (2 if res is not None else 1)) + 4.0 * K * Cin * Cout
best = min(ts[1:]) if len(ts) > 1 else ts[0]
print(json.dumps({"warm": warm, "mode": mode, "Cin": Cin, "Cout": Cout, "rate": rate, "K": K, "dil": dil, "L": L, "B": B, "us": [round(t, 1) for t in ts],
                  "best_us": round(best, 1), "tflops": round(fl / best / 1e6, 1), "layer_gbs": round(by / best / 1e3, 1)}))
