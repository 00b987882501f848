"""Launch one conv shape through the C ABI a few times (for ncu / timing).
usage: python tools/profile_conv.py MODE C K DIL L [B] [reps]   MODE in {ffma, tf32, fp32, bf16};  C may be "Cin:Cout"."""
import math
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emotivoice_b200 import _abi, packing, build

build.build(verbose=False)
lib = _abi.load()
mode, K, dil, L = sys.argv[1], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
Cin, Cout = (int(v) for v in sys.argv[2].split(":")) if ":" in sys.argv[2] else (int(sys.argv[2]), int(sys.argv[2]))
C = Cout
B = int(sys.argv[6]) if len(sys.argv) > 6 else 1
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randn(B, L, Cin, generator=g).to(dev)
w = torch.randn(K, Cin, Cout, generator=g) / math.sqrt(Cin * K)
wd = (packing.to_tc16_layout(w) if mode == "bf16" else (packing.to_tc_layout(w) if mode != "ffma" else w)).to(dev)
b = torch.randn(C, generator=g).to(dev)
res = torch.randn(B, L, C, generator=g).to(dev)
out = torch.empty(B, L, C, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
st = torch.cuda.current_stream().cuda_stream
ts = []
for i in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if mode == "ffma":
        _abi.check(lib.ev_op_conv1d(x.data_ptr(), wd.data_ptr(), b.data_ptr(), 0, res.data_ptr(), out.data_ptr(), B, L, Cin, Cout, K, dil,
                                    None, 1, 1, 0.1, 0, 0, 1.0, st))
    else:
        _abi.check(lib.ev_op_conv1d_tc(x.data_ptr(), wd.data_ptr(), {"fp32": 1, "tf32": 0, "bf16": 2}[mode], b.data_ptr(), 0, res.data_ptr(),
                                       out.data_ptr(), B, L, Cin, Cout, K, dil, None, 1, 1, 0.1, 0, 0, 1.0, None, 0, st))
    e1.record()
    e1.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
fl = 2.0 * B * L * Cin * Cout * K
print("%s C=%d:%d K=%d dil=%d L=%d B=%d: %s us  -> %.1f TFLOP/s (best)" % (mode, Cin, Cout, K, dil, L, B, ["%.1f" % t for t in ts], fl / min(ts) / 1e6))
