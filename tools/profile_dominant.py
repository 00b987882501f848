"""Launch bench.py's dominant launch (bench.dominant_launch: the grouped stage-2 ResBlock convolutions of the batch-1 step) a few
times, for ncu:   ncu --set full -k regex:conv1d_gp -s 2 -c 1 python tools/profile_dominant.py [precision] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from emotivoice_b200 import _abi, build

build.build(verbose=False)
lib = _abi.load()
dev = torch.device("cuda:0")
precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 511
d = bench.dominant_launch(lib, dev, frames, precision)
flush = torch.empty(64 * 1024 * 1024, device=dev)
ts = []
for i in range(5):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _abi.check(d["call"]())
    e1.record()
    e1.synchronize()
    ts.append(round(e0.elapsed_time(e1) * 1e3, 1))
print({"kernel": d["kname"], "us": ts, "tflops": round(d["flops"] / min(ts[1:]) / 1e6, 1), "flops": d["flops"], "alg_bytes": d["alg_bytes"]})
