"""Per-step view of bench.py's `value` loop (distinct corpus utterances, L2 flush between steps): device ms, frames, cudaMalloc / cudaFree
calls of torch's caching allocator during the step.   python tools/diag_value_loop.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emotivoice_b200 import synth
from emotivoice_b200.config import default_config
from emotivoice_b200.modules import JETSGenerator
conf = default_config(); dev = torch.device("cuda:0")
m = JETSGenerator(conf).to(dev); m.load_state_dict(synth.make_state_dict(conf)); m.eval()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
bs = [{k: v.to(dev) for k, v in synth.collate_utterances([synth.corpus_utterance(i)]).items()} for i in range(n)]
flush = torch.empty(64 * 1024 * 1024, device=dev)
for rnd in range(2):
    for i, b in enumerate(bs):
        flush.zero_()
        s0 = torch.cuda.memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = m(**b); e1.record(); e1.synchronize()
        s1 = torch.cuda.memory_stats()
        print("round %d step %2d  %6.2f ms  F=%4d  cudaMalloc +%d  cudaFree +%d  reserved %.0f MB" % (
            rnd, i, e0.elapsed_time(e1), out["dec_outputs"].shape[1], s1["num_device_alloc"] - s0["num_device_alloc"],
            s1["num_device_free"] - s0["num_device_free"], s1["reserved_bytes.all.current"] / 2**20), flush=True)
