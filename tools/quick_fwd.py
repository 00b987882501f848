"""One B=1 forward + parity vs the committed fixture (debug helper)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from emotivoice_b200 import synth
from emotivoice_b200.config import default_config
from emotivoice_b200.modules import JETSGenerator
conf = default_config(); dev = torch.device("cuda:0")
m = JETSGenerator(conf).to(dev); m.load_state_dict(synth.make_state_dict(conf)); m.eval()
m.precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
b = {k: v.to(dev) for k, v in synth.make_batch([100], seed=synth.SEED).items()}
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "b1_t100.npz"))
for i in range(3):
    o = m(**b); torch.cuda.synchronize()
ts = []
for i in range(10):
    t0 = time.perf_counter(); o = m(**b); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
wav = torch.from_numpy(z["wav"])
err = float(((o["wav_predictions"].cpu() - wav).double().pow(2).mean().sqrt() / wav.double().pow(2).mean().sqrt()))
print("%s: median %.3f ms, dur equal %s, wav rms-rel %.2e" % (m.precision, 1e3 * sorted(ts)[5],
      bool(torch.equal(o["log_duration_predictions"].cpu(), torch.from_numpy(z["durations"]))), err), flush=True)
