"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED
reference (``/root/reference``), and pins oracle/jets_oracle.py against it.

Run in the build container (the GPU box has no reference tree):

    python oracle/make_golden.py

For every case the reference's own ``JETSGenerator`` (jets.py:26) is instantiated
with the reference's own config.yaml, loaded (strict) with the seeded synthetic
state dict of ``emotivoice_b200.synth.make_state_dict`` and called exactly like
inference_am_vocoder_joint.py:120-129 calls it.  Inputs and outputs are stored;
the oracle restatement must reproduce them (asserted here, and again by
tests/test_oracle_golden.py on every run).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from emotivoice_b200.config import default_config          # noqa: E402
from emotivoice_b200 import synth                           # noqa: E402
from oracle import jets_oracle as O                          # noqa: E402
from oracle import refshim                                   # noqa: E402

CASES = {
    # name: (phoneme counts, input seed); one padded batch, literal reference batched forward
    "b1_t12": ([12], 1240),
    "b1_t50": ([50], 1242),          # BASELINE.json configs[0] shape
    "b1_t100": ([100], synth.SEED),  # BASELINE.json configs[1] shape == the bench.py workload
    "b3_padded": ([9, 23, 14], 1243),  # literal padded-batch semantics
}
VOC_CASE = ("voc_b2_f40", 2, 40)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    conf = default_config()
    sd = synth.make_state_dict(conf)
    JETS = refshim.import_reference_jets()
    ref = JETS(refshim.load_reference_config(conf.n_vocab, conf.n_speaker)).eval()
    ref.load_state_dict(sd, strict=True)
    meta = dict(state_dict_digest=synth.state_dict_digest(sd), torch=torch.__version__,
                numpy=np.__version__, seed=synth.SEED, cases={})
    for name, (lens, seed) in CASES.items():
        batch = synth.make_batch(lens, seed=seed)
        with torch.no_grad():
            r = ref(**{k: v.clone() for k, v in batch.items()})
        o = O.jets_forward(sd, conf, **batch)
        assert torch.equal(r["log_duration_predictions"], o["log_duration_predictions"]), name
        for k in ("dec_outputs", "wav_predictions"):
            err = (r[k] - o[k]).abs().max().item()
            assert err <= 1e-6 * max(1.0, r[k].abs().max().item()), (name, k, err)
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            **{k: v.numpy() for k, v in batch.items()},
            durations=r["log_duration_predictions"].numpy(),
            pitch=r["pitch_predictions"].reshape(len(lens), -1).numpy(),
            energy=r["energy_predictions"].reshape(len(lens), -1).numpy(),
            mel=r["dec_outputs"].numpy(),
            wav=r["wav_predictions"].numpy())
        meta["cases"][name] = dict(lens=lens, seed=seed, frames=int(r["dec_outputs"].shape[1]),
                                   mel_absmax=float(r["dec_outputs"].abs().max()),
                                   wav_absmax=float(r["wav_predictions"].abs().max()))
        print(name, lens, "F=%d" % r["dec_outputs"].shape[1], "oracle==reference OK")
    # vocoder only (Generator.forward, hifigan/models.py:115-131)
    name, B, Fr = VOC_CASE
    mel = synth.make_mel(B, Fr, seed=synth.SEED + 7)
    with torch.no_grad():
        w = ref.generator(mel)
    wo = O.vocoder(sd, conf.model, mel)
    assert (w - wo).abs().max().item() <= 1e-7, (w - wo).abs().max().item()
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), mel=mel.numpy(), wav=w.numpy())
    meta["cases"][name] = dict(batch=B, frames=Fr, seed=synth.SEED + 7, wav_absmax=float(w.abs().max()))
    print(name, "oracle==reference OK")
    # length regulator edge cases (GaussianUpsampling.forward, alignment.py:180-211), the reference module itself
    if refshim.REF_ROOT not in sys.path:
        sys.path.insert(0, refshim.REF_ROOT)
    from models.prompt_tts_modified.modules.alignment import GaussianUpsampling
    up = GaussianUpsampling()
    g = torch.Generator().manual_seed(99)
    hs = torch.randn(3, 6, 16, generator=g)
    valid = torch.tensor([[1, 1, 1, 1, 1, 1], [1, 1, 1, 0, 0, 0], [1, 0, 0, 0, 0, 0]], dtype=torch.bool)
    edge = {
        "all_zero": torch.zeros(3, 6, dtype=torch.int64),                                   # :187-191 -> every token gets 1 (pads included)
        "zero_tokens": torch.tensor([[0, 3, 0, 2, 0, 1], [2, 0, 0, 0, 0, 0], [4, 0, 0, 0, 0, 0]]),   # zero-duration tokens still get weight
        "one_item_zero": torch.tensor([[1, 2, 1, 1, 2, 1], [0, 0, 0, 0, 0, 0], [3, 0, 0, 0, 0, 0]]),  # batch sum != 0: the guard does NOT fire
    }
    arrays = {"hs": hs.numpy(), "valid": valid.numpy()}
    for name, ds in edge.items():
        with torch.no_grad():
            want = up(hs.clone(), ds.clone(), None, valid)
        got, mel_lens = O.gaussian_upsampling(hs.clone(), ds.clone(), valid)
        assert want.shape == got.shape and torch.equal(torch.nan_to_num(want, nan=7.0), torch.nan_to_num(got, nan=7.0)), name
        arrays["ds_" + name], arrays["out_" + name], arrays["mel_lens_" + name] = ds.numpy(), want.numpy(), mel_lens.numpy()
        print("upsample edge", name, tuple(want.shape), "oracle==reference OK")
    np.savez_compressed(os.path.join(out_dir, "upsample_edge.npz"), **arrays)
    with open(os.path.join(out_dir, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
