"""The oracle restatement (oracle/jets_oracle.py) must reproduce the fixtures that
oracle/make_golden.py generated from the UNMODIFIED reference modules.  This is what pins
the oracle; the GPU parity tests then compare the CUDA path with the oracle."""
import pytest
import torch

from conftest import load_golden, rel_max
from emotivoice_b200 import synth
from oracle import jets_oracle as O

KEYS = ("inputs_ling", "input_lengths", "inputs_speaker", "inputs_style_embedding", "inputs_content_embedding")


def test_synthetic_weights_are_reproducible(sd, golden_meta):
    assert synth.state_dict_digest(sd) == golden_meta["state_dict_digest"]
    assert len(sd) == 422 and sum(v.numel() for v in sd.values()) == 53277015


@pytest.mark.parametrize("name", ["b1_t12", "b1_t50", "b3_padded"])
def test_oracle_reproduces_reference_fixture(name, sd, conf, golden_meta):
    g = load_golden(name)
    o = O.jets_forward(sd, conf, **{k: g[k] for k in KEYS})
    assert torch.equal(o["log_duration_predictions"], g["durations"])
    assert o["dec_outputs"].shape == g["mel"].shape
    assert o["wav_predictions"].shape == g["wav"].shape
    assert o["wav_predictions"].shape[-1] == 256 * g["mel"].shape[1]
    assert rel_max(o["dec_outputs"], g["mel"]) <= 2e-6
    assert rel_max(o["wav_predictions"], g["wav"]) <= 2e-6
    B = g["inputs_ling"].shape[0]
    assert rel_max(o["pitch_predictions"].reshape(B, -1), g["pitch"]) <= 2e-6
    assert rel_max(o["energy_predictions"].reshape(B, -1), g["energy"]) <= 2e-6


def test_oracle_vocoder_fixture(sd, conf):
    g = load_golden("voc_b2_f40")
    w = O.vocoder(sd, conf.model, g["mel"])
    assert w.shape == g["wav"].shape == (2, 1, 40 * 256)
    assert rel_max(w, g["wav"]) <= 2e-6


def test_oracle_fp64_noise_floor(sd, conf):
    """fp32 reference vs fp64 run of the same algorithm: the oracle's own noise floor
    (~1e-6 relative); durations identical."""
    g = load_golden("b1_t12")
    o64 = O.jets_forward(sd, conf, **{k: g[k] for k in KEYS}, dtype=torch.float64)
    assert torch.equal(o64["log_duration_predictions"], g["durations"])
    assert rel_max(o64["dec_outputs"].float(), g["mel"]) <= 2e-5
    assert rel_max(o64["wav_predictions"].float(), g["wav"]) <= 2e-5


def test_padded_batch_is_not_batch_invariant_in_reference(sd, conf):
    """the literal padded forward leaks padding into shorter items,
    which is why the engine's default contract is 'each item == its B=1 call'."""
    g = load_golden("b3_padded")
    per = O.jets_forward_per_utterance(sd, conf, {k: g[k] for k in KEYS})
    lens = g["input_lengths"].tolist()
    longest = max(range(len(lens)), key=lambda i: lens[i])
    diffs = []
    for b, r in enumerate(per):
        Fb = r["dec_outputs"].shape[1]
        if not torch.equal(r["log_duration_predictions"][0], g["durations"][b, :lens[b]]):
            diffs.append(1.0)
            continue
        diffs.append(rel_max(g["mel"][b, :Fb], r["dec_outputs"][0]))
    assert diffs[longest] < 1e-4
    assert max(d for i, d in enumerate(diffs) if i != longest) > 1e-3


@pytest.mark.parametrize("case", ["all_zero", "zero_tokens", "one_item_zero"])
def test_length_regulator_edge_cases_match_reference_fixture(case):
    """GaussianUpsampling.forward (alignment.py:180-211) on the durations a bad predictor can emit: all zero (the batch-wide
    guard of :187-191 turns every duration, pads included, into 1), zero-duration tokens (they still receive weight), and one
    all-zero item inside a non-zero batch (the guard does not fire).  Fixture: the reference module itself (oracle/make_golden.py)."""
    g = load_golden("upsample_edge")
    out, mel_lens = O.gaussian_upsampling(g["hs"].clone(), g["ds_" + case].clone(), g["valid"])
    assert torch.equal(mel_lens, g["mel_lens_" + case])
    assert out.shape == g["out_" + case].shape
    assert torch.equal(torch.nan_to_num(out, nan=7.0), torch.nan_to_num(g["out_" + case], nan=7.0))
