"""Pins oracle/align_oracle.py (monotonic alignment search, durations, bin loss, per-token averaging: SURVEY.md s8f rank 4)
against fixtures generated from the reference's own numba functions (alignment.py:90-177) by oracle/make_golden_align.py.
Paths and durations are integers: bit-exact.  bin_loss / averages are float32 means: 1e-6."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import align_oracle as AO


@pytest.mark.parametrize("name", ["align_b3", "align_b2_ties", "align_b4_long"])
def test_alignment_oracle_matches_reference_fixture(name):
    g = {k: v.numpy() for k, v in load_golden(name).items()}
    tt, tf = g["text_lengths"].tolist(), g["feats_lengths"].tolist()
    ds, bl = AO.viterbi_decode(g["log_p_attn"], tt, tf)
    assert np.array_equal(ds, g["durations"])
    assert abs(float(bl) - float(g["bin_loss"])) <= 1e-6 * abs(float(g["bin_loss"]))
    for b in range(len(tt)):
        path = AO.monotonic_alignment_search(g["log_p_attn"][b, :tf[b], :tt[b]])
        assert np.array_equal(path, g["paths"][b, :tf[b]])
        assert path[0] == 0 and path[-1] == tt[b] - 1 and (np.diff(path) >= 0).all() and (np.diff(path) <= 1).all()      # monotonic, surjective
        assert ds[b].sum() == tf[b]
    assert np.abs(AO.average_by_duration(ds, g["xs"], tt, tf) - g["averaged"]).max() <= 1e-6


def test_alignment_module_oracle_matches_reference_fixture():
    """AlignmentModule.forward (alignment.py:33-56) restated in oracle/align_oracle.py vs the unmodified reference module's
    output (oracle/make_golden_align.py), with the product's host-side beta-binomial prior (the same scipy call as the
    reference's); get_segments vs the reference's segments incl. the t < segment_size zero padding."""
    import torch
    from emotivoice_b200 import synth
    from emotivoice_b200.align import AlignmentModule
    g = load_golden("alignmod_b3")
    sd = synth.make_alignment_state_dict(384, 80)
    tl, fl = g["text_lengths"], g["feats_lengths"]
    T = g["text"].shape[1]
    x_masks = torch.arange(T)[None, :] >= tl[:, None]
    prior = AlignmentModule(384, 80)._generate_prior
    lp = AO.alignment_module_forward(sd, g["text"], g["feats"], tl, fl, x_masks, prior_fn=prior)
    fin = torch.isfinite(g["log_p_attn"])
    assert torch.equal(fin, torch.isfinite(lp)) and (lp[fin] - g["log_p_attn"][fin]).abs().max() <= 1e-5
    assert np.array_equal(AO.get_segments(g["z"].numpy(), g["starts"].numpy(), 32), g["seg"].numpy())
    assert np.array_equal(AO.get_segments(g["z"].numpy()[:, :, :20], np.array([0, 3, 19]), 32), g["seg_short"].numpy())
