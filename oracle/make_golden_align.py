"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/align_*.npz from the reference's own numba functions
(models/prompt_tts_modified/modules/alignment.py:90-177: _monotonic_alignment_search via viterbi_decode, average_by_duration)
and pins oracle/align_oracle.py against them.  Run in the build container:  python oracle/make_golden_align.py
Inputs are log-softmax rows like AlignmentModule.forward produces (alignment.py:47) -- one case quantised so that ties between
the two predecessor cells are frequent (the tie rule decides the integer path)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import align_oracle as AO      # noqa: E402
from oracle import refshim                 # noqa: E402

CASES = {
    # name: (text lengths, feats lengths, seed, quantise)
    "align_b3": ([17, 9, 25], [80, 33, 140], 7001, False),
    "align_b2_ties": ([12, 30], [64, 200], 7002, True),
    "align_b4_long": ([100, 57, 3, 120], [537, 260, 3, 600], 7003, False),     # incl. T_mel == T_inp and the bench utterance's shape
}


def main():
    if refshim.REF_ROOT not in sys.path:
        sys.path.insert(0, refshim.REF_ROOT)
    from models.prompt_tts_modified.modules import alignment as R
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, (tt, tf, seed, quant) in CASES.items():
        rng = np.random.default_rng(seed)
        B, T, F = len(tt), max(tt), max(tf)
        score = rng.normal(size=(B, F, T)).astype(np.float32) * 2.0
        for b in range(B):
            score[b, :, tt[b]:] = -np.inf                                  # x_masks (alignment.py:43-45)
        lp = torch.log_softmax(torch.from_numpy(score), dim=-1)
        if quant:
            lp = torch.round(lp * 4) / 4
        tl, fl = torch.tensor(tt), torch.tensor(tf)
        ds, bin_loss = R.viterbi_decode(lp, tl, fl)
        xs = torch.from_numpy(rng.normal(size=(B, F)).astype(np.float32))
        avg = R.average_by_duration(ds, xs, tl, fl)
        paths = np.full((B, F), -1, np.int32)
        for b in range(B):
            paths[b, :tf[b]] = R._monotonic_alignment_search(lp[b, :tf[b], :tt[b]].numpy())
        o_ds, o_bl = AO.viterbi_decode(lp.numpy(), tt, tf)
        o_avg = AO.average_by_duration(o_ds, xs.numpy(), tt, tf)
        assert np.array_equal(o_ds, ds.numpy()) and abs(float(o_bl) - float(bin_loss)) <= 1e-6 * abs(float(bin_loss)), name
        assert np.abs(o_avg - avg.numpy()).max() <= 1e-6, name
        for b in range(B):
            assert np.array_equal(AO.monotonic_alignment_search(lp[b, :tf[b], :tt[b]].numpy()), paths[b, :tf[b]]), name
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), log_p_attn=lp.numpy(), text_lengths=np.asarray(tt, np.int64),
                            feats_lengths=np.asarray(tf, np.int64), xs=xs.numpy(), paths=paths, durations=ds.numpy(),
                            bin_loss=np.float32(bin_loss), averaged=avg.numpy())
        print(name, "ok: bin_loss %.6f, durations sum %s" % (float(bin_loss), ds.sum(1).tolist()))


def make_module_fixture():
    """AlignmentModule.forward + get_random_segments of the UNMODIFIED reference on seeded inputs -> tests/golden/alignmod_b3.npz."""
    if refshim.REF_ROOT not in sys.path:
        sys.path.insert(0, refshim.REF_ROOT)
    from models.prompt_tts_modified.modules import alignment as R
    from models.hifigan import get_random_segments as RS
    from emotivoice_b200 import synth
    adim, odim = 384, 80
    mod = R.AlignmentModule(adim, odim).eval()
    mod.load_state_dict(synth.make_alignment_state_dict(adim, odim), strict=True)       # seeded: the fixture need not carry the weights
    tt, tf = [23, 9, 31], [120, 40, 187]
    B, T, F = len(tt), max(tt), max(tf)
    rng = np.random.default_rng(8001)
    text = torch.from_numpy(rng.normal(size=(B, T, adim)).astype(np.float32))
    feats = torch.from_numpy((rng.normal(size=(B, F, odim)) * 1.2).astype(np.float32))
    tl, fl = torch.tensor(tt), torch.tensor(tf)
    x_masks = torch.arange(T)[None, :] >= tl[:, None]                         # True = pad (model_open_source.py:164-173)
    with torch.no_grad():
        lp = mod(text, feats, tl, fl, x_masks)
    sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    o = AO.alignment_module_forward(sd, text, feats, tl, fl, x_masks, prior_fn=mod._generate_prior)
    fin = torch.isfinite(lp)
    assert torch.equal(fin, torch.isfinite(o)) and (lp[fin] - o[fin]).abs().max() <= 1e-5, "oracle restatement differs from the reference module"
    torch.manual_seed(99)
    z = torch.from_numpy(rng.normal(size=(B, odim, F)).astype(np.float32))
    seg, starts, size = RS.get_random_segments(z, fl, 32)
    assert np.array_equal(AO.get_segments(z.numpy(), starts.numpy(), 32), seg.numpy())
    short = RS.get_segments(z[:, :, :20], torch.tensor([0, 3, 19]), 32)        # t < segment_size: zero padded
    arrays = {"text": text.numpy(), "feats": feats.numpy(), "text_lengths": tl.numpy().astype(np.int64), "feats_lengths": fl.numpy().astype(np.int64),
              "log_p_attn": lp.numpy(), "z": z.numpy(), "seg": seg.numpy(), "starts": starts.numpy().astype(np.int64), "seg_short": short.numpy()}
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "alignmod_b3.npz"), **arrays)
    print("alignmod_b3 ok: log_p_attn", tuple(lp.shape), "starts", starts.tolist())


if __name__ == "__main__":
    main()
    make_module_fixture()
