"""CPU model of the fused ResBlock-layer kernel's tiling (csrc/resblock_tc.cu): the same index expressions -- tile origin,
halo offsets, row-shifted tap views of the two operand tiles, the R = 128*MT - (k-1) rows a tile may store, the zero mask of
the intermediate outside [0, len) -- walked in numpy and compared with the layer computed directly.  It pins the arithmetic the
kernel's correctness rests on (the barrier protocol itself can only be exercised on the GPU), and the planner's invariants."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from emotivoice_b200 import _abi

BM = 128


def layer_direct(x, w1, b1, w2, b2, K, dil, n):
    """x (L, C) with n valid rows -> x + c2(lrelu(c1(lrelu(x)))) on the valid rows, zeros after (the B=1 semantics)."""
    xv = torch.from_numpy(x[:n].T[None])
    t = F.conv1d(F.leaky_relu(xv, 0.1), torch.from_numpy(w1).permute(2, 1, 0), torch.from_numpy(b1), padding=(K - 1) // 2 * dil, dilation=dil)
    y = F.conv1d(F.leaky_relu(t, 0.1), torch.from_numpy(w2).permute(2, 1, 0), torch.from_numpy(b2), padding=(K - 1) // 2) + xv
    out = np.zeros_like(x)
    out[:n] = y[0].T.numpy()
    return out


def layer_tiled(x, w1, b1, w2, b2, K, dil, n, MT):
    L, C = x.shape
    lrelu = lambda a: np.where(a > 0, a, a * np.float32(0.1))
    h2 = (K - 1) // 2
    h1 = h2 * dil
    R = BM * MT - (K - 1)
    rows_a1 = BM * MT + (K - 1) * dil
    rows2 = BM * MT + (K - 1)
    out = np.full_like(x, np.nan)
    for tm in range((L + R - 1) // R):
        t0 = tm * R
        if t0 >= n:                                     # padding tile: zeros
            out[t0:min(t0 + R, L)] = 0
            continue
        # producers: A1 row r <-> time t0 - h2 - h1 + r, zero outside [0, n), LeakyReLU applied while staging
        a1 = np.zeros((rows_a1, C), np.float32)
        for r in range(rows_a1):
            t = t0 - h2 - h1 + r
            if 0 <= t < n:
                a1[r] = lrelu(x[t])
        # c1: accumulator row r reads A1 rows r + j*dil
        acc1 = np.zeros((BM * MT, C), np.float32)
        for j in range(K):
            acc1 += a1[j * dil:j * dil + BM * MT] @ w1[j]
        # epilogue 1: A2 row r <-> time t0 - h2 + r; rows c1 never produced stay garbage
        a2 = np.full((rows2, C), np.nan, np.float32)
        for r in range(BM * MT):
            t = t0 - h2 + r
            a2[r] = lrelu(acc1[r] + b1) if 0 <= t < n else 0
        # c2: accumulator row rl reads A2 rows rl + j
        acc2 = np.zeros((BM * MT, C), np.float32)
        for j in range(K):
            acc2 += np.nan_to_num(a2[j:j + BM * MT], nan=1e30) @ w2[j]      # garbage rows poison only rows that are never stored
        for rl in range(R):
            row = t0 + rl
            if row < L:
                out[row] = (acc2[rl] + b2 + x[row]) if row < n else 0
    return out


@pytest.mark.parametrize("K,dil,MT,L,n", [(3, 1, 1, 300, 300), (11, 5, 1, 500, 431), (7, 3, 2, 900, 640), (11, 1, 4, 1300, 1300),
                                          (3, 5, 2, 260, 9), (11, 5, 2, 246 * 2, 246 * 2)])
def test_tiled_walk_equals_the_layer(K, dil, MT, L, n):
    rng = np.random.default_rng(K * 100 + dil * 10 + MT)
    C = 8
    x = rng.normal(size=(L, C)).astype(np.float32)
    x[n:] = 7.0                                          # rows past the item's length must never be read
    w1 = (rng.normal(size=(K, C, C)) / np.sqrt(C * K)).astype(np.float32)
    w2 = (rng.normal(size=(K, C, C)) / np.sqrt(C * K)).astype(np.float32)
    b1, b2 = rng.normal(size=C).astype(np.float32), rng.normal(size=C).astype(np.float32)
    want = layer_direct(x, w1, b1, w2, b2, K, dil, n)
    got = layer_tiled(x, w1, b1, w2, b2, K, dil, n, MT)
    assert not np.isnan(got).any()                       # every row of [0, L) is stored by exactly one tile
    assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max())


NAMES = "MT KBG a_stages b_stages groups tmem smem tiles R rows1_pad rows2_pad".split()


def _plan(lib, B, L, C, K, dil, mode):
    v = (ctypes.c_int * 11)()
    rc = lib.ev_debug_resblock_plan(B, L, C, K, dil, mode, v)
    return None if rc != 0 else dict(zip(NAMES, list(v)))


def test_planner_invariants_and_batch_independent_reduction_order():
    from emotivoice_b200 import build
    build.build(verbose=False)
    lib = _abi.load()
    seen = 0
    for C in (32, 64, 128):
        for K in (3, 7, 11):
            for dil in (1, 3, 5):
                for mode in (0, 1, 2):
                    plans = [_plan(lib, B, L, C, K, dil, mode) for B, L in ((1, 300), (1, 137472), (32, 20000), (3, 70001))]
                    if C == 128 and mode == 1:
                        assert all(p is None for p in plans)            # the 3xTF32 operand tile of c2 does not fit
                        continue
                    assert all(p is not None for p in plans)
                    seen += 1
                    assert len({p["KBG"] for p in plans}) == 1          # the only order-relevant parameter: shape-only
                    tcv = (ctypes.c_int * 11)()
                    for d in (dil, 1):                                  # == the K granules of both unfused convolutions
                        assert lib.ev_debug_tc_plan(1, 300, C, C, K, d, mode, 0, tcv) == 0 and tcv[2] == plans[0]["KBG"]
                    for p in plans:
                        assert p["groups"] <= p["a_stages"] <= 8 and 2 <= p["b_stages"] <= 8      # barrier-parity safety
                        assert p["smem"] <= 227 * 1024 and 2 * p["MT"] * C <= p["tmem"] <= 512
                        assert p["R"] == 128 * p["MT"] - (K - 1)
                        assert p["rows1_pad"] >= 128 * p["MT"] + (K - 1) * dil and p["rows2_pad"] >= 128 * p["MT"] + K - 1
    assert seen == 3 * 3 * 3 * 3 - 9
    assert _plan(lib, 1, 1000, 48, 3, 1, 1) is None and _plan(lib, 1, 1000, 256, 3, 1, 0) is None
