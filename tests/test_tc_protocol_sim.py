"""Discrete-event model of the mbarrier protocols of the two tcgen05 kernels (csrc/conv1d_tc.cu, csrc/conv1d_gp.cu).

Both deadlocks of round 1 were protocol bugs that a GPU can only show as a hang (producer groups running two ring phases
ahead of a parity wait; epilogue warps releasing an accumulator they never waited for).  This model replays the kernels' role
loops -- producers (in groups), weight loader, MMA issuer, epilogue warps -- as coroutines over barriers with the hardware's
semantics (arrival count, phase bit, `try_wait.parity(P)` passes iff the current phase parity != P; `tcgen05.commit` arrives
when every MMA issued before it has completed; MMAs read their operands as late as their completion) under many random
schedules, and checks: no deadlock, every slot holds the expected contents when it is read, no slot is overwritten before its
last reader is done.  The role loops are transcribed from the kernels; the planner (ring depths, groups) is the real one.
"""
import ctypes
import random

import pytest

from emotivoice_b200 import _abi

NEPI_WARPS = 8
NPWARPS = 6


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passes(self, parity):
        return (self.phase & 1) != parity


class Sim:
    """Cooperative scheduler.  Roles are generators yielding ('wait', bar, parity) | ('arrive', bar) | ('write', slot, tag)
    | ('read', slot, tag) | ('mma_read', slot, tag) | ('commit', bar).  MMA reads are deferred to the next commit's
    completion, and commits complete in order at a random later time: the most adversarial timing the hardware allows."""

    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.slots = {}
        self.roles = []
        self.inflight = []           # [(reads, bars)] groups of MMAs closed by a commit, oldest first
        self.open_reads = []

    def add(self, name, gen):
        self.roles.append([name, gen, None])

    def _complete_oldest(self):
        reads, bars = self.inflight.pop(0)
        for slot, tag in reads:
            assert self.slots.get(slot) == tag, "MMA read %s: holds %r, expected %r" % (slot, self.slots.get(slot), tag)
        for b in bars:
            b.arrive()

    def run(self, max_steps=2_000_000):
        live = list(self.roles)
        for _ in range(max_steps):
            if not live and not self.inflight:
                return
            runnable = [r for r in live if r[2] is None or r[2][0].passes(r[2][1])]
            if self.inflight and (not runnable or self.rng.random() < 0.3):
                self._complete_oldest()
                continue
            if not runnable:
                raise AssertionError("deadlock: %r" % ([(r[0], r[2][1], r[2][0].phase) for r in live],))
            role = self.rng.choice(runnable)
            name, gen = role[0], role[1]
            role[2] = None
            try:
                ev = next(gen)
            except StopIteration:
                live.remove(role)
                continue
            kind = ev[0]
            if kind == "wait":
                role[2] = (ev[1], ev[2])
            elif kind == "arrive":
                ev[1].arrive()
            elif kind == "write":
                self.slots[ev[1]] = ev[2]
            elif kind == "read":
                assert self.slots.get(ev[1]) == ev[2], "%s read %s: holds %r, expected %r" % (name, ev[1], self.slots.get(ev[1]), ev[2])
            elif kind == "mma_read":
                self.open_reads.append((ev[1], ev[2]))
            elif kind == "commit":
                if self.inflight and not self.open_reads:
                    self.inflight[-1][1].append(ev[1])      # back-to-back commits track the same MMAs
                else:
                    self.inflight.append((self.open_reads, [ev[1]]))
                    self.open_reads = []
        raise AssertionError("simulation did not finish")


# ------------------------------------------------------------------------------------------------------------------
# conv1d_tc.cu
# ------------------------------------------------------------------------------------------------------------------
def sim_conv(seed, tiles, n_cb, K, a_stages, b_stages, ngroups, epi_idle_warps=0, idle_warps_skip_wait=False):
    """tiles: list of booleans (True = active tile, False = padding tile that every role skips)."""
    sim = Sim(seed)
    wpg = NPWARPS // ngroups
    a_full = [Bar(wpg) for _ in range(a_stages)]            # one arrival per producer warp here (the kernel: per thread)
    a_empty = [Bar(1) for _ in range(a_stages)]
    b_full = [Bar(1) for _ in range(b_stages)]
    b_empty = [Bar(1) for _ in range(b_stages)]
    acc_full = [Bar(1), Bar(1)]
    acc_empty = [Bar(NEPI_WARPS), Bar(NEPI_WARPS)]

    def producer(grp, w):
        a_cnt = 0
        for ti, active in enumerate(tiles):
            if not active:
                continue
            for cb in range(n_cb):
                if a_cnt % ngroups == grp:
                    s = a_cnt % a_stages
                    yield ("wait", a_empty[s], ((a_cnt // a_stages) & 1) ^ 1)
                    if w == 0:
                        yield ("write", ("A", s), (ti, cb))
                    yield ("arrive", a_full[s])
                a_cnt += 1

    def loader():
        b_cnt = 0
        for ti, active in enumerate(tiles):
            if not active:
                continue
            for cb in range(n_cb):
                for j in range(K):
                    sb = b_cnt % b_stages
                    yield ("wait", b_empty[sb], ((b_cnt // b_stages) & 1) ^ 1)
                    yield ("write", ("B", sb), (ti, cb, j))
                    yield ("arrive", b_full[sb])           # expect_tx + complete_tx of the bulk copy
                    b_cnt += 1

    def mma():
        a_cnt = b_cnt = tile_cnt = 0
        for ti, active in enumerate(tiles):
            if not active:
                continue
            buf = tile_cnt & 1
            yield ("wait", acc_empty[buf], ((tile_cnt >> 1) & 1) ^ 1)
            for cb in range(n_cb):
                sa = a_cnt % a_stages
                yield ("wait", a_full[sa], (a_cnt // a_stages) & 1)
                for j in range(K):
                    sb = b_cnt % b_stages
                    yield ("wait", b_full[sb], (b_cnt // b_stages) & 1)
                    yield ("mma_read", ("A", sa), (ti, cb))
                    yield ("mma_read", ("B", sb), (ti, cb, j))
                    yield ("commit", b_empty[sb])
                    b_cnt += 1
                yield ("commit", a_empty[sa])
                a_cnt += 1
            yield ("write", ("ACC", buf), ti)                # (written as the MMAs complete; conservatively at issue)
            yield ("commit", acc_full[buf])
            tile_cnt += 1

    def epilogue(w):
        tile_cnt = 0
        for ti, active in enumerate(tiles):
            if not active:
                continue
            buf = tile_cnt & 1
            if not (idle_warps_skip_wait and w < epi_idle_warps):
                yield ("wait", acc_full[buf], (tile_cnt >> 1) & 1)  # idle warps (no columns) wait too: the round-1 fix
            if w >= epi_idle_warps:
                yield ("read", ("ACC", buf), ti)
            yield ("arrive", acc_empty[buf])
            tile_cnt += 1

    for g in range(ngroups):
        for w in range(wpg):
            sim.add("producer%d.%d" % (g, w), producer(g, w))
    sim.add("loader", loader())
    sim.add("mma", mma())
    for w in range(NEPI_WARPS):
        sim.add("epilogue%d" % w, epilogue(w))
    sim.run()


def _tc_plan(lib, B, L, Cin, Cout, K, dil, mode, ksplit=0):
    v = (ctypes.c_int * 11)()
    assert lib.ev_debug_tc_plan(B, L, Cin, Cout, K, dil, mode, ksplit, v) == 0
    return dict(zip("BN MT KBG a_stages b_stages groups ksplit tmem smem tiles rows_pad".split(), list(v)))


@pytest.fixture(scope="module")
def lib():
    from emotivoice_b200 import build
    build.build(verbose=False)
    return _abi.load()


@pytest.mark.parametrize("shape", [(1, 537, 384, 1152, 1, 1), (1, 537, 1536, 384, 3, 1), (1, 34368, 128, 128, 11, 5),
                                   (1, 137472, 32, 32, 3, 1), (1, 4296, 256, 256, 7, 3), (1, 100, 384, 384, 3, 1)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_conv_protocol_with_the_real_plans(lib, shape, mode):
    B, L, Cin, Cout, K, dil = shape
    pl = _tc_plan(lib, B, L, Cin, Cout, K, dil, mode)
    cpg = 8 if mode == 2 else 4
    n_cb = -(-Cin // (cpg * pl["KBG"]))
    for seed in range(6):
        rng = random.Random(seed)
        tiles = [rng.random() > 0.2 for _ in range(rng.randint(1, 5))]       # tiles of ONE persistent CTA, some of them padding
        sim_conv(seed, tiles, n_cb, K, pl["a_stages"], pl["b_stages"], pl["groups"], epi_idle_warps=4 if Cout <= 32 else 0)


def test_conv_protocol_model_catches_the_round1_bugs():
    """The model is only worth something if it fails on the two protocols that hung the GPU in round 1."""
    with pytest.raises(AssertionError):                       # 6 producer groups on a 2-deep ring: parity cannot tell phases apart
        for seed in range(20):
            sim_conv(seed, [True, True, True], n_cb=12, K=3, a_stages=2, b_stages=4, ngroups=6)
    with pytest.raises(AssertionError):                       # column-less epilogue warps handing back an accumulator they never waited for
        for seed in range(50):
            sim_conv(seed, [True] * 6, n_cb=2, K=3, a_stages=2, b_stages=4, ngroups=2, epi_idle_warps=4, idle_warps_skip_wait=True)


# ------------------------------------------------------------------------------------------------------------------
# conv1d_gp.cu: A loader (bulk copies) -> a_full -> transform warps (in place) -> a_ready -> MMA -> a_empty -> A loader
# ------------------------------------------------------------------------------------------------------------------
NTW = 4


def sim_gp(seed, tiles, n_cb, K, a_stages, b_stages, n_work_items=2, skip_idle_wait=False):
    """tiles: list of booleans (True = active tile, False = padding tile that every role skips).  n_work_items: epilogue
    work items (MT * BN/32) per lane quadrant; warp `half` of a quadrant takes items half, half+2, ...: with one item the
    second warp of each quadrant has nothing to read but must still follow the accumulator phases."""
    sim = Sim(seed)
    taps = (lambda ti: K[ti]) if isinstance(K, (list, tuple)) else (lambda ti: K)     # grouped launch: the taps differ from tile to tile
    a_full = [Bar(1) for _ in range(a_stages)]
    a_ready = [Bar(NTW) for _ in range(a_stages)]           # one arrival per transform warp here (the kernel: per thread)
    a_empty = [Bar(1) for _ in range(a_stages)]
    b_full = [Bar(1) for _ in range(b_stages)]
    b_empty = [Bar(1) for _ in range(b_stages)]
    acc_full = [Bar(1), Bar(1)]
    acc_empty = [Bar(NEPI_WARPS), Bar(NEPI_WARPS)]

    def aloader():
        a_cnt = 0
        for ti, active in enumerate(tiles):
            if not active:
                continue
            for cb in range(n_cb):
                s = a_cnt % a_stages
                yield ("wait", a_empty[s], ((a_cnt // a_stages) & 1) ^ 1)
                yield ("write", ("A", s), ("raw", ti, cb))       # expect_tx + the bulk copies' complete_tx
                yield ("arrive", a_full[s])
                a_cnt += 1

    def xform(w):
        a_cnt = 0
        for ti, active in enumerate(tiles):
            if not active:
                continue
            for cb in range(n_cb):
                s = a_cnt % a_stages
                yield ("wait", a_full[s], (a_cnt // a_stages) & 1)
                yield ("read", ("A", s), ("raw", ti, cb) if w == 0 else sim.slots.get(("A", s)))
                if w == 0:
                    yield ("write", ("A", s), ("op", ti, cb))
                yield ("arrive", a_ready[s])
                a_cnt += 1

    def bloader():
        b_cnt = 0
        for ti, active in enumerate(tiles):
            if not active:
                continue
            for cb in range(n_cb):
                for j in range(taps(ti)):
                    sb = b_cnt % b_stages
                    yield ("wait", b_empty[sb], ((b_cnt // b_stages) & 1) ^ 1)
                    yield ("write", ("B", sb), (ti, cb, j))
                    yield ("arrive", b_full[sb])
                    b_cnt += 1

    def mma():
        a_cnt = b_cnt = tile_cnt = 0
        for ti, active in enumerate(tiles):
            if not active:
                continue
            buf = tile_cnt & 1
            yield ("wait", acc_empty[buf], ((tile_cnt >> 1) & 1) ^ 1)
            for cb in range(n_cb):
                sa = a_cnt % a_stages
                yield ("wait", a_ready[sa], (a_cnt // a_stages) & 1)
                for j in range(taps(ti)):
                    sb = b_cnt % b_stages
                    yield ("wait", b_full[sb], (b_cnt // b_stages) & 1)
                    yield ("mma_read", ("A", sa), ("op", ti, cb))
                    yield ("mma_read", ("B", sb), (ti, cb, j))
                    yield ("commit", b_empty[sb])
                    b_cnt += 1
                yield ("commit", a_empty[sa])
                a_cnt += 1
            yield ("write", ("ACC", buf), ti)
            yield ("commit", acc_full[buf])
            tile_cnt += 1

    def epilogue(w):
        half = w >> 2
        has_work = half < n_work_items
        tile_cnt = 0
        for ti, active in enumerate(tiles):
            if not active:
                continue
            buf = tile_cnt & 1
            if has_work or not skip_idle_wait:
                yield ("wait", acc_full[buf], (tile_cnt >> 1) & 1)
            if has_work:
                yield ("read", ("ACC", buf), ti)
            yield ("arrive", acc_empty[buf])
            tile_cnt += 1

    sim.add("aloader", aloader())
    for w in range(NTW):
        sim.add("xform%d" % w, xform(w))
    sim.add("bloader", bloader())
    sim.add("mma", mma())
    for w in range(NEPI_WARPS):
        sim.add("epilogue%d" % w, epilogue(w))
    sim.run()


def _gp_plan(lib, B, L, Cin, Cout, K, dil, rate, mode):
    v = (ctypes.c_int * 11)()
    assert lib.ev_debug_gp_plan(B, L, Cin, Cout, K, dil, rate, mode, v) == 0, lib.ev_last_error()
    return dict(zip("BN MT KBG a_stages b_stages ntw planes tmem smem tiles rows_pad".split(), list(v)))


GP_SHAPES = [(1, 537, 80, 512, 7, 1, 1), (1, 537, 512, 2048, 3, 1, 8), (1, 4296, 256, 256, 11, 5, 1), (1, 34368, 128, 128, 11, 1, 1),
             (1, 34368, 128, 128, 3, 1, 2), (1, 68736, 64, 64, 7, 3, 1), (1, 137472, 32, 32, 3, 1, 1), (32, 262144, 32, 32, 11, 5, 1),
             (8, 65536, 128, 128, 11, 5, 1), (128, 32768, 256, 256, 7, 1, 1)]


@pytest.mark.parametrize("shape", GP_SHAPES)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gp_plans_respect_the_hardware_limits_and_the_protocol(lib, shape, mode):
    B, L, Cin, Cout, K, dil, rate = shape
    pl = _gp_plan(lib, B, L, Cin, Cout, K, dil, rate, mode)
    cpg = 8 if mode == 2 else 4
    assert pl["smem"] <= 227 * 1024 and 2 <= pl["a_stages"] <= 8 and 1 <= pl["b_stages"] <= 8
    assert pl["tmem"] in (32, 64, 128, 256, 512) and 2 * pl["MT"] * pl["BN"] <= pl["tmem"]
    assert pl["BN"] % 32 == 0 and Cout % pl["BN"] == 0 and pl["planes"] == (2 if mode == 1 else 1)
    rows = 128 * pl["MT"] + (K - 1) * dil
    assert pl["rows_pad"] >= rows and pl["rows_pad"] % 8 == 0
    # one stage's transaction count must fit the mbarrier tx-count field (2^20 - 1 bytes)
    assert pl["KBG"] * pl["rows_pad"] * 16 < (1 << 20) and pl["planes"] * pl["KBG"] * pl["BN"] * 16 < (1 << 20)
    n_cb = -(-Cin // (cpg * pl["KBG"]))
    items = pl["MT"] * (pl["BN"] // 32)
    for seed in range(5):
        rng = random.Random(seed)
        tiles = [rng.random() > 0.2 for _ in range(rng.randint(1, 5))]
        sim_gp(seed, tiles, min(n_cb, 6), K, pl["a_stages"], pl["b_stages"], n_work_items=items)


def test_gp_summation_order_parameters_do_not_depend_on_batch_or_length(lib):
    """KBG (channels per pipeline stage) fixes the order of each output element's reduction: it must be a function of the
    layer shape only, or a batch would not be bitwise equal to its items' B=1 runs."""
    for (Cin, Cout, K, dil, rate) in [(80, 512, 7, 1, 1), (512, 2048, 3, 1, 8), (256, 256, 11, 5, 1), (128, 128, 7, 3, 1), (32, 32, 11, 5, 1)]:
        for mode in (0, 1, 2):
            kbgs = {_gp_plan(lib, B, L, Cin, Cout, K, dil, rate, mode)["KBG"] for (B, L) in [(1, 64), (1, 5000), (3, 70000), (32, 300000)]}
            assert len(kbgs) == 1, (Cin, Cout, K, dil, mode, kbgs)


def _gp_group_plan(lib, Ks, dils, B, L, Cin, Cout, mode):
    n = len(Ks)
    v = (ctypes.c_int * 11)()
    IA = ctypes.c_int * n
    assert lib.ev_debug_gp_group_plan(n, IA(*Ks), IA(*dils), B, L, Cin, Cout, mode, v) == 0, lib.ev_last_error()
    return dict(zip("BN MT KBG a_stages b_stages ntw planes tmem smem tiles rows_pad".split(), list(v)))


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(1, 4296, 256), (1, 34368, 128), (2, 3000, 128), (3, 900, 64)])
def test_gp_grouped_launch_plans(lib, shape, mode):
    """The grouped launch (three convolutions of one shape in a kernel) must use each member's own K granules per stage (the
    reduction order: bitwise equality with the members' own launches), size its stages for the widest halo, and carry every
    member's tiles; and it must fit the same hardware limits as a single launch."""
    B, L, C = shape
    Ks, dils = (3, 7, 11), (1, 3, 5)
    g = _gp_group_plan(lib, Ks, dils, B, L, C, C, mode)
    solo = [_gp_plan(lib, B, L, C, C, K, d, 1, mode) for K, d in zip(Ks, dils)]
    assert {p["KBG"] for p in solo} == {g["KBG"]}
    assert g["smem"] <= 227 * 1024 and 2 * g["MT"] * g["BN"] <= g["tmem"] <= 512
    assert g["rows_pad"] >= 128 * g["MT"] + (11 - 1) * 5 and g["rows_pad"] % 8 == 0
    tiles_one = B * -(-L // (128 * g["MT"])) * (C // g["BN"])
    assert g["tiles"] == 3 * tiles_one
    n_cb = -(-C // ((8 if mode == 2 else 4) * g["KBG"]))
    for seed in range(3):
        rng = random.Random(seed)
        n = rng.randint(2, 6)
        sim_gp(seed, [True] * n, min(n_cb, 4), [rng.choice(Ks) for _ in range(n)], g["a_stages"], g["b_stages"], n_work_items=g["MT"] * (g["BN"] // 32))


@pytest.mark.parametrize("mode", [0, 2, 3])
@pytest.mark.parametrize("C,mul", [(256, 8), (128, 64)])
def test_gp_grouped_launch_plans_over_utterance_lengths(lib, C, mul, mode):
    """The grouped convolution plan over every length a batch-1 step can have: members' own K granules, all tiles carried, limits kept."""
    for F in range(40, 2401, 17):
        L = F * mul
        g = _gp_group_plan(lib, (3, 7, 11), (1, 3, 5), 1, L, C, C, mode)
        assert g["KBG"] == _gp_plan(lib, 1, L, C, C, 11, 5, 1, mode)["KBG"] == _gp_plan(lib, 1, L, C, C, 3, 1, 1, mode)["KBG"]
        assert g["tiles"] == 3 * -(-L // (128 * g["MT"])) * (C // g["BN"]) and g["MT"] in (1, 2, 4)
        assert g["smem"] <= 227 * 1024 and 2 * g["MT"] * g["BN"] <= g["tmem"] <= 512 and g["rows_pad"] >= 128 * g["MT"] + 50


def test_gp_grouped_launch_plan_fills_the_machine_at_batch_1(lib):
    """HiFi-GAN stage 1 at batch 1 in the fp32 mode: 3 x 68 one-accumulator tiles (204 > 148 SMs) instead of 3 x 34 two-accumulator
    ones -- the plan is picked by simulating the round-robin deal, where the k = 11 member's double tile would be the critical path."""
    g = _gp_group_plan(lib, (3, 7, 11), (1, 3, 5), 1, 4296, 256, 256, 3)
    assert g["MT"] == 1 and g["tiles"] == 204


def test_gp_grouped_launch_protocol():
    """A grouped launch (conv1d_gp_group: the same-index convolutions of three parallel ResBlocks in one kernel) changes the number of
    taps -- weight stages per activation stage -- from tile to tile; the weight loader and the MMA issuer derive it from the same
    tile decode, so the rings stay in step for any mix and any ring depth."""
    for seed in range(30):
        rng = random.Random(seed)
        n = rng.randint(2, 7)
        tiles = [rng.random() > 0.15 for _ in range(n)]
        taps = [rng.choice((3, 7, 11)) for _ in range(n)]
        sim_gp(seed, tiles, rng.randint(1, 4), taps, rng.randint(2, 4), rng.randint(2, 8), n_work_items=rng.choice((1, 2, 4)))


@pytest.mark.parametrize("mode", [0, 2, 3])
@pytest.mark.parametrize("C,mul", [(64, 128), (32, 256)])
def test_resblock_gp_grouped_launch_plans_over_utterance_lengths(lib, C, mul, mode):
    """Host logic of the grouped fused launch over the lengths a batch-1 step can have (every frame count from 40 to 2400 in steps of
    13): members in launch order heaviest first, each with ITS OWN rows per tile R = 128 MT - (K - 1) and row-tile count, tile
    ranges contiguous and complete, stage sizes for the widest halo, hardware limits respected.  A wrong first-tile index or tile
    count would make the roles of the kernel walk different tile sequences (a hang), so this is checked here, on the CPU."""
    Ks, dils = (3, 7, 11), (1, 3, 5)
    IA = ctypes.c_int * 3
    seen_mt = set()
    for F in range(40, 2401, 13):
        L = F * mul
        v = (ctypes.c_int * 16)()
        rc = lib.ev_debug_resblock_gp_group_plan(3, IA(*Ks), IA(*dils), 1, L, C, mode, v)
        solo = []
        for K, d in zip(Ks, dils):
            w = (ctypes.c_int * 11)()
            solo.append(list(w) if lib.ev_debug_resblock_gp_plan(1, L, C, K, d, mode, w) == 0 else None)
        if any(x is None or x[0] < 2 for x in solo):
            assert rc != 0          # a member that would not be fused on its own is never grouped
            continue
        assert rc == 0, lib.ev_last_error()
        mt, kbg, total, rows1_pad, rows2_pad, smem, tmem = list(v)[:7]
        members = [tuple(v[7 + 3 * i: 10 + 3 * i]) for i in range(3)]
        seen_mt.add(mt)
        assert mt in (2, 4) and {x[1] for x in solo} == {kbg}
        assert [m[0] for m in members] == [11, 7, 3]                      # heaviest first
        t0 = 0
        for K, tiles_m, tile0 in members:
            R = 128 * mt - (K - 1)
            assert tiles_m == -(-L // R) and tile0 == t0
            t0 += tiles_m
        assert total == t0
        assert rows1_pad >= 128 * mt + 10 * 5 and rows2_pad >= 128 * mt + 10 and rows1_pad % 8 == 0 and rows2_pad % 8 == 0
        assert smem <= 227 * 1024 and 4 * mt * C <= tmem <= 512
    assert seen_mt


def test_resblock_gp_grouped_launch_protocol():
    """Grouped fused-ResBlock launch: consecutive tiles of a CTA may belong to layers with different taps; the weight loader streams
    w1 of the NEXT tile's layer, then w2 of the current one, exactly as the MMA issuer consumes them."""
    for seed in range(30):
        rng = random.Random(seed)
        n = rng.randint(1, 6)
        sim_pair(seed, n, rng.randint(1, 3), [rng.choice((3, 7, 11)) for _ in range(n)], rng.randint(2, 4), rng.randint(2, 8))


def test_gp_protocol_model_is_sensitive():
    """The model must fail when an epilogue warp without work items releases an accumulator set it never waited for
    (the round-1 hang), which is why the kernel's idle warps still wait on acc_full."""
    with pytest.raises(AssertionError):
        for seed in range(40):
            sim_gp(seed, [True] * 5, n_cb=1, K=1, a_stages=2, b_stages=2, n_work_items=1, skip_idle_wait=True)


# ------------------------------------------------------------------------------------------------------------------
# resblock_gp.cu: C1(0) | C1(i+1), C2(i); epi1 writes the xt tile between them; separate epilogue warp groups
# ------------------------------------------------------------------------------------------------------------------
def sim_pair(seed, n_tiles, n_cb, K, a_stages, b_stages):
    sim = Sim(seed)
    taps = (lambda ti: K[ti]) if isinstance(K, (list, tuple)) else (lambda ti: K)     # grouped launch: the taps differ from tile to tile
    a_full = [Bar(1) for _ in range(a_stages)]
    a_ready = [Bar(NTW) for _ in range(a_stages)]
    a_empty = [Bar(1) for _ in range(a_stages)]
    b_full = [Bar(1) for _ in range(b_stages)]
    b_empty = [Bar(1) for _ in range(b_stages)]
    acc1_full, acc1_empty = [Bar(1), Bar(1)], [Bar(4), Bar(4)]
    acc2_full, acc2_empty = [Bar(1), Bar(1)], [Bar(4), Bar(4)]
    a2_full, a2_empty = Bar(4), Bar(1)                  # one arrival per epi1 warp here (the kernel: per thread)

    def order():                                         # the MMA issuer's / weight loader's schedule
        if n_tiles:
            yield (1, 0)
        for i in range(n_tiles):
            if i + 1 < n_tiles:
                yield (1, i + 1)
            yield (2, i)

    def aloader():
        a_cnt = 0
        for ti in range(n_tiles):
            for cb in range(n_cb):
                s = a_cnt % a_stages
                yield ("wait", a_empty[s], ((a_cnt // a_stages) & 1) ^ 1)
                yield ("write", ("X", s), ("raw", ti, cb))
                yield ("arrive", a_full[s])
                a_cnt += 1

    def xform(w):
        a_cnt = 0
        for ti in range(n_tiles):
            for cb in range(n_cb):
                s = a_cnt % a_stages
                yield ("wait", a_full[s], (a_cnt // a_stages) & 1)
                if w == 0:
                    yield ("read", ("X", s), ("raw", ti, cb))
                    yield ("write", ("X", s), ("op", ti, cb))
                yield ("arrive", a_ready[s])
                a_cnt += 1

    def bloader():
        b_cnt = 0
        for which, ti in order():
            for cb in range(n_cb):
                for j in range(taps(ti)):
                    sb = b_cnt % b_stages
                    yield ("wait", b_empty[sb], ((b_cnt // b_stages) & 1) ^ 1)
                    yield ("write", ("B", sb), (which, ti, cb, j))
                    yield ("arrive", b_full[sb])
                    b_cnt += 1

    def mma():
        a_cnt = b_cnt = 0
        for which, ti in order():
            buf = ti & 1
            if which == 1:
                yield ("wait", acc1_empty[buf], ((ti >> 1) & 1) ^ 1)
                for cb in range(n_cb):
                    sa = a_cnt % a_stages
                    yield ("wait", a_ready[sa], (a_cnt // a_stages) & 1)
                    for j in range(taps(ti)):
                        sb = b_cnt % b_stages
                        yield ("wait", b_full[sb], (b_cnt // b_stages) & 1)
                        yield ("mma_read", ("X", sa), ("op", ti, cb))
                        yield ("mma_read", ("B", sb), (1, ti, cb, j))
                        yield ("commit", b_empty[sb])
                        b_cnt += 1
                    yield ("commit", a_empty[sa])
                    a_cnt += 1
                yield ("write", ("ACC1", buf), ti)
                yield ("commit", acc1_full[buf])
            else:
                yield ("wait", a2_full, ti & 1)
                yield ("wait", acc2_empty[buf], ((ti >> 1) & 1) ^ 1)
                for cb in range(n_cb):
                    for j in range(taps(ti)):
                        sb = b_cnt % b_stages
                        yield ("wait", b_full[sb], (b_cnt // b_stages) & 1)
                        yield ("mma_read", ("A2",), ti)
                        yield ("mma_read", ("B", sb), (2, ti, cb, j))
                        yield ("commit", b_empty[sb])
                        b_cnt += 1
                yield ("write", ("ACC2", buf), ti)
                yield ("commit", a2_empty)
                yield ("commit", acc2_full[buf])

    def epi1(w):
        for ti in range(n_tiles):
            buf = ti & 1
            yield ("wait", acc1_full[buf], (ti >> 1) & 1)
            yield ("wait", a2_empty, (ti & 1) ^ 1)
            yield ("read", ("ACC1", buf), ti)
            if w == 0:
                yield ("write", ("A2",), ti)
            yield ("arrive", a2_full)
            yield ("arrive", acc1_empty[buf])

    def epi2(w):
        for ti in range(n_tiles):
            buf = ti & 1
            yield ("wait", acc2_full[buf], (ti >> 1) & 1)
            yield ("read", ("ACC2", buf), ti)
            yield ("arrive", acc2_empty[buf])

    sim.add("aloader", aloader())
    for w in range(NTW):
        sim.add("xform%d" % w, xform(w))
    sim.add("bloader", bloader())
    sim.add("mma", mma())
    for w in range(4):
        sim.add("epi1_%d" % w, epi1(w))
        sim.add("epi2_%d" % w, epi2(w))
    sim.run()


@pytest.mark.parametrize("C,K,dil", [(32, 3, 1), (32, 11, 5), (64, 7, 3), (64, 11, 5), (128, 11, 1)])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_resblock_gp_protocol_with_the_real_plans(lib, C, K, dil, mode):
    v = (ctypes.c_int * 11)()
    assert lib.ev_debug_resblock_gp_plan(1, 137472, C, K, dil, mode, v) == 0, lib.ev_last_error()
    pl = dict(zip("MT KBG a_stages b_stages ntw tmem smem tiles R rows1_pad rows2_pad".split(), list(v)))
    assert pl["smem"] <= 227 * 1024 and 4 * pl["MT"] * C <= pl["tmem"] <= 512 and pl["a_stages"] >= 2 and pl["b_stages"] >= 2
    assert pl["R"] == 128 * pl["MT"] - (K - 1) and pl["rows1_pad"] >= 128 * pl["MT"] + (K - 1) * dil and pl["rows2_pad"] >= 128 * pl["MT"] + K - 1
    cpg = 8 if mode == 2 else 4
    n_cb = -(-C // (cpg * pl["KBG"]))
    for seed in range(5):
        sim_pair(seed, random.Random(seed).randint(1, 5), n_cb, K, pl["a_stages"], pl["b_stages"])


def test_resblock_gp_kbg_matches_the_unfused_kernel(lib):
    """The fused layer must reduce in the same order as the two launches it replaces: same channels per pipeline stage."""
    for C, K, dil in [(32, 3, 1), (32, 11, 5), (64, 7, 3), (64, 11, 5), (128, 11, 5)]:
        for mode in (0, 1, 2, 3):
            v, u = (ctypes.c_int * 11)(), (ctypes.c_int * 11)()
            assert lib.ev_debug_resblock_gp_plan(4, 50000, C, K, dil, mode, v) == 0
            assert lib.ev_debug_gp_plan(4, 50000, C, C, K, dil, 1, mode, u) == 0
            assert v[1] == u[2], (C, K, dil, mode)
