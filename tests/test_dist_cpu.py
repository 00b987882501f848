"""Host-side multi-GPU logic on CPU: LPT sharding and the one-time weight broadcast over a
world_size-2 gloo group (the GPU box uses the same code over NCCL)."""
import os
import socket

import torch
import torch.multiprocessing as mp

from emotivoice_b200 import runner, synth
from emotivoice_b200.config import default_config


def test_plan_shards_is_balanced_and_complete():
    import numpy as np
    rng = np.random.default_rng(0)
    lens = rng.integers(20, 201, size=1000).tolist()
    for world in (1, 2, 4, 8):
        shards = runner.plan_shards(lens, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(1000))
        loads = [sum(lens[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lens)
        assert shards == runner.plan_shards(lens, world)       # deterministic
    batches = runner.bucket_batches(shards[0], lens, 32)
    assert sum(len(b) for b in batches) == len(shards[0])
    assert all(lens[b[0]] >= lens[b[-1]] for b in batches)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    conf = default_config(n_vocab=20, n_speaker=6)
    sd = synth.make_state_dict(conf, seed=99) if rank == 0 else None
    got = runner.broadcast_state_dict(sd, conf, torch.device("cpu"), src=0)
    q.put((rank, synth.state_dict_digest(got), len(got)))
    dist.barrier()
    dist.destroy_process_group()


def test_weight_broadcast_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    conf = default_config(n_vocab=20, n_speaker=6)
    want = synth.state_dict_digest(synth.make_state_dict(conf, seed=99))
    assert res[0][1] == want and res[1][1] == want and res[0][2] == 422


def test_corpus_utterances_do_not_depend_on_corpus_size_or_sharding():
    """bench.py's multi-GPU claim -- every utterance's PCM is bit-identical at every world size -- rests on utterance i being the
    same whatever the corpus size / shard: each has its own generator seeded by (seed, index)."""
    import numpy as np
    a = synth.corpus_utterance(37)
    b = synth.corpus_utterance(37, seed=synth.SEED)
    assert np.array_equal(a["ids"], b["ids"]) and a["speaker"] == b["speaker"] and np.array_equal(a["style"], b["style"])
    assert 20 <= len(a["ids"]) <= 200 and a["ids"][0] == 1 and a["ids"][-1] == 1
    lens = synth.corpus_lengths(64)
    assert lens == [len(synth.corpus_utterance(i)["ids"]) for i in range(64)]
    en, zh = synth._ids_from(synth.EN_ID_RANGES), synth._ids_from(synth.ZH_ID_RANGES)
    assert set(synth.corpus_utterance(10)["ids"][1:-1].tolist()) <= set(en.tolist())        # even index: EN symbols
    assert set(synth.corpus_utterance(11)["ids"][1:-1].tolist()) <= set(zh.tolist())        # odd index: ZH symbols
    assert not (set(en.tolist()) & set(zh.tolist())) and max(en.max(), zh.max()) <= 416     # 417-501 are unused placeholders
    fixed = synth.corpus_utterance(5, n_phonemes=100)
    assert len(fixed["ids"]) == 100
    batch = synth.collate_utterances([synth.corpus_utterance(i) for i in (3, 4)])
    assert batch["inputs_ling"].shape == (2, max(lens[3], lens[4])) and batch["input_lengths"].tolist() == [lens[3], lens[4]]


def test_digest_combination_is_independent_of_sharding():
    import numpy as np
    pcm = {i: np.random.default_rng(i).integers(-3000, 3000, size=100 + i).astype(np.int16) for i in range(12)}
    pairs = [(i, runner.utterance_digest(p)) for i, p in pcm.items()]
    whole = runner.combine_digests(pairs)
    for world in (2, 4):
        shards = runner.plan_shards([len(p) for p in pcm.values()], world)
        gathered = [pr for s in reversed(shards) for pr in pairs if pr[0] in s]          # ranks report in any order
        assert runner.combine_digests(gathered) == whole
    pcm[3][0] += 1
    assert runner.combine_digests([(i, runner.utterance_digest(p)) for i, p in pcm.items()]) != whole


def test_bench_shards_give_every_rank_the_same_number_of_steps():
    """bench.py deals world * (warmup + steps) equal-length utterances over the ranks with plan_shards (weak scaling)."""
    for world in (1, 2, 4, 8):
        n = world * 25
        shards = runner.plan_shards(synth.corpus_lengths(n, n_phonemes=100), world)
        assert [len(s) for s in shards] == [25] * world and sorted(i for s in shards for i in s) == list(range(n))
