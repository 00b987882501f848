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
