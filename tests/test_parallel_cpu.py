"""N>1 host logic on CPU: world_size-2 gloo processes shard utterances, replicate a state dict from
rank 0 and gather per-utterance results back into order (no GPU, no data-path collective)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fish_speech_b200 import parallel as P

    utts = [f"utt{i}" for i in range(7)]
    mine = P.shard(utts, rank, world)
    sd = {"a.weight": torch.arange(6.0).view(2, 3), "b": torch.tensor([1, 2, 3])} if rank == 0 else None
    sd = P.broadcast_state_dict(sd, src=0)
    local = [(u, float(sd["a.weight"].sum()) + len(u)) for u in mine]
    allr = P.gather_objects(local)
    q.put((rank, mine, [u for u, _ in allr], sd["b"].tolist()))
    dist.destroy_process_group()


def test_shard_broadcast_gather_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    utts = [f"utt{i}" for i in range(7)]
    assert res[0][1] == utts[0::2] and res[1][1] == utts[1::2]
    for r in res:
        assert r[2] == utts and r[3] == [1, 2, 3]


def test_unshard_inverts_shard():
    from fish_speech_b200.parallel import shard, unshard

    for n in (0, 1, 5, 8):
        for world in (1, 2, 3, 8):
            items = list(range(n))
            assert unshard([shard(items, r, world) for r in range(world)]) == items
