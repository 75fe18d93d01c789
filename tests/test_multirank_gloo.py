"""World-size-2 CPU test (gloo) of the N>1 host logic: disjoint game-id shards per rank and the end-of-run
reduce of statistics.  The per-rank work is done by the oracle here (no GPU in this test); the sum over ranks must
equal a single-process run over the union of the shards."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stats_for(first_id, n, seed):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kc_oracle
    recs, _, _ = kc_oracle.playout_run(5, 5, 4, seed, first_id, n, planes=False, threads=2)
    moved = recs[recs["movePos"] >= 0]
    fin = moved[(moved["status"] >> 8) & 1 == 1]
    w = (fin["status"] >> 9) & 3
    return [len(moved), len(moved), len(fin), int((w == 1).sum()), int((w == 2).sum()), int((w == 0).sum())]


def _worker(rank, world, port, n, seed, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from katacoffee_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    local = _stats_for(shard.first_game_id(rank), n, seed)
    total = shard.reduce_stats(local)
    slowest = shard.max_over_ranks(1.0 + rank)
    out.put((rank, local, total, slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_and_reduce():
    sys.path.insert(0, ROOT)
    from katacoffee_b200 import shard
    world, n, seed = 2, 300, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert shard.first_game_id(1) - shard.first_game_id(0) >= 1 << 32          # shards cannot overlap, refills included
    locals_ = [r[1] for r in res]
    assert res[0][2] == list(np.sum(locals_, axis=0))                           # rank 0 holds the sum
    assert all(r[3] == 2.0 for r in res)                                        # max over ranks visible everywhere
    assert locals_[0] != locals_[1]                                             # the shards really are different games
    assert locals_[0] == _stats_for(shard.first_game_id(0), n, seed)            # and deterministic
