"""N>1 host logic on CPU: world_size-2 gloo processes exercise the weight broadcast and the example gather."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from othellozero_b200 import dist as ozd
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    blob = np.arange(1000, dtype=np.float32) if rank == 0 else None
    t = ozd.broadcast_weights(blob, src=0)
    ok_w = bool((t.numpy() == np.arange(1000, dtype=np.float32)).all())
    g = 3 + rank
    rec = dict(n_moves=np.array([2] * g), winner=np.array([rank] * g),
               action=np.full((g, 64), 7 + rank, dtype=np.uint8), player=np.zeros((g, 64), dtype=np.uint8),
               black=np.full((g, 64), 100 + rank, dtype=np.uint64), white=np.full((g, 64), 200 + rank, dtype=np.uint64))
    allrows = ozd.gather_examples(ozd.pack_records(rec))
    ids = ozd.shard_game_ids(10, rank, world)
    q.put((rank, ok_w, allrows.shape, int(allrows[:, 0].sum()), ids.tolist()))
    dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    for rank, ok_w, shape, s, ids in res:
        assert ok_w
        assert shape == ((3 + 4) * 2, 3)
        assert s == 3 * 2 * 100 + 4 * 2 * 101
    assert res[0][4] == [0, 2, 4, 6, 8] and res[1][4] == [1, 3, 5, 7, 9]
