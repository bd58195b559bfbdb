#!/usr/bin/env python
"""One full training iteration (BASELINE.json configs[4]; reference: main.py:71-148):

    self-play sharded by game  ->  C2 gather of example records  ->  train step (rank 0, PyTorch autograd)
    ->  C1 broadcast of the new weights  ->  arena: new net vs old net (main.py:103-148 semantics)

Prints one JSON line with the wall time of every phase.  Single GPU:  python tools/full_iteration.py
Multi GPU:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/full_iteration.py
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--board", type=int, default=8)
    ap.add_argument("--channels", type=int, default=512)
    ap.add_argument("--games", type=int, default=1024, help="self-play games per GPU")
    ap.add_argument("--sims", type=int, default=100)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--max-examples", type=int, default=32768)
    ap.add_argument("--arena-games", type=int, default=64)
    ap.add_argument("--arena-sims", type=int, default=25)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from othellozero_b200 import arena, dist as ozd, net as oznet, selfplay, train
    from othellozero_b200 import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, C = args.board, args.channels
    t = {}

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    old = oznet.B200NNet((n, n), C, device=local, max_batch=max(args.games, args.arena_games), seed=0)
    # ---- self-play, sharded by global game id ------------------------------------------------------------------
    sync(); t0 = time.perf_counter()
    sp = selfplay.SelfPlay(n, old, 1.0, max_games=args.games, num_simulations=args.sims, device=local, seed=1)
    ids = np.arange(rank * args.games, (rank + 1) * args.games, dtype=np.uint64)
    rec = sp.play(args.games, 1.0, 0.9, game_ids=ids)
    ctr = sp.engine.counters()
    sp.close()
    sync(); t["selfplay_s"] = time.perf_counter() - t0
    # ---- C2: gather packed example records -----------------------------------------------------------------------
    t0 = time.perf_counter()
    packed = ozd.pack_records(rec)
    allrows = ozd.gather_examples(packed) if world > 1 else packed
    sync(); t["gather_s"] = time.perf_counter() - t0
    # ---- train on rank 0 (Net/NNet.py:53-68) ---------------------------------------------------------------------
    t0 = time.perf_counter()
    new_blob = None
    hist = None
    if rank == 0:
        rows = np.asarray(allrows[:args.max_examples])
        meta = rows[:, 2].astype(np.int64)
        act, player, win = meta & 0xFF, (meta >> 8) & 0xFF, (meta >> 16) & 0xFF
        boards = selfplay.bits_to_boards(rows[:, 0], rows[:, 1], n)
        pols = np.zeros((rows.shape[0], n, n)); pols[np.arange(rows.shape[0]), act >> 3, act & 7] = 1
        examples = list(zip(boards, pols, np.where(win == player, 1, -1).tolist()))
        new_blob, hist = train.train_blob(old.blob, examples, n, C, epochs=args.epochs, device=f"cuda:{local}")
    sync(); t["train_s"] = time.perf_counter() - t0
    # ---- C1: broadcast the new weights ---------------------------------------------------------------------------
    t0 = time.perf_counter()
    if world > 1:
        wt = ozd.broadcast_weights(new_blob, src=0)
        new_blob = wt.cpu().numpy()
    new = oznet.B200NNet((n, n), C, device=local, max_batch=args.arena_games, blob=new_blob)
    sync(); t["broadcast_load_s"] = time.perf_counter() - t0
    # ---- arena: new (BLACK) vs old (WHITE) and the reverse, main.py:103-148 --------------------------------------
    t0 = time.perf_counter()
    half = args.arena_games // 2
    r1 = arena.pit(n, new, old, args.arena_sims, 1, n_games=half, device=local)
    r2 = arena.pit(n, old, new, args.arena_sims, 1, n_games=half, device=local)
    new_wins = int((r1["winner"] == 0).sum() + (r2["winner"] == 1).sum())
    sync(); t["arena_s"] = time.perf_counter() - t0
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"config": vars(args), "n_gpus": world, "phases": t, "selfplay_sims": ctr["sims"],
                          "selfplay_sims_per_s_per_gpu": ctr["sims"] / t["selfplay_s"], "examples": int(allrows.shape[0]),
                          "train_history": hist, "arena_new_wins": new_wins, "arena_games": 2 * half}))


if __name__ == "__main__":
    main()
