#!/usr/bin/env python
"""Per-phase throughput of one complete batch of self-play games (the bench's e2e leg), sampled every STEP engine steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from othellozero_b200 import engine as E, net as oznet

G, C, sims = int(os.environ.get("G", "4096")), 512, 100
STEP = int(os.environ.get("STEP", "200"))
eng = E.Engine(8, max_games=G, nodes_per_game=sims * 61 + 64, prior_mode=E.PRIOR_NET, eval_cache_log2=int(os.environ.get("CACHE", "24")))
eng.load_weights(oznet.init_weights(8, C, seed=0), C)
sb, sw, sp = bench.synthetic_starts(E, G, 1, 1 << 24, 0)
ids = np.arange(1 << 24, (1 << 24) + G, dtype=np.uint64)
eng.selfplay_begin(G, sims, 1.0, 0.9, -1, sb, sw, sp, ids)
prev = eng.counters(); t_prev = time.perf_counter(); t0 = t_prev
k = 0
while True:
    active = eng.selfplay_run(STEP)
    torch.cuda.synchronize()
    now = time.perf_counter(); c = eng.counters()
    ds = c["sims"] - prev["sims"]; dn = c["nodes"] - prev["nodes"]
    de = dn - (c["cache_hits"] - prev["cache_hits"]) - (c["cache_aliases"] - prev["cache_aliases"])
    dt = now - t_prev
    k += STEP
    print(f"steps {k:6d} t {now - t0:6.2f}s active {active:5d} moves {c['moves']:7d} sims/s {ds / dt / 1e6:6.3f}M evals/s {de / dt / 1e6:6.3f}M "
          f"evals/sim {de / max(1, ds):5.3f} terminal/sim {(c['terminal_visits'] - prev['terminal_visits']) / max(1, ds):5.3f} ms/step {dt / STEP * 1e3:6.3f} sims/step {ds / STEP:7.1f}", flush=True)
    prev, t_prev = c, now
    if active == 0:
        break
print(f"total {c['sims'] / (now - t0) / 1e6:.3f} M sims/s, {now - t0:.2f} s")
eng.close()
