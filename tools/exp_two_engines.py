#!/usr/bin/env python
"""Experiment: does running two half-size game groups on two streams hide the memory-bound / small kernels of one group
under the conv GEMMs of the other?  (a) one engine, G games; (b) two engines, G/2 games each, driven by two host threads."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from othellozero_b200 import engine as E, net as oznet

G, C, sims = int(os.environ.get("G", "4096")), 512, 100
NE = int(os.environ.get("NE", "2"))
STEPS = int(os.environ.get("STEPS", "300"))
cache = int(os.environ.get("CACHE", "0"))
blob = oznet.init_weights(8, C, seed=0)
engs = []
for i in range(NE):
    e = E.Engine(8, max_games=G // NE, nodes_per_game=sims * 61 + 64, prior_mode=E.PRIOR_NET, eval_cache_log2=cache, seed=i)
    e.load_weights(blob, C)
    ids = np.arange(i * (G // NE), (i + 1) * (G // NE), dtype=np.uint64)
    e.selfplay_begin(G // NE, sims, 1.0, 0.9, -1, None, None, None, ids)
    engs.append(e)

def run(e, steps):
    e.selfplay_run(steps)

def both(steps):
    th = [threading.Thread(target=run, args=(e, steps)) for e in engs]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()

both(100)
c0 = [e.counters() for e in engs]
t0 = time.perf_counter()
both(STEPS)
dt = time.perf_counter() - t0
c1 = [e.counters() for e in engs]
ds = sum(b["sims"] - a["sims"] for a, b in zip(c0, c1))
dn = sum(b["nodes"] - a["nodes"] - (b["cache_hits"] - a["cache_hits"]) - (b["cache_aliases"] - a["cache_aliases"]) for a, b in zip(c0, c1))
print(f"engines={NE} games={G} cache={cache} pdl={'off' if os.environ.get('OZ_NET_NO_PDL')=='1' else 'on'} steps={STEPS}: "
      f"{ds/dt/1e6:.3f} M sims/s, {dn/dt/1e6:.3f} M evals/s, {dt/STEPS*1e3:.3f} ms/step")
for e in engs: e.close()
