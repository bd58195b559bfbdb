#!/usr/bin/env python
"""Kernel timeline of a few engine steps (OZ_NET_TRACE=<slots> makes every tower kernel record its %globaltimer span;
the engine prints the table to stderr when it is destroyed).  Usage: OZ_NET_TRACE=64 OZ_NET_CHUNKS=2 python tools/trace_forward.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from othellozero_b200 import engine as E, net as oznet

G, C, sims = 4096, 512, 100
eng = E.Engine(8, max_games=G, nodes_per_game=sims * 61 + 64, prior_mode=E.PRIOR_NET, eval_cache_log2=0)
eng.load_weights(oznet.init_weights(8, C, seed=0), C)
eng.selfplay_begin(G, sims, 1.0, 0.9)
eng.selfplay_run(int(os.environ.get("STEPS", "6")))
eng.close()
