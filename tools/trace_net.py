#!/usr/bin/env python
"""Kernel timeline of plain network forwards (no tree): OZ_NET_TRACE=<slots> python tools/trace_net.py [boards]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from othellozero_b200 import engine as E, net as oznet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = E.Engine(8, max_games=B, nodes_per_game=2, prior_mode=E.PRIOR_NET, eval_cache_log2=0)
eng.load_weights(oznet.init_weights(8, 512, seed=0), 512)
po = E.perft_playouts(B, 8, seed=1, first_game_id=0, max_moves=20)
own = np.where(po["player"] == 1, po["white"], po["black"]); opp = np.where(po["player"] == 1, po["black"], po["white"])
for _ in range(int(os.environ.get("REPS", "4"))):
    eng.net_forward(own, opp, want_logits=False)
eng.close()
