"""Visit-count parity at the BENCHMARK shape (BASELINE.json configs[2]): 8x8, C = 512 network priors (OZ_PRIOR_NET), 100
simulations per move, evaluation cache on, >= 16 distinct start positions, more games than slots (device-side queue) - the very
path bench.py times.  The oracle search is fed the DEVICE network's priors (batch-1 forwards of the same weights; the
tower is row-independent, so they are bit-identical to the batched ones), so moves and per-move visit counts must be
bit-exact (north_star: "MCTS visit counts must be bit-exact when both sides are fed identical network priors")."""
import numpy as np
import pytest

import oracle
from gpu_util import sq8, visits_to_grid

pytestmark = pytest.mark.gpu


def _device_predict(engine_mod, e, n):
    from othellozero_b200 import net as oznet
    cache = {}

    def predict(board):
        o, p = oznet.boards_to_bits(board)
        key = (int(o[0]), int(p[0]))
        if key not in cache:
            pi, _, v = e.net_forward(o, p, want_logits=False)
            cache[key] = (pi[0].reshape(n, n).copy(), np.float32(v[0]))
        return cache[key]
    return predict


@pytest.mark.parametrize("temperature,e_greedy,max_moves,games,slots", [(1.0, 0.9, 6, 22, 8), (0.0, 1.0, 5, 20, 6)])
def test_config2_shape_visit_counts_match_oracle(temperature, e_greedy, max_moves, games, slots):
    from othellozero_b200 import engine as E, net as oznet
    n, C, sims = 8, 512, 100
    blob = oznet.init_weights(n, C, seed=0)                      # the bench's weights (Keras default init)
    st = E.perft_playouts(games, n, seed=31, max_moves=8)         # >= 16 distinct mid-opening starts
    black, white, player = st["black"], st["white"], st["player"].astype(np.int32)
    # the last four games start where the first four do (other game ids, so other e-greedy draws): identical positions in
    # different games are what the cross-game evaluation cache exists for
    black[-4:], white[-4:], player[-4:] = black[:4], white[:4], player[:4]
    assert len({(int(b), int(w)) for b, w in zip(black, white)}) >= games - 4 >= 12
    ids = np.arange(700, 700 + games, dtype=np.uint64)
    e = E.Engine(n, max_games=slots, nodes_per_game=sims * 61 + 64, prior_mode=E.PRIOR_NET, seed=5, log_visits=True,
                 eval_cache_log2=16)
    e.load_weights(blob, C)
    e.selfplay_begin(games, sims, temperature, e_greedy, max_moves, black, white, player, ids)
    assert e.selfplay_run(-1) == 0
    rec = e.selfplay_records()
    c = e.counters()
    assert c["cache_hits"] + c["cache_aliases"] > 0               # the cache really served evaluations
    # a second, network-only engine provides the priors the oracle search consumes
    f = E.Engine(n, max_games=4, nodes_per_game=2, prior_mode=E.PRIOR_NET)
    f.load_weights(blob, C)
    predict = _device_predict(E, f, n)
    total_nodes = 0
    for g in range(games):
        ref = oracle.execute_episode(n, sims, c=1.0, temperature=temperature, e_greedy=e_greedy, predict=predict, seed=5,
                                     game_id=int(ids[g]), start_board=oracle.bits_to_board(int(black[g]), int(white[g]), n),
                                     start_player=int(player[g]), max_moves=max_moves, log_visits=True)
        k = int(rec["n_moves"][g])
        assert k == len(ref["moves"]) == max_moves
        assert [int(a) for a in rec["action"][g][:k]] == [sq8(a, n) for a in ref["moves"]], f"game {g}"
        for p in range(k):
            assert visits_to_grid(rec["visits"][g][p], n).tolist() == ref["visits"][p].tolist(), f"game {g} move {p}"
        total_nodes += ref["net_calls"]
    assert c["nodes"] == total_nodes                              # one expansion per oracle net call (othelo_mcts.py:82-88)
    e.close(); f.close()


def test_config2_full_game_visit_counts_match_oracle():
    """One complete 8x8 / 100-sim / C=512 episode: every move's visit counts, the winner and the final position."""
    from othellozero_b200 import engine as E, net as oznet
    n, C, sims = 8, 512, 100
    blob = oznet.init_weights(n, C, seed=0)
    e = E.Engine(n, max_games=2, nodes_per_game=sims * 61 + 64, prior_mode=E.PRIOR_NET, seed=1, log_visits=True,
                 eval_cache_log2=14)
    e.load_weights(blob, C)
    e.selfplay_begin(2, sims, 1.0, 1.0, game_ids=[0, 1])          # e_greedy 1: both games are the same deterministic game
    assert e.selfplay_run(-1) == 0
    rec = e.selfplay_records()
    predict = _device_predict(E, e, n)
    ref = oracle.execute_episode(n, sims, predict=predict, log_visits=True)
    e.close()
    for g in range(2):
        k = int(rec["n_moves"][g])
        assert [int(a) for a in rec["action"][g][:k]] == [sq8(a, n) for a in ref["moves"]]
        assert int(rec["winner"][g]) == ref["winner"]
        for p in range(k):
            assert visits_to_grid(rec["visits"][g][p], n).tolist() == ref["visits"][p].tolist(), f"move {p}"
