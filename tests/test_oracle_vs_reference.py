"""Oracle against the LIVE reference (imported from /root/reference).  Skipped where the
reference does not exist (the GPU box); the committed golden fixtures cover that case."""
import random

import numpy as np
import pytest

import oracle
import prior_fns
from tools import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference not present")


@pytest.fixture(scope="module")
def R():
    return ref_loader.load()


@pytest.mark.parametrize("n", [6, 8])
def test_rules_random_game(R, n):
    Game, Player = R.Othello.OthelloGame, R.Othello.OthelloPlayer
    rng = random.Random(n)
    g = Game(n)
    while not g.has_finished():
        board = g.board(R.Othello.BoardView.TWO_CHANNELS)
        for ch, pl in ((0, Player.BLACK), (1, Player.WHITE)):
            ref_acts = [tuple(int(x) for x in a) for a in Game.get_player_valid_actions(board, pl)]
            assert oracle.valid_actions(board, ch) == ref_acts
            for r, c in ref_acts:
                nb = np.copy(board)
                Game.flip_board_squares(nb, pl, r, c)
                assert np.array_equal(oracle.flip_board(board, ch, r, c), nb.astype(np.uint8))
        acts = list(g.get_valid_actions())
        a = acts[rng.randrange(len(acts))]
        g.play(int(a[0]), int(a[1]))


def test_search_visits_nondyadic_prior(R):
    """Non-dyadic float32 priors: exercises numpy's pairwise np.sum order and f32/f64 Q updates."""
    n, sims = 6, 60
    net = ref_loader.StubNet(prior_fns.sha_prior)
    mcts = R.othelo_mcts.OthelloMCTS(n, net, 1)
    g = R.Othello.OthelloGame(n)
    st = g.board(R.Othello.BoardView.TWO_CHANNELS)
    m = oracle.Mcts(n, 1.0, prior_fns.sha_prior)
    for _ in range(sims):
        mcts.simulate(st, R.Othello.OthelloPlayer.BLACK)
        m.simulate(st, 0)
    ns, v = m.visits(st)
    assert ns == mcts.N(st)
    h = R.MCTS.hash_ndarray(st)
    q, p, tag = m.node_stats(st)
    for a in mcts.get_state_actions(st):
        assert v[a] == mcts.N(st, a)
        assert q[a] == float(mcts._Qsa[h][a])
        assert p[a] == float(mcts._Psa[h][a])
    assert m.net_calls == net.calls
