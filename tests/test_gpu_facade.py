"""The reference-named Python surface (OthelloGame / OthelloMCTS / execute_episode / Worker) over the C-ABI,
checked against golden vectors produced by the reference and against the oracle."""
import numpy as np
import pytest

import oracle
import prior_fns

pytestmark = pytest.mark.gpu


def test_othello_game_replays_golden_playouts(golden_playouts):
    from othellozero_b200.othello import BoardView, OthelloGame, OthelloPlayer
    for rec in golden_playouts[:6] + golden_playouts[-4:]:
        n = rec["n"]
        g = OthelloGame(n)
        for mv in rec["moves"]:
            acts = [tuple(int(x) for x in a) for a in g.get_valid_actions()]
            assert (mv // n, mv % n) in acts
            assert acts == oracle.valid_actions(g.board(BoardView.TWO_CHANNELS),
                                                0 if g.current_player is OthelloPlayer.BLACK else 1)
            g.play(mv // n, mv % n)
        assert g.has_finished()
        assert oracle.board_to_bits(g.board(BoardView.TWO_CHANNELS)) == (int(rec["b"], 16), int(rec["w"], 16))
        wp, pts = g.get_winning_player()
        assert (0 if wp is OthelloPlayer.BLACK else 1, int(pts)) == (rec["winner"], rec["points"])
        with pytest.raises(AssertionError):
            g.play(0, 0)  # 'Game has ended' (Othello/__init__.py:143)


def test_othello_static_api_and_aliases(golden_rules):
    from othellozero_b200.othello import OthelloGame, OthelloPlayer
    for rec in golden_rules["8"][5:40:7]:
        board = oracle.bits_to_board(int(rec["b"], 16), int(rec["w"], 16), 8).astype(bool)
        for ch, pl in ((0, OthelloPlayer.BLACK), (1, OthelloPlayer.WHITE)):
            exp = rec[f"moves{ch}"]
            acts = [tuple(int(x) for x in a) for a in OthelloGame.get_player_valid_actions(board, pl)]
            assert [r * 8 + c for r, c in acts] == [m[0] for m in exp]
            assert OthelloGame.getValidMoves(board, pl).sum() == len(exp)
            for (r, c), m in zip(acts, exp):
                nb = board.copy()
                OthelloGame.flip_board_squares(nb, pl, r, c)
                assert oracle.board_to_bits(nb) == (int(m[1], 16), int(m[2], 16))
        assert OthelloGame.has_board_finished(board) == rec["finished"]
        assert OthelloGame.getGameEnded(board) == (0 if not rec["finished"] else (1 if rec["winner"] == 0 else -1))
    with pytest.raises(TypeError):
        OthelloGame.get_player_valid_actions(board, 0)
    assert np.array_equal(OthelloGame.getCanonicalForm(board, OthelloPlayer.WHITE)[..., 0], board[..., 1])
    assert np.array_equal(OthelloGame.getInitBoard(6), oracle.initial_board(6).astype(bool))


def test_mcts_facade_hash_prior(golden_roots):
    from othellozero_b200.mcts import HashPriorNet, OthelloMCTS
    from othellozero_b200.othello import OthelloGame, OthelloPlayer
    rec = golden_roots[0]
    n = rec["n"]
    m = OthelloMCTS(n, HashPriorNet(), 1)
    st = OthelloGame.initial_board(n)
    for _ in range(rec["sims"]):
        m.simulate(st, OthelloPlayer.BLACK)
    assert m.N(st) == rec["ns"]
    for a in m.get_state_actions(st):
        assert m.N(st, a) == rec["visits"][a[0] * n + a[1]]
    pol = m.get_policy_action_probabilities(st, 1)
    exp = np.array(rec["visits"], dtype=float).reshape(n, n)
    assert np.array_equal(pol, exp / exp.sum())
    assert m.getActionProb(st, 0).sum() == 1
    m.close()


def test_mcts_facade_any_predict_object(golden_episodes):
    """A plain object with .predict (like the reference's NNetWrapper) is called once per expanded node."""
    from othellozero_b200.mcts import OthelloMCTS
    from othellozero_b200.net import NeuralNets
    from othellozero_b200.othello import OthelloGame, OthelloPlayer

    class Net:
        network_type = NeuralNets.ONN
        calls = 0

        def predict(self, board):
            Net.calls += 1
            return prior_fns.sha_prior(board)

    rec = golden_episodes["sha_6_25"]
    m = OthelloMCTS(6, Net(), 1)
    st = OthelloGame.initial_board(6)
    m.simulate(st, OthelloPlayer.BLACK, num_simulations=rec["sims"])
    got = [m.N(st, (r, c)) if (r, c) in m.get_state_actions(st) else 0 for r in range(6) for c in range(6)]
    assert got == rec["visits"][0]
    assert Net.calls == rec["sims"]  # one node (= one predict) per simulation from a fresh tree
    m.close()


def test_execute_episode_dropin(golden_episodes):
    from othellozero_b200.mcts import HashPriorNet
    from othellozero_b200.selfplay import execute_episode
    rec = golden_episodes["hash_6_25"]
    n = rec["n"]
    ex = execute_episode(n, HashPriorNet(), 1, rec["sims"], 1, 1.0)
    assert len(ex) == 8 * len(rec["moves"])
    moves = [int(np.argmax(ex[8 * i + 7][1])) for i in range(len(rec["moves"]))]  # identity symmetry is last
    assert moves == rec["moves"]
    for i, (board, pol, z) in enumerate(ex):
        assert board.shape == (n, n, 2) and pol.shape == (n, n) and pol.sum() == 1 and z in (1, -1)
    # z = +1 iff the mover is the winner
    for i, p in enumerate(rec["players"]):
        assert ex[8 * i][2] == (1 if p == rec["winner"] else -1)
    # first example of the first move: rot90 + fliplr of the initial board
    b0 = oracle.initial_board(n).astype(bool)
    assert np.array_equal(ex[0][0], np.fliplr(np.rot90(b0, k=1)))


def test_b200_worker_batches_episodes():
    from othellozero_b200.mcts import HashPriorNet
    from othellozero_b200.selfplay import B200Worker, WorkType
    w = B200Worker()
    w.run(WorkType.EXECUTE_EPISODE, 5, 6, HashPriorNet(), 1, 10, 1, 1.0)
    w.wait()
    res = w.get_results()
    assert len(res) == 5 and all(len(r) % 8 == 0 and len(r) > 0 for r in res)
    assert all(len(r) == len(res[0]) for r in res)  # e_greedy = 1.0 -> identical deterministic games


class _FirstChoice:
    @staticmethod
    def choice(seq):
        return seq[0]


def _oracle_duel(n, sims, c_black=1.0, c_white=1.0):
    """duel_between_agents (agents.py:71-84) with two NeuralNetworkOthelloAgents, restated with the oracle."""
    trees = [oracle.Mcts(n, c_black), oracle.Mcts(n, c_white)]
    board, player, moves = oracle.initial_board(n), 0, []
    while not oracle.has_finished(board):
        m = trees[player]
        for _ in range(sims):
            m.simulate(board, player)
        canon = board if player == 0 else board[..., ::-1]
        _, v = m.visits(np.ascontiguousarray(canon))
        a = int(np.argmax(v.ravel()))            # T = 0, first of the arg-max set
        moves.append(a)
        board = oracle.flip_board(board, player, a // n, a % n)
        nxt = 1 - player
        if not oracle.valid_actions(board, nxt):
            nxt = player
        player = nxt
    return moves, oracle.winner(board), oracle.board_to_bits(board)


def test_arena_pit_matches_oracle_duel():
    from othellozero_b200.arena import pit
    from othellozero_b200.mcts import HashPriorNet
    n, sims = 6, 20
    moves, (wch, pts), (fb, fw) = _oracle_duel(n, sims)
    out = pit(n, HashPriorNet(), HashPriorNet(), sims, 1, n_games=3, rng=_FirstChoice)
    assert out["winner"].tolist() == [wch] * 3
    assert [int(x) for x in out["black"]] == [fb] * 3 and [int(x) for x in out["white"]] == [fw] * 3
    assert out["plies"].tolist() == [len(moves)] * 3 and out["points"].tolist() == [pts] * 3


def test_agents_mirror_duel(monkeypatch):
    import random
    from othellozero_b200.arena import NeuralNetworkOthelloAgent, RandomOthelloAgent, duel_between_agents
    from othellozero_b200.mcts import HashPriorNet
    from othellozero_b200.othello import OthelloGame
    n, sims = 6, 20
    moves, (wch, pts), (fb, fw) = _oracle_duel(n, sims)
    monkeypatch.setattr(random, "choice", lambda seq: seq[0])
    game = OthelloGame(n)
    a1 = NeuralNetworkOthelloAgent(game, HashPriorNet(), sims, 1)
    a2 = NeuralNetworkOthelloAgent(game, HashPriorNet(), sims, 1)
    winner, points = duel_between_agents(game, a1, a2)
    assert (winner is a1) == (wch == 0) and int(points) == pts
    from othellozero_b200.othello import BoardView
    assert oracle.board_to_bits(game.board(BoardView.TWO_CHANNELS)) == (fb, fw)
    g2 = OthelloGame(n)
    w2, _ = duel_between_agents(g2, RandomOthelloAgent(g2), RandomOthelloAgent(g2))
    assert g2.has_finished()


class _FirstChoiceRng(_FirstChoice):
    """rng for pit(): first element of every choice; colour coin flips alternate."""
    def __init__(self):
        self.k = 0

    def random(self):
        self.k += 1
        return 0.25 if self.k % 2 else 0.75


def _oracle_duel_vs_first_move(n, sims, net_is_black):
    """A NeuralNetworkOthelloAgent (T=0, first arg-max) against an agent that always plays its first legal move in
    row-major order (RandomOthelloAgent with random.choice -> seq[0]), restated with the oracle."""
    tree = oracle.Mcts(n, 1.0)
    board, player, plies = oracle.initial_board(n), 0, 0
    while not oracle.has_finished(board):
        if (player == 0) == net_is_black:
            for _ in range(sims):
                tree.simulate(board, player)
            canon = board if player == 0 else board[..., ::-1]
            _, v = tree.visits(np.ascontiguousarray(canon))
            a = int(np.argmax(v.ravel()))
        else:
            r, c = oracle.valid_actions(board, player)[0]
            a = int(r) * n + int(c)
        board = oracle.flip_board(board, player, a // n, a % n)
        plies += 1
        nxt = 1 - player
        if not oracle.valid_actions(board, nxt):
            nxt = player
        player = nxt
    return oracle.winner(board), oracle.board_to_bits(board), plies


@pytest.mark.parametrize("net_is_black", [True, False])
def test_pit_network_vs_random_agent_matches_oracle(net_is_black):
    from othellozero_b200.arena import RANDOM_AGENT, pit
    from othellozero_b200.mcts import HashPriorNet
    n, sims = 6, 16
    (wch, pts), (fb, fw), plies = _oracle_duel_vs_first_move(n, sims, net_is_black)
    sides = (HashPriorNet(), RANDOM_AGENT) if net_is_black else (RANDOM_AGENT, HashPriorNet())
    out = pit(n, sides[0], sides[1], sims, 1, n_games=2, rng=_FirstChoice)
    assert out["winner"].tolist() == [wch] * 2 and out["plies"].tolist() == [plies] * 2
    assert [int(x) for x in out["black"]] == [fb] * 2 and [int(x) for x in out["white"]] == [fw] * 2


def test_batched_arena_drivers_and_worker_work_types():
    """workers.py:18-21,72-79: the three work types through the worker seam, one entry per iteration."""
    import random
    from othellozero_b200 import arena
    from othellozero_b200.mcts import HashPriorNet
    from othellozero_b200.selfplay import B200Worker, WorkType
    n, sims = 6, 12
    net = HashPriorNet()
    # evaluate: alternating colours, every random choice = first element -> two distinct deterministic games
    wins = arena.evaluate_neural_network(n, 6, net, sims, 1, rng=_FirstChoiceRng())
    (w_b, _), _, _ = _oracle_duel_vs_first_move(n, sims, True)
    (w_w, _), _, _ = _oracle_duel_vs_first_move(n, sims, False)
    assert wins == 3 * int(w_b == 0) + 3 * int(w_w == 1)
    assert arena.evaluate_neural_network(n, 4, net, sims, 1, rng=random.Random(3), repeats=3).__len__() == 3
    with pytest.raises(TypeError):
        arena.evaluate_neural_network(n, 2, net, sims, 1, agent_class=arena.NeuralNetworkOthelloAgent)
    w = B200Worker()
    w.run(WorkType.DUEL_BETWEEN_NEURAL_NETWORKS, 4, n, net, net, 1, sims)
    w.wait()
    duels = w.get_results()
    assert len(duels) == 4 and set(duels) <= {0, 1}
    w.run(WorkType.EVALUATE_NEURAL_NETWORK, 2, n, 5, net, sims, 1, arena.RandomOthelloAgent, ())
    w.wait()
    ev = w.get_results()
    assert len(ev) == 2 and all(0 <= x <= 5 for x in ev)
    with pytest.raises(TypeError):
        w._run("no such work", 1, (), {})


# ---- example stream at the drop-in boundary, against the reference's own stream ----------------------------------------
def test_execute_episode_stream_equals_reference_digest(golden_examples):
    """selfplay.execute_episodes (device self-play -> records -> examples) must reproduce training.execute_episode's list
    byte for byte: the digests were taken from the reference's own output (tools/gen_golden.py examples), with the
    engine's draws injected for T = 0 / e_greedy < 1."""
    from example_digest import examples_digest
    from othellozero_b200.mcts import HashPriorNet
    from othellozero_b200.selfplay import execute_episodes
    done = 0
    for g in golden_examples:
        if g["prior"] != "hash":
            continue  # host-evaluated priors: covered through the oracle in tests/test_examples_reference_cpu.py
        kw = dict(seed=g["seed"], game_ids=[g["game_id"]])
        aliased = execute_episodes(1, g["n"], HashPriorNet(), 1, g["sims"], g["T"], g["e_greedy"], reference_aliasing=True, **kw)[0]
        snap = execute_episodes(1, g["n"], HashPriorNet(), 1, g["sims"], g["T"], g["e_greedy"], **kw)[0]
        assert len(aliased) == len(snap) == g["n_examples"]
        assert examples_digest(aliased) == g["sha256_reference_stream"]
        assert examples_digest(snap) == g["sha256_snapshot_stream"]
        done += 1
    assert done >= 4


def test_workers_and_runs_never_replay_episodes():
    """every worker / every WorkerManager.run() used to replay the same e-greedy streams (fixed seed 0, ids
    0..n-1).  Now ids come from a process-wide allocator and the default seed is drawn per process."""
    from othellozero_b200.mcts import HashPriorNet
    from othellozero_b200.selfplay import WorkType, make_b200_worker

    def moves_of(results):
        return {tuple(int(np.argmax(ex[8 * i + 7][1])) for i in range(len(ex) // 8)) for ex in results}

    w1, w2 = make_b200_worker()(), make_b200_worker()()
    runs = []
    for w in (w1, w2, w1):
        w.run(WorkType.EXECUTE_EPISODE, 6, 6, HashPriorNet(), 1, 8, 1, 0.4)
        w.wait()
        res = w.get_results()
        assert len(res) == 6
        runs.append(moves_of(res))
    assert len(runs[0]) >= 5                                   # e_greedy 0.4: episodes of one run differ
    assert not (runs[0] & runs[1]) and not (runs[0] & runs[2]) and not (runs[1] & runs[2])
    # an explicit seed + explicit ids stay reproducible
    from othellozero_b200.selfplay import execute_episodes
    a = execute_episodes(3, 6, HashPriorNet(), 1, 8, 1, 0.4, seed=5, game_ids=[1, 2, 3])
    b = execute_episodes(3, 6, HashPriorNet(), 1, 8, 1, 0.4, seed=5, game_ids=[1, 2, 3])
    assert moves_of(a) == moves_of(b) and len(moves_of(a)) == 3
