"""The oracle (oracle/oz_oracle.c) against golden vectors produced by executing the reference
(tools/gen_golden.py) and the known-answer vectors of SURVEY Appendix B."""
import numpy as np
import pytest

import oracle


def _bits(board):
    return oracle.board_to_bits(board)


def test_perft_known_answers():
    assert [oracle.perft(8, d) for d in range(1, 7)] == [4, 12, 56, 244, 1396, 8200]
    assert [oracle.perft(6, d) for d in range(1, 8)] == [4, 12, 56, 244, 1364, 7604, 47712]


def test_quirk_and_draw_witness():
    # row 0 = [., W, B, W, B] ; BLACK plays (0,0) -> flips (0,1) AND (0,3)
    b = np.zeros((8, 8, 2), dtype=np.uint8)
    b[0, 1, 1] = b[0, 3, 1] = 1
    b[0, 2, 0] = b[0, 4, 0] = 1
    f = oracle.flip_squares(b, 0, 0, 0)
    assert sorted(zip(*np.nonzero(f))) == [(0, 1), (0, 3)]
    d = np.zeros((4, 4, 2), dtype=np.uint8)
    d[:2, :, 0] = 1
    d[2:, :, 1] = 1
    assert oracle.winner(d) == (0, 8)


@pytest.mark.parametrize("n", [4, 6, 8])
def test_rules_golden(golden_rules, n):
    for rec in golden_rules[str(n)]:
        board = oracle.bits_to_board(int(rec["b"], 16), int(rec["w"], 16), n)
        for ch in (0, 1):
            exp = rec[f"moves{ch}"]
            acts = oracle.valid_actions(board, ch)
            assert [r * 8 + c for r, c in acts] == [m[0] for m in exp]
            for (r, c), m in zip(acts, exp):
                nb = oracle.flip_board(board, ch, r, c)
                assert _bits(nb) == (int(m[1], 16), int(m[2], 16))
        assert oracle.has_finished(board) == rec["finished"]
        assert oracle.winner(board) == (rec["winner"], rec["points"])


def test_playouts_golden(golden_playouts):
    for rec in golden_playouts:
        out = oracle.playout(rec["n"], rec["seed"], rec["game_id"])
        assert out["moves"] == rec["moves"]
        assert (out["black"], out["white"]) == (int(rec["b"], 16), int(rec["w"], 16))
        assert out["finished"]
        assert oracle.winner(out["board"]) == (rec["winner"], rec["points"])


def test_root_visits_golden(golden_roots):
    for rec in golden_roots:
        m = oracle.Mcts(rec["n"], 1.0)
        b = oracle.initial_board(rec["n"])
        for _ in range(rec["sims"]):
            m.simulate(b, 0)
        ns, v = m.visits(b)
        assert ns == rec["ns"]
        assert v.ravel().tolist() == rec["visits"]
        assert m.net_calls == rec["net_calls"]


def _prior(name):
    import prior_fns
    kind = name.split("_")[0]
    return {"hash": None, "sha": prior_fns.sha_prior, "zero": prior_fns.zero_prior}[kind]


@pytest.mark.parametrize("name", ["hash_6_25", "hash_4_40", "sha_4_30", "sha_6_25", "zero_6_10", "hash_6_60_c2",
                                  "hash_8_100", "sha_8_50"])
def test_episode_golden(golden_episodes, name):
    rec = golden_episodes[name]
    out = oracle.execute_episode(rec["n"], rec["sims"], c=rec["c"], temperature=rec["T"], e_greedy=1.0,
                                 predict=_prior(name), log_visits=True)
    assert out["moves"] == rec["moves"]
    assert out["players"] == rec["players"]
    assert out["visits"].tolist() == rec["visits"]
    assert out["winner"] == rec["winner"]
    assert out["net_calls"] == rec["net_calls"]


def test_q_values_and_types_golden(golden_episodes):
    """Q after the first move's simulations: value AND dynamic type (python float vs np.float32)."""
    import prior_fns
    for name in ("hash_4_40", "sha_4_30", "hash_6_25"):
        rec = golden_episodes[name]
        n = rec["n"]
        m = oracle.Mcts(n, rec["c"], _prior(name))
        b = oracle.initial_board(n)
        for _ in range(rec["sims"]):
            m.simulate(b, 0)
        q, p, tag = m.node_stats(b)
        for k, (val, typ) in rec["q"][0].items():
            sq = int(k)
            assert q.ravel()[sq] == val
            want = {"int": 0, "float": 1, "float32": 2}[typ]
            assert tag.ravel()[sq] == want


def test_np_sum_matches_numpy():
    rng = np.random.default_rng(0)
    for n in (16, 36, 64):
        for _ in range(200):
            a = rng.random(n) * (rng.random(n) < 0.3)
            side = int(round(n ** 0.5))
            assert oracle.np_sum(a) == float(np.sum(a.reshape(side, side)))


def test_rng_episodes_golden(golden_rng_episodes):
    """The reference's OWN execute_episode (training.py:26-72), run with the engine's counter RNG injected into its three
    draw sites (tools/gen_golden.py::EngineRng): temperature 0 with random.choice over tied arg-max actions
    (othelo_mcts.py:54-62), the e-greedy coin and the random legal action (training.py:51-56)."""
    import prior_fns
    ties = 0
    for rec in golden_rng_episodes:
        predict = None if rec["prior"] == "hash" else prior_fns.sha_prior
        out = oracle.execute_episode(rec["n"], rec["sims"], c=rec["c"], temperature=rec["T"], e_greedy=rec["e_greedy"],
                                     predict=predict, seed=rec["seed"], game_id=rec["game_id"])
        assert out["moves"] == rec["moves"], rec
        assert [1 if out["winner"] == p else -1 for p in out["players"]] == rec["z"]
        assert out["net_calls"] == rec["net_calls"]
        ties += rec["rng_calls"]["ties"]
    assert ties > 50  # the tie-break draw really decided moves
