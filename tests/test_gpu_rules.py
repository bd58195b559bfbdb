"""K1-K3 (CUDA bitboard rules, through the C-ABI) against the oracle and the golden vectors. Bit-exact."""
import numpy as np
import pytest

import oracle
from gpu_util import sq8

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from othellozero_b200 import engine
    return engine


@pytest.mark.parametrize("n", [4, 6, 8])
def test_legal_and_apply_golden(eng, golden_rules, n):
    recs = golden_rules[str(n)]
    own, opp, exp_legal = [], [], []
    a_own, a_opp, a_sq, a_exp = [], [], [], []
    for rec in recs:
        b, w = int(rec["b"], 16), int(rec["w"], 16)
        for ch, (o, p) in ((0, (b, w)), (1, (w, b))):
            own.append(o); opp.append(p)
            m = 0
            for sq, fb, fw in rec[f"moves{ch}"]:
                m |= 1 << sq
                a_own.append(o); a_opp.append(p); a_sq.append(sq)
                fb, fw = int(fb, 16), int(fw, 16)
                a_exp.append((fb, fw) if ch == 0 else (fw, fb))  # (mover, other) after the flips
            exp_legal.append(m)
    legal = eng.legal_moves(np.array(own, dtype=np.uint64), np.array(opp, dtype=np.uint64), n)
    assert [int(x) for x in legal] == exp_legal
    o2, p2, fl, nl = eng.apply_moves(np.array(a_own, dtype=np.uint64), np.array(a_opp, dtype=np.uint64),
                                     np.array(a_sq, dtype=np.int32), n)
    for i, (mo, ot) in enumerate(a_exp):
        got = (int(p2[i]), int(o2[i])) if fl[i] & 1 else (int(o2[i]), int(p2[i]))
        assert got == (mo, ot)
        # flags / next_legal agree with the oracle's turn logic
        board = oracle.bits_to_board(mo, ot, n)
        opp_can, own_can = bool(oracle.valid_actions(board, 1)), bool(oracle.valid_actions(board, 0))
        want = 1 if opp_can else (2 if own_can else 4)
        assert int(fl[i]) == want


def test_apply_rejects_illegal(eng):
    b, w = oracle.board_to_bits(oracle.initial_board(8))
    o2, p2, fl, nl = eng.apply_moves([b], [w], [0], 8)
    assert fl[0] == 0x80000000 and (int(o2[0]), int(p2[0])) == (b, w)


def test_empty_batch(eng):
    assert eng.legal_moves(np.zeros(0, np.uint64), np.zeros(0, np.uint64), 8).size == 0


def test_playouts_golden(eng, golden_playouts):
    for rec in golden_playouts:
        n = rec["n"]
        out = eng.perft_playouts(1, n, seed=rec["seed"], first_game_id=rec["game_id"], want_moves=True)
        k = int(out["plies"][0])
        assert [int(m) for m in out["moves"][0][:k]] == [sq8(a, n) for a in rec["moves"]]
        assert (int(out["black"][0]), int(out["white"][0])) == (int(rec["b"], 16), int(rec["w"], 16))
        assert out["finished"][0]


@pytest.mark.parametrize("n,count", [(8, 4096), (6, 2048)])
def test_playouts_vs_oracle(eng, n, count):
    """>= 4096 sampled games replayed by the oracle: move lists, final boards, plies, winner (SURVEY §8d cfg 2)."""
    out = eng.perft_playouts(count, n, seed=99, first_game_id=500000, want_moves=True)
    for g in range(count):
        ref = oracle.playout(n, 99, 500000 + g)
        k = int(out["plies"][g])
        assert k == len(ref["moves"])
        assert [int(m) for m in out["moves"][g][:k]] == [sq8(a, n) for a in ref["moves"]]
        assert (int(out["black"][g]), int(out["white"][g])) == (ref["black"], ref["white"])
        assert int(out["player"][g]) == ref["player"]


def test_playouts_full_size_properties(eng):
    """Config 2 at full size (1M games): size-independent invariants + determinism + a sampled oracle check."""
    n_games = 1 << 20
    a = eng.perft_playouts(n_games, 8, seed=0)
    b = eng.perft_playouts(n_games, 8, seed=0)
    assert np.array_equal(a["black"], b["black"]) and np.array_equal(a["white"], b["white"])
    assert a["finished"].all()
    assert not (a["black"] & a["white"]).any()
    discs = np.array([bin(int(x)).count("1") for x in (a["black"][:4096] | a["white"][:4096])])
    assert (discs == a["plies"][:4096] + 4).all()          # one disc placed per ply, none ever removed
    assert np.array_equal(eng.legal_moves(a["black"], a["white"], 8), np.zeros(n_games, np.uint64))
    assert np.array_equal(eng.legal_moves(a["white"], a["black"], 8), np.zeros(n_games, np.uint64))
    for g in (0, 1, 777777, n_games - 1):
        ref = oracle.playout(8, 0, g)
        assert (int(a["black"][g]), int(a["white"][g])) == (ref["black"], ref["white"])
    # shard independence: games are keyed by global id
    part = eng.perft_playouts(1000, 8, seed=0, first_game_id=500000)
    assert np.array_equal(part["black"], a["black"][500000:501000])


def test_host_entry_points_from_many_threads_and_sizes():
    """The host-buffer rules calls keep one staging buffer per calling thread (grown on demand): concurrent callers with
    different batch sizes must not see each other's data."""
    import threading
    from othellozero_b200 import engine as E
    n = 8
    base = E.perft_playouts(3000, n, seed=5, max_moves=11)
    own = np.where(base["player"] == 0, base["black"], base["white"])
    opp = np.where(base["player"] == 0, base["white"], base["black"])
    ref_legal = E.legal_moves(own, opp, n)
    ref_score = E.score(base["black"], base["white"])
    errors = []

    def worker(tid):
        try:
            rng = np.random.default_rng(tid)
            for it in range(30):
                k = int(rng.integers(1, 3000))            # sizes shrink and grow -> the scratch is reused and regrown
                lo = int(rng.integers(0, 3000 - k + 1))
                sl = slice(lo, lo + k)
                assert np.array_equal(E.legal_moves(own[sl], opp[sl], n), ref_legal[sl])
                cb, cw = E.score(base["black"][sl], base["white"][sl])
                assert np.array_equal(cb, ref_score[0][sl]) and np.array_equal(cw, ref_score[1][sl])
                out = E.perft_playouts(k, n, seed=5, first_game_id=lo, max_moves=11)
                assert np.array_equal(out["black"], base["black"][sl]) and np.array_equal(out["white"], base["white"][sl])
        except Exception as ex:  # noqa: BLE001
            errors.append((tid, repr(ex)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
