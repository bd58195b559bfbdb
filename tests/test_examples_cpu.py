"""Example emission: the vectorised records -> examples path must reproduce the per-game path
(training.py:58-72 + the 8 symmetries of training.py:13-23) value for value, type for type, in the same order."""
import numpy as np
import pytest

from othellozero_b200 import selfplay
from othellozero_b200.net import bits_to_board


def _random_records(n, G, seed):
    rng = np.random.default_rng(seed)
    rec = dict(black=np.zeros((G, 64), np.uint64), white=np.zeros((G, 64), np.uint64),
               action=np.zeros((G, 64), np.uint8), player=np.zeros((G, 64), np.uint8),
               n_moves=np.zeros(G, np.int32), winner=np.zeros(G, np.int32))
    for g in range(G):
        k = int(rng.integers(0, n * n - 3)) if g else 0          # game 0 has no moves at all
        rec["n_moves"][g] = k
        rec["winner"][g] = rng.integers(-1, 2)                   # -1 = unfinished
        b = rng.integers(0, 2**63, k, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, k, dtype=np.uint64)
        rec["black"][g][:k] = b
        rec["white"][g][:k] = ~b & rng.integers(0, 2**63, k, dtype=np.uint64)
        rec["action"][g][:k] = (rng.integers(0, n, k) << 3) | rng.integers(0, n, k)
        rec["player"][g][:k] = rng.integers(0, 2, k)
    return rec


@pytest.mark.parametrize("n", [4, 6, 8])
def test_batched_examples_equal_per_game_examples(n):
    rec = _random_records(n, 40, seed=n)
    per_game = [selfplay.records_to_examples(rec, g, n) for g in range(40)]
    batched = selfplay.records_to_examples_batch(rec, n)
    assert [len(x) for x in batched] == [8 * int(k) for k in rec["n_moves"]]
    for ga, gb in zip(per_game, batched):
        assert len(ga) == len(gb)
        for (b1, p1, z1), (b2, p2, z2) in zip(ga, gb):
            assert b2.dtype == bool and b2.shape == (n, n, 2) and np.array_equal(b1, b2)
            assert p2.dtype == np.float64 and np.array_equal(p1, p2) and p2.sum() == 1
            assert z1 == z2 and type(z2) is int


def test_bits_to_boards_matches_scalar_conversion():
    rng = np.random.default_rng(0)
    b = rng.integers(0, 2**63, 50, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    w = ~b
    for n in (6, 8):
        got = selfplay.bits_to_boards(b, w, n)
        for i in range(50):
            assert np.array_equal(got[i], bits_to_board(int(b[i]), int(w[i]), n))


def test_batched_examples_subset_of_games():
    rec = _random_records(8, 10, seed=3)
    sub = selfplay.records_to_examples_batch(rec, 8, games=[7, 2])
    assert len(sub) == 2
    for got, g in zip(sub, (7, 2)):
        ref = selfplay.records_to_examples(rec, g, 8)
        assert len(got) == len(ref) and all(np.array_equal(a[0], b[0]) and a[2] == b[2] for a, b in zip(got, ref))
