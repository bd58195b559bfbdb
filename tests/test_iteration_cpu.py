"""configs[4] host logic that needs no GPU: packed rows -> training arrays (8 symmetries), buffer accounting."""
import numpy as np

import oracle
from example_digest import oracle_episode_as_records
from othellozero_b200 import dist as ozd, iteration, selfplay


def test_training_arrays_equal_the_example_stream():
    n = 6
    out = oracle.execute_episode(n, 12, e_greedy=0.8, seed=4, game_id=9)
    rec = oracle_episode_as_records(out, n)
    rows = ozd.pack_records(rec)
    assert rows.shape == (len(out["moves"]), 3)
    boards, pols, z = iteration.training_arrays(rows, n)
    ex = selfplay.records_to_examples(rec, 0, n)
    assert boards.shape == (len(ex), n, n, 2) and pols.shape == (len(ex), n * n) and z.shape == (len(ex),)
    for i, (b, p, zz) in enumerate(ex):
        assert np.array_equal(boards[i].astype(bool), b) and np.array_equal(pols[i].reshape(n, n), p) and z[i] == zz


def test_unpack_rows_round_trip():
    rec = dict(n_moves=np.array([2, 1]), winner=np.array([1, 0]),
               action=np.array([[9, 18] + [255] * 62, [27] + [255] * 63], dtype=np.uint8),
               player=np.array([[0, 1] + [0] * 62, [0] * 64], dtype=np.uint8),
               black=np.arange(128, dtype=np.uint64).reshape(2, 64), white=np.arange(128, 256, dtype=np.uint64).reshape(2, 64))
    b, w, a, z = iteration.unpack_rows(ozd.pack_records(rec))
    assert b.tolist() == [0, 1, 64] and w.tolist() == [128, 129, 192] and a.tolist() == [9, 18, 27]
    assert z.tolist() == [-1, 1, 1]        # winner WHITE: the BLACK mover lost, the WHITE mover won; game 2: BLACK won
