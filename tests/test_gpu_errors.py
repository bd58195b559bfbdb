"""Error behaviour and handle isolation of the C-ABI on a real device (SURVEY §8b 'Errors' / 'Threading')."""
import threading

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def test_invalid_arguments_raise_like_the_reference():
    from othellozero_b200 import _lib, engine as E
    with pytest.raises(AssertionError):          # Othello/__init__.py:36 'Board size must be even' style asserts
        E.Engine(7, 1, 64)
    with pytest.raises(AssertionError):
        E.legal_moves([1], [2], board_size=5)
    with pytest.raises(AssertionError):
        E.Engine(8, 0, 64)
    e = E.Engine(8, 2, 64, E.PRIOR_HASH)
    with pytest.raises(AssertionError):
        e.reset(3)                                # more games than slots
    with pytest.raises(_lib.OzError):
        e.search(10)                              # search before reset
    e.reset(2)
    with pytest.raises(AssertionError):
        e.selfplay_begin(2, 1)                    # num_sims = 1: the reference's policy would be all-zero
    with pytest.raises(AssertionError):
        e.selfplay_begin(2, 10, temperature=-1.0)
    with pytest.raises(AssertionError):           # a start whose side to move cannot move (here: an empty board)
        e.selfplay_begin(2, 10, black=[0x0000001008000000, 0], white=[0x0000000810000000, 0], player=[0, 0])
    with pytest.raises(AssertionError):           # overlapping discs
        e.selfplay_begin(1, 10, black=[0x0000001818000000], white=[0x0000000810000000], player=[0])
    with pytest.raises(_lib.OzError):
        e.net_forward([1], [2])                   # no network in this engine
    with pytest.raises(AssertionError):
        E.check(e._L.oz_selfplay_get_records(e._h, None, None, None, None, None, None, None))  # null buffers
    e.close()
    n = E.Engine(8, 4, 2, E.PRIOR_NET)
    with pytest.raises(_lib.OzError):
        n.net_forward([1], [2])                   # weights not loaded
    with pytest.raises(AssertionError):
        n.load_weights(np.zeros(10, dtype=np.float32), 512)   # wrong blob size
    with pytest.raises(AssertionError):
        n.load_weights(np.zeros(10, dtype=np.float32), 100)   # channels not a multiple of 128
    n.close()
    h = E.Engine(6, 1, 256, E.PRIOR_HOST)
    h.reset(1)
    with pytest.raises(_lib.OzError):
        h.search(5)                               # host-prior engine needs a predict callback
    h.close()


def test_engines_are_independent_across_threads():
    """One handle per thread (the reference starts one thread per worker, workers.py:33-37)."""
    from othellozero_b200 import engine as E
    ref = oracle.execute_episode(6, 12)
    results = [None] * 4

    def work(i):
        e = E.Engine(6, 3, 12 * 40, E.PRIOR_HASH)
        e.selfplay_begin(3, 12, 1.0, 1.0)
        e.selfplay_run(-1)
        rec = e.selfplay_records()
        k = int(rec["n_moves"][0])
        results[i] = [int(a) for a in rec["action"][0][:k]]
        e.close()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    want = [(a // 6) * 8 + a % 6 for a in ref["moves"]]
    assert all(r == want for r in results)


def test_engine_reuse_after_reset_is_clean():
    from othellozero_b200 import engine as E
    e = E.Engine(6, 2, 25 * 40, E.PRIOR_HASH, log_visits=True)
    outs = []
    for _ in range(2):
        e.selfplay_begin(2, 25, 1.0, 1.0)
        assert e.selfplay_run(-1) == 0
        rec = e.selfplay_records()
        outs.append((rec["action"].copy(), rec["visits"].copy(), e.counters()["nodes"]))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2]
    e.close()
