"""Virtual-loss wave mode (vl_width > 1, BASELINE configs[3]).  It changes visit counts relative to the sequential
reference BY DESIGN, so the checks are: vl_width = 1 is the bit-exact path, and for vl_width > 1 the search is
deterministic, conserves visits, leaves no virtual loss behind, is independent of where the priors come from, and
plays legal complete games."""
import numpy as np
import pytest

import oracle
from gpu_util import canon_board, visits_to_grid

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    from othellozero_b200 import engine
    return engine


def test_vl_width_one_is_the_sequential_path(E, golden_roots):
    rec = golden_roots[1]
    e = E.Engine(rec["n"], 1, rec["sims"] + 8, E.PRIOR_HASH, vl_width=1)
    e.reset(1)
    e.search(rec["sims"])
    v, ns = e.visits()
    assert int(ns[0]) == rec["ns"] and visits_to_grid(v[0], rec["n"]).tolist() == rec["visits"]
    e.close()


@pytest.mark.parametrize("V", [2, 4, 8])
def test_wave_search_conserves_visits_and_is_deterministic(E, V):
    sims = 203
    outs = []
    for _ in range(2):
        e = E.Engine(8, 3, sims + 16, E.PRIOR_HASH, vl_width=V)
        e.reset(3)
        e.search(sims)
        v, ns = e.visits()
        c = e.counters()
        outs.append((v.copy(), ns.copy()))
        assert (ns == sims - 1).all()                 # every sim but the root expansion passes the root; vns == 0
        assert (v.sum(axis=1) == ns).all()
        assert c["sims"] == 3 * sims
        e.search(50)                                  # a second batch on the same tree keeps the invariants
        v2, ns2 = e.visits()
        assert (ns2 == sims + 49).all() and (v2.sum(axis=1) == ns2).all()
        e.close()
    assert np.array_equal(outs[0][0], outs[1][0])


def test_wave_search_is_independent_of_the_prior_source(E):
    """Host-fed priors (several leaves per game per step) give the same tree as the device-evaluated hash prior."""
    n, sims, V = 6, 90, 4
    dev = E.Engine(n, 1, sims + 16, E.PRIOR_HASH, vl_width=V)
    dev.reset(1); dev.search(sims)
    want, ns_want = dev.visits()
    dev.close()
    batch_sizes = []

    def predict_batch(own, opp):
        batch_sizes.append(len(own))
        outs = [oracle.hash_prior(canon_board(o, p, n)) for o, p in zip(own, opp)]
        return np.stack([o[0].ravel() for o in outs]), np.array([o[1] for o in outs], dtype=np.float32)

    host = E.Engine(n, 1, sims + 16, E.PRIOR_HOST, vl_width=V)
    host.reset(1); host.search(sims, predict_batch)
    got, ns_got = host.visits()
    host.close()
    assert np.array_equal(got, want) and np.array_equal(ns_got, ns_want)
    assert max(batch_sizes) > 1 and max(batch_sizes) <= V


def test_wave_selfplay_plays_legal_complete_games(E):
    from othellozero_b200 import engine as eng_mod
    n, sims, G, V = 8, 64, 256, 8
    e = E.Engine(n, G, sims * 61 + 64, E.PRIOR_HASH, seed=3, vl_width=V)
    e.selfplay_begin(G, sims, 1.0, 0.9)
    assert e.selfplay_run(-1) == 0
    rec = e.selfplay_records()
    c = e.counters()
    fb, fw, _ = e.positions()
    e.close()
    nm = rec["n_moves"]
    assert (rec["winner"] >= 0).all() and c["sims"] == sims * int(nm.sum())
    for p in range(int(nm.max())):
        idx = np.nonzero(nm > p)[0]
        b, w, pl = rec["black"][idx, p], rec["white"][idx, p], rec["player"][idx, p].astype(np.int64)
        own = np.where(pl == 0, b, w); opp = np.where(pl == 0, w, b)
        o2, p2, fl, _ = eng_mod.apply_moves(own, opp, rec["action"][idx, p].astype(np.int32), n)
        assert not (fl & 0x80000000).any()
        npl = np.where((fl & 1).astype(bool), 1 - pl, pl)
        nb = np.where(npl == 0, o2, p2); nw = np.where(npl == 0, p2, o2)
        last = nm[idx] == p + 1
        nxt = idx[~last]
        assert np.array_equal(nb[~last], rec["black"][nxt, p + 1]) and np.array_equal(nw[~last], rec["white"][nxt, p + 1])
        assert np.array_equal(nb[last], fb[idx[last]]) and np.array_equal(nw[last], fw[idx[last]])


def test_wave_selfplay_with_net_and_cache(E):
    """configs[3] shape in miniature: few games, several leaves in flight per game, device net, evaluation cache."""
    from othellozero_b200 import net
    n, C, sims, G, V = 6, 128, 40, 16, 4
    blob = net.init_weights(n, C, seed=21, randomize_bn=True)
    outs = []
    for lg in (0, 14):
        e = E.Engine(n, G, sims * 40, E.PRIOR_NET, seed=9, vl_width=V, log_visits=True, eval_cache_log2=lg)
        e.load_weights(blob, C)
        e.selfplay_begin(G, sims, 1.0, 0.85)
        assert e.selfplay_run(-1) == 0
        outs.append(e.selfplay_records())
        assert (outs[-1]["winner"] >= 0).all()
        e.close()
    for k in ("action", "n_moves", "winner", "visits"):
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_wave_mode_game_queue_matches_one_batch(E):
    """Queued self-play in wave mode: a game's moves and visit counts do not depend on the slot it ran in."""
    n, sims, V, total, slots = 6, 24, 4, 21, 5
    ids = np.arange(500, 500 + total, dtype=np.uint64)
    outs = []
    for mg in (slots, total):
        e = E.Engine(n, max_games=mg, nodes_per_game=sims * 36 + 64, prior_mode=E.PRIOR_HASH, vl_width=V,
                     log_visits=True)
        e.selfplay_begin(total, sims, 1.0, 0.85, game_ids=ids)
        assert e.selfplay_run(-1) == 0
        outs.append(e.selfplay_records())
        e.close()
    a, b = outs
    assert np.array_equal(a["n_moves"], b["n_moves"]) and np.array_equal(a["winner"], b["winner"])
    for g in range(total):
        k = int(a["n_moves"][g])
        assert np.array_equal(a["action"][g][:k], b["action"][g][:k])
        assert np.array_equal(a["visits"][g][:k], b["visits"][g][:k])
