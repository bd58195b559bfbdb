"""K4-K6/K11 (CUDA PUCT search + self-play driver) against the oracle and the reference's golden episodes.
Visit counts, Q values (incl. their float32/float64 typing), moves and winners must be bit-exact."""
import numpy as np
import pytest

import oracle
import prior_fns
from gpu_util import canon_board, sq8, visits_to_grid

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    from othellozero_b200 import engine
    return engine


def test_root_visits_golden(E, golden_roots):
    for rec in golden_roots:
        n = rec["n"]
        e = E.Engine(n, max_games=1, nodes_per_game=rec["sims"] + 8, prior_mode=E.PRIOR_HASH)
        e.reset(1)
        e.search(rec["sims"])
        v, ns = e.visits()
        assert int(ns[0]) == rec["ns"]
        assert visits_to_grid(v[0], n).tolist() == rec["visits"]
        assert e.counters()["nodes"] == rec["net_calls"]
        e.close()


def test_sims_one_at_a_time_equals_batch(E):
    """OthelloMCTS.simulate is called once per simulation by the reference (training.py:42-43)."""
    a = E.Engine(8, 1, 256, E.PRIOR_HASH)
    b = E.Engine(8, 1, 256, E.PRIOR_HASH)
    a.reset(1); b.reset(1)
    a.search(60)
    for _ in range(60):
        b.search(1)
    assert np.array_equal(a.visits()[0], b.visits()[0])
    qa, pa, ta = a.root_stats(0)
    qb, pb, tb = b.root_stats(0)
    assert np.array_equal(qa, qb) and np.array_equal(pa, pb) and np.array_equal(ta, tb)


@pytest.mark.parametrize("name", ["hash_4_40", "hash_6_25", "hash_6_60_c2", "hash_8_100"])
def test_selfplay_hash_prior_golden(E, golden_episodes, name):
    rec = golden_episodes[name]
    n = rec["n"]
    e = E.Engine(n, max_games=2, nodes_per_game=rec["sims"] * (n * n) + 64, prior_mode=E.PRIOR_HASH,
                 c_puct=rec["c"], log_visits=True)
    e.selfplay_begin(2, rec["sims"], temperature=rec["T"], e_greedy=1.0)
    assert e.selfplay_run(-1) == 0
    out = e.selfplay_records()
    for g in range(2):  # both games are deterministic and identical
        k = int(out["n_moves"][g])
        assert [int(a) for a in out["action"][g][:k]] == [sq8(a, n) for a in rec["moves"]]
        assert [int(p) for p in out["player"][g][:k]] == rec["players"]
        assert int(out["winner"][g]) == rec["winner"]
        for p in range(k):
            assert visits_to_grid(out["visits"][g][p], n).tolist() == rec["visits"][p]
    assert e.counters()["nodes"] == 2 * rec["net_calls"]
    e.close()


@pytest.mark.parametrize("name", ["sha_4_30", "sha_6_25", "zero_6_10", "sha_8_50"])
def test_host_prior_episode_golden(E, golden_episodes, name):
    """Host-fed NON-dyadic float32 priors: numpy's pairwise-sum order and the float32 Q arithmetic."""
    rec = golden_episodes[name]
    n = rec["n"]
    fn = prior_fns.sha_prior if name.startswith("sha") else prior_fns.zero_prior

    def predict_batch(own, opp):
        outs = [fn(canon_board(o, p, n)) for o, p in zip(own, opp)]
        return np.stack([o[0].ravel() for o in outs]), np.array([o[1] for o in outs], dtype=np.float32)

    e = E.Engine(n, max_games=1, nodes_per_game=rec["sims"] * (n * n) + 64, prior_mode=E.PRIOR_HOST, c_puct=rec["c"])
    board = oracle.initial_board(n)
    b, w = oracle.board_to_bits(board)
    e.reset(1, [b], [w], [0])
    moves = []
    cur = oracle.as_board(board).copy()
    player = 0
    for p, exp_vis in enumerate(rec["visits"]):
        e.search(rec["sims"], predict_batch)
        v, ns = e.visits()
        assert visits_to_grid(v[0], n).tolist() == exp_vis, f"move {p}"
        a = rec["moves"][p]
        moves.append(a)
        # advance with the oracle's rules (OthelloGame.play)
        nb = oracle.flip_board(cur, player, a // n, a % n)
        nxt = 1 - player
        if not oracle.valid_actions(nb, nxt):
            nxt = player if oracle.valid_actions(nb, player) else nxt
        cur, player = nb, nxt
        if oracle.has_finished(cur):
            break
        b, w = oracle.board_to_bits(cur)
        e.set_roots([b], [w], [player])
    assert e.counters()["nodes"] == rec["net_calls"]
    e.close()


def test_q_values_and_types(E, golden_episodes):
    for name in ("hash_4_40", "hash_6_25"):
        rec = golden_episodes[name]
        n = rec["n"]
        e = E.Engine(n, 1, 4096, E.PRIOR_HASH, c_puct=rec["c"])
        e.reset(1)
        e.search(rec["sims"])
        q, p, tag = e.root_stats(0)
        for k, (val, typ) in rec["q"][0].items():
            s = sq8(int(k), n)
            assert q[s] == val
            assert tag[s] == {"int": 0, "float": 1, "float32": 2}[typ]
        e.close()


def test_many_games_vs_oracle_with_random_moves(E):
    """e_greedy < 1 with the engine RNG; distinct start positions; every game checked against the oracle."""
    n, sims, G = 6, 20, 64
    starts = [oracle.playout(n, 5, g, max_moves=g % 5) for g in range(G)]
    black = [s["black"] for s in starts]
    white = [s["white"] for s in starts]
    player = [s["player"] for s in starts]
    ids = [1000 + g for g in range(G)]
    e = E.Engine(n, G, sims * 40 + 64, E.PRIOR_HASH, seed=42, log_visits=True)
    e.selfplay_begin(G, sims, 1.0, 0.7, -1, black, white, player, ids)
    assert e.selfplay_run(-1) == 0
    out = e.selfplay_records()
    for g in range(G):
        ref = oracle.execute_episode(n, sims, e_greedy=0.7, seed=42, game_id=ids[g],
                                     start_board=starts[g]["board"], start_player=player[g], log_visits=True)
        k = int(out["n_moves"][g])
        assert [int(a) for a in out["action"][g][:k]] == [sq8(a, n) for a in ref["moves"]], f"game {g}"
        assert int(out["winner"][g]) == ref["winner"]
        assert [int(x) for x in out["black"][g][:k]] == ref["black"]
        for p in range(k):
            assert visits_to_grid(out["visits"][g][p], n).tolist() == ref["visits"][p].tolist()
    e.close()


def test_pool_exhaustion_is_reported(E):
    e = E.Engine(8, 1, 16, E.PRIOR_HASH)
    e.selfplay_begin(1, 50)
    with pytest.raises(MemoryError):
        e.selfplay_run(-1)
    e.close()


def test_pool_exhaustion_is_reported_by_the_search_api(E):
    """oz_search_begin must not return thinner visit counts silently when a game's node pool fills up."""
    e = E.Engine(8, 2, 16, E.PRIOR_HASH)
    e.reset(2)
    with pytest.raises(MemoryError):
        e.search(400)
    e.close()


def test_selfplay_full_size_properties(E):
    """BASELINE configs[2] size (4096 games x 100 sims/move, 8x8) with the closed-form priors: size-independent
    invariants for every game + oracle comparison for a sample."""
    from othellozero_b200 import engine as eng_mod
    G, sims, n = 4096, 100, 8
    starts = [eng_mod.perft_playouts(G, n, seed=3, first_game_id=0, max_moves=k) for k in range(8)]
    sel = np.arange(G) % 8
    black = np.choose(sel, [s["black"] for s in starts]).astype(np.uint64)
    white = np.choose(sel, [s["white"] for s in starts]).astype(np.uint64)
    player = np.choose(sel, [s["player"] for s in starts]).astype(np.int32)
    ids = np.arange(G, dtype=np.uint64)
    e = E.Engine(n, G, sims * 61 + 64, E.PRIOR_HASH, seed=17)
    e.selfplay_begin(G, sims, 1.0, 0.9, -1, black, white, player, ids)
    assert e.selfplay_run(-1) == 0
    rec = e.selfplay_records()
    c = e.counters()
    fb, fw, fp = e.positions()
    e.close()
    nm = rec["n_moves"]
    assert (rec["winner"] >= 0).all() and (nm > 0).all() and (nm <= 60).all()
    assert c["moves"] == int(nm.sum())
    assert c["sims"] == sims * int(nm.sum())            # exactly `sims` simulations before every move
    # replay every recorded move through the K2 rules kernel: positions must chain and end at the final position
    for p in range(int(nm.max())):
        act = nm > p
        idx = np.nonzero(act)[0]
        b, w, pl = rec["black"][idx, p], rec["white"][idx, p], rec["player"][idx, p].astype(np.int64)
        own = np.where(pl == 0, b, w); opp = np.where(pl == 0, w, b)
        o2, p2, fl, _ = eng_mod.apply_moves(own, opp, rec["action"][idx, p].astype(np.int32), n)
        assert not (fl & 0x80000000).any()               # every recorded action was legal
        swapped = (fl & 1).astype(bool)
        npl = np.where(swapped, 1 - pl, pl)
        nb = np.where(npl == 0, o2, p2); nw = np.where(npl == 0, p2, o2)
        last = nm[idx] == p + 1
        assert ((fl & 4) != 0).tolist() == last.tolist() # finished exactly at the last recorded move
        nxt = idx[~last]
        assert np.array_equal(nb[~last], rec["black"][nxt, p + 1]) and np.array_equal(nw[~last], rec["white"][nxt, p + 1])
        assert np.array_equal(npl[~last], rec["player"][nxt, p + 1].astype(np.int64))
        assert np.array_equal(nb[last], fb[idx[last]]) and np.array_equal(nw[last], fw[idx[last]])
    pc = np.array([bin(int(x)).count("1") for x in fb]) >= np.array([bin(int(x)).count("1") for x in fw])
    assert np.array_equal(rec["winner"], np.where(pc, 0, 1))  # draw -> BLACK
    for g in (0, 1, 2047, 4095):
        ref = oracle.execute_episode(n, sims, e_greedy=0.9, seed=17, game_id=g,
                                     start_board=oracle.bits_to_board(int(black[g]), int(white[g]), n),
                                     start_player=int(player[g]))
        k = int(nm[g])
        assert [int(a) for a in rec["action"][g][:k]] == [sq8(a, n) for a in ref["moves"]]
        assert int(rec["winner"][g]) == ref["winner"]


def _rec_equal(a, b, ga, gb):
    k = int(a["n_moves"][ga])
    assert k == int(b["n_moves"][gb]) and int(a["winner"][ga]) == int(b["winner"][gb])
    for key in ("black", "white", "action", "player"):
        assert np.array_equal(a[key][ga][:k], b[key][gb][:k]), key
    if a.get("visits") is not None and b.get("visits") is not None:
        assert np.array_equal(a["visits"][ga][:k], b["visits"][gb][:k])


@pytest.mark.parametrize("n,sims,total,slots", [(6, 20, 37, 8), (8, 16, 50, 16)])
def test_game_queue_matches_separate_batches(E, n, sims, total, slots):
    """More episodes than slots: finished slots take the next queued game.  Every game (moves, positions, visit counts,
    winner) must equal the same game played in an engine of its own batch - the slot and its neighbours do not matter."""
    rng = np.random.default_rng(7)
    ids = rng.integers(1, 2**62, size=total, dtype=np.uint64)
    # de-correlated starts: a few random plies from the initial position
    st = E.perft_playouts(total, n, seed=11, max_moves=3)
    black, white, player = st["black"], st["white"], st["player"].astype(np.int32)
    npg = sims * (n * n) + 64
    q = E.Engine(n, max_games=slots, nodes_per_game=npg, prior_mode=E.PRIOR_HASH, log_visits=True)
    q.selfplay_begin(total, sims, 1.0, 0.8, black=black, white=white, player=player, game_ids=ids)
    assert q.selfplay_run(-1) == 0
    rq = q.selfplay_records()
    assert rq["n_moves"].shape == (total,) and (rq["winner"] >= 0).all()
    q.close()
    ref = E.Engine(n, max_games=total, nodes_per_game=npg, prior_mode=E.PRIOR_HASH, log_visits=True)
    ref.selfplay_begin(total, sims, 1.0, 0.8, black=black, white=white, player=player, game_ids=ids)
    assert ref.selfplay_run(-1) == 0
    rr = ref.selfplay_records()
    ref.close()
    for g in range(total):
        _rec_equal(rq, rr, g, g)
    # and against the oracle for a few games (hash priors, same RNG stream)
    assert int(rq["n_moves"].sum()) > total * 10


def test_game_queue_default_ids_and_partial_run(E):
    """Default game ids are the game indices (also for queued games); a step-limited run reports unfinished games."""
    n, sims, total, slots = 6, 12, 20, 4
    q = E.Engine(n, max_games=slots, nodes_per_game=sims * 36 + 64, prior_mode=E.PRIOR_HASH)
    q.selfplay_begin(total, sims, 1.0, 0.9)
    assert q.selfplay_run(-1) == 0
    rq = q.selfplay_records()
    q.close()
    ref = E.Engine(n, max_games=total, nodes_per_game=sims * 36 + 64, prior_mode=E.PRIOR_HASH)
    ref.selfplay_begin(total, sims, 1.0, 0.9)
    assert ref.selfplay_run(-1) == 0
    rr = ref.selfplay_records()
    ref.close()
    for g in range(total):
        _rec_equal(rq, rr, g, g)
    # max_moves truncation also frees the slot for the next game
    t = E.Engine(n, max_games=slots, nodes_per_game=sims * 36 + 64, prior_mode=E.PRIOR_HASH)
    t.selfplay_begin(total, sims, 1.0, 1.0, max_moves=5)
    assert t.selfplay_run(-1) == 0
    rt = t.selfplay_records()
    assert (rt["n_moves"] == 5).all() and (rt["winner"] == -1).all()
    t.close()


# ---- temperature 0 and the engine RNG (main.py:73-76 switches self-play to T = 0 after `temperature_threshold`) ----------
def test_selfplay_rng_golden(E, golden_rng_episodes):
    """Device self-play against the REFERENCE'S OWN execute_episode run with the engine's draws injected
    (tools/gen_golden.py::EngineRng): T = 0 tie-breaks over the arg-max set, e-greedy coin, random legal action."""
    groups = {}
    for rec in golden_rng_episodes:
        if rec["prior"] == "hash":
            groups.setdefault((rec["n"], rec["sims"], rec["c"], rec["T"], rec["e_greedy"], rec["seed"]), []).append(rec)
    assert sum(r["rng_calls"]["ties"] for g in groups.values() for r in g) > 40
    for (n, sims, c, T, eg, seed), recs in groups.items():
        ids = [r["game_id"] for r in recs]
        e = E.Engine(n, max_games=2, nodes_per_game=sims * n * n + 64, prior_mode=E.PRIOR_HASH, c_puct=c, seed=seed)
        e.selfplay_begin(len(ids), sims, temperature=T, e_greedy=eg, game_ids=ids)   # 2 slots: the others are queued
        assert e.selfplay_run(-1) == 0
        out = e.selfplay_records()
        assert e.counters()["nodes"] == sum(r["net_calls"] for r in recs)
        e.close()
        for g, rec in enumerate(recs):
            k = int(out["n_moves"][g])
            assert [int(a) for a in out["action"][g][:k]] == [sq8(a, n) for a in rec["moves"]], (n, sims, T, eg, seed, g)
            z = [1 if int(out["winner"][g]) == int(p) else -1 for p in out["player"][g][:k]]
            assert z == rec["z"]


@pytest.mark.parametrize("n,sims,eg", [(6, 4, 1.0), (6, 25, 1.0), (8, 6, 1.0), (8, 30, 0.85)])
def test_temperature_zero_vs_oracle(E, n, sims, eg):
    """T = 0 episodes equal the oracle move for move (positions, movers, per-move visit counts, winner); few simulations
    per move make ties in the arg-max set frequent, so the tie-break draw is exercised on most moves."""
    G, slots = 24, 8
    starts = [oracle.playout(n, 9, g, max_moves=g % 4) for g in range(G)]
    black = [s["black"] for s in starts]; white = [s["white"] for s in starts]; player = [s["player"] for s in starts]
    ids = [5000 + 3 * g for g in range(G)]
    e = E.Engine(n, slots, sims * n * n + 64, E.PRIOR_HASH, seed=77, log_visits=True)
    e.selfplay_begin(G, sims, 0.0, eg, -1, black, white, player, ids)
    assert e.selfplay_run(-1) == 0
    out = e.selfplay_records()
    e.close()
    tie_moves = 0
    for g in range(G):
        ref = oracle.execute_episode(n, sims, temperature=0.0, e_greedy=eg, seed=77, game_id=ids[g],
                                     start_board=starts[g]["board"], start_player=player[g], log_visits=True)
        k = int(out["n_moves"][g])
        assert [int(a) for a in out["action"][g][:k]] == [sq8(a, n) for a in ref["moves"]], f"game {g}"
        assert [int(p) for p in out["player"][g][:k]] == ref["players"]
        assert [int(x) for x in out["black"][g][:k]] == ref["black"] and [int(x) for x in out["white"][g][:k]] == ref["white"]
        assert int(out["winner"][g]) == ref["winner"]
        for p in range(k):
            v = visits_to_grid(out["visits"][g][p], n)
            assert v.tolist() == ref["visits"][p].tolist()
            tie_moves += int((v == v.max()).sum() > 1)
    assert tie_moves > (20 if sims <= 6 else 0)


def test_final_position_is_recorded(E):
    """Entry n_moves of a game's position row = the position the episode ended in (include/oz_b200.h)."""
    n, sims, G = 6, 8, 12
    e = E.Engine(n, 4, sims * 36 + 64, E.PRIOR_HASH, seed=3)
    e.selfplay_begin(G, sims, 1.0, 0.9)
    assert e.selfplay_run(-1) == 0
    rec = e.selfplay_records()
    e.close()
    for g in range(G):
        k = int(rec["n_moves"][g])
        pl = int(rec["player"][g][k - 1])
        b, w = int(rec["black"][g][k - 1]), int(rec["white"][g][k - 1])
        own, opp = (w, b) if pl else (b, w)
        o2, p2, fl, _ = E.apply_moves([own], [opp], [int(rec["action"][g][k - 1])], n)
        assert int(fl[0]) & 4
        fb, fw = (int(p2[0]), int(o2[0])) if pl else (int(o2[0]), int(p2[0]))
        assert (int(rec["black"][g][k]), int(rec["white"][g][k])) == (fb, fw)


def test_seed_and_game_id_do_not_commute(E):
    """(seed 1, id g) used to replay (seed 0, id g ^ 1): stream keys now mix seed and id non-commutatively."""
    n, sims, G = 6, 6, 16

    def play(seed, ids):
        e = E.Engine(n, G, sims * 36 + 64, E.PRIOR_HASH, seed=seed)
        e.selfplay_begin(G, sims, 1.0, 0.3, game_ids=ids)
        assert e.selfplay_run(-1) == 0
        r = e.selfplay_records()
        e.close()
        return {int(i): bytes(r["action"][g]) for g, i in enumerate(ids)}

    a = play(0, list(range(G)))
    b = play(1, list(range(G)))
    assert len(set(a.values())) > G // 2                       # e_greedy 0.3: games differ from each other
    assert sum(a[g] == b[g ^ 1] for g in range(G)) <= 1        # ... and seed 1 is not a permutation of seed 0
    assert len(set(a.values()) & set(b.values())) <= 1
    c = play(0, list(range(G)))
    assert a == c                                              # same (seed, ids) -> same games


def test_roots_with_more_than_32_legal_moves(E):
    """Artificial (unreachable) positions with 33-35 legal moves: a lane then scores two children per level (the rare
    second-candidate path of the descent) and the chosen square comes from the high half of the k-th-set-bit ballot.
    Root visit counts, Ns, node counts and Q values/types must equal the oracle's."""
    n, sims = 8, 300
    roots = [(0x40202c04468c2000, 0x5e400a10624c00), (0x8e42000678420, 0x76125202087200),
             (0x1108282441460180, 0x5446503220e600)]
    for own, opp in roots:
        board = oracle.bits_to_board(own, opp, n)                 # own = BLACK, BLACK to move
        k = len(oracle.valid_actions(board, 0))
        assert k > 32
        e = E.Engine(n, 1, 4 * sims, E.PRIOR_HASH)                 # pools are sized for 16 children per node: these nodes have 30+
        e.reset(1, [own], [opp], [0])
        e.search(sims)
        v, ns = e.visits()
        q, p, tag = e.root_stats(0)
        m = oracle.Mcts(n)
        for _ in range(sims):
            m.simulate(board, 0)
        rns, rv = m.visits(board)
        assert int(ns[0]) == rns and visits_to_grid(v[0], n).tolist() == rv.reshape(-1).tolist(), hex(own)
        assert e.counters()["nodes"] == m.nodes
        rq, rp, rtag = m.node_stats(board)
        for r in range(n):
            for c in range(n):
                if rtag[r, c] >= 0:
                    assert q[r * 8 + c] == rq[r, c] and p[r * 8 + c] == rp[r, c] and tag[r * 8 + c] == rtag[r, c]
        assert int((rv > 0).sum()) >= 33 - 1                      # the search really spreads over > 32 children
        e.close()
