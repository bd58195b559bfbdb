"""Generate tests/golden/*.json by EXECUTING the reference (imported from /root/reference).

Run in the authoring container only:  python tools/gen_golden.py
The fixtures pin the oracle (oracle/oz_oracle.c) and, through it, the CUDA kernels.
Board encoding in fixtures: hex bitboards, bit r*8+c, "b" = BLACK (ch0), "w" = WHITE (ch1).
"""
from __future__ import annotations

import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from tools import ref_loader  # noqa: E402
import prior_fns  # noqa: E402
from example_digest import examples_digest  # noqa: E402

R = ref_loader.load()
Game = R.Othello.OthelloGame
Player = R.Othello.OthelloPlayer
M64 = 2**64 - 1


def sm64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def bits(board):
    n = board.shape[0]
    b = w = 0
    for r in range(n):
        for c in range(n):
            if board[r, c, 0]:
                b |= 1 << (r * 8 + c)
            if board[r, c, 1]:
                w |= 1 << (r * 8 + c)
    return b, w


def hx(x):
    return f"{x:016x}"


def hash_prior(board):
    n = board.shape[0]
    own, opp = bits(board)
    key = sm64(own ^ sm64(opp))
    pi = np.zeros((n, n), dtype=np.float32)
    for r in range(n):
        for c in range(n):
            pi[r, c] = np.float32(((sm64((key + r * 8 + c) & M64) >> 48) + 1) / 65536)
    v = np.float32(((sm64(key ^ 0xABCDEF) >> 48) - 32768) / 32768)
    return pi, v


def gen_rules(n, n_games, rng):
    """Random games through OthelloGame.play; at each position record both players' legal
    moves and every resulting flipped board, plus terminal/winner info."""
    out = []
    for _ in range(n_games):
        g = Game(n)
        while not g.has_finished():
            board = g.board(R.Othello.BoardView.TWO_CHANNELS)
            b, w = bits(board)
            rec = {"b": hx(b), "w": hx(w), "to_move": 0 if g.current_player is Player.BLACK else 1}
            for ch, pl in ((0, Player.BLACK), (1, Player.WHITE)):
                acts = [tuple(int(x) for x in a) for a in Game.get_player_valid_actions(board, pl)]
                res = []
                for (r, c) in acts:
                    nb = np.copy(board)
                    Game.flip_board_squares(nb, pl, r, c)
                    fb, fw = bits(nb)
                    res.append([r * 8 + c, hx(fb), hx(fw)])
                rec[f"moves{ch}"] = res
            rec["finished"] = bool(Game.has_board_finished(board))
            wp, pts = Game.get_board_winning_player(board)
            rec["winner"] = 0 if wp is Player.BLACK else 1
            rec["points"] = int(pts)
            out.append(rec)
            acts = list(g.get_valid_actions())
            a = acts[rng.randrange(len(acts))]
            g.play(int(a[0]), int(a[1]))
        board = g.board(R.Othello.BoardView.TWO_CHANNELS)
        b, w = bits(board)
        wp, pts = Game.get_board_winning_player(board)
        out.append({"b": hx(b), "w": hx(w), "to_move": 0 if g.current_player is Player.BLACK else 1,
                    "moves0": [], "moves1": [], "finished": True, "winner": 0 if wp is Player.BLACK else 1,
                    "points": int(pts)})
    return out


def stream_key(seed, gid):
    """Engine spec (include/oz_b200.h): seed and game id mixed non-commutatively."""
    return sm64((sm64(seed) + gid) & M64)


def episode_draw(base, p, which):
    return sm64((base + 4 * p + which) & M64)


def pick_index(z, cnt):
    return ((z >> 32) * cnt) >> 32


def gen_playouts(n, seed, ids):
    """Engine-RNG playouts replayed through the reference's OthelloGame.play."""
    out = []
    for gid in ids:
        g = Game(n)
        base = stream_key(seed, gid)
        p = 0
        moves = []
        while not g.has_finished():
            acts = [tuple(int(x) for x in a) for a in g.get_valid_actions()]
            z = sm64((base + p) & M64)
            k = pick_index(z, len(acts))
            r, c = acts[k]
            moves.append(r * n + c)
            g.play(r, c)
            p += 1
        b, w = bits(g.board(R.Othello.BoardView.TWO_CHANNELS))
        wp, pts = g.get_winning_player()
        out.append({"n": n, "seed": seed, "game_id": gid, "moves": moves, "b": hx(b), "w": hx(w),
                    "winner": 0 if wp is Player.BLACK else 1, "points": int(pts)})
    return out


def run_episode(n, sims, prior, c=1, T=1):
    """training.execute_episode semantics with e_greedy=1.0 (deterministic), logging root visit
    counts per move (the reference function itself does not expose them)."""
    net = ref_loader.StubNet(prior)
    g = Game(n)
    mcts = R.othelo_mcts.OthelloMCTS(n, net, c)
    moves, visits, players, qlog = [], [], [], []
    while not g.has_finished():
        state = g.board(R.Othello.BoardView.TWO_CHANNELS)
        for _ in range(sims):
            mcts.simulate(state, g.current_player)
        if g.current_player == Player.WHITE:
            state = Game.invert_board(state)
        policy = mcts.get_policy_action_probabilities(state, T)
        action = np.argwhere(policy == policy.max())[0]
        cnt = [0] * (n * n)
        qs = {}
        for a in mcts.get_state_actions(state):
            cnt[a[0] * n + a[1]] = int(mcts.N(state, a))
            h = R.MCTS.hash_ndarray(state)
            q = mcts._Qsa[h][a]
            qs[a[0] * n + a[1]] = [float(q), type(q).__name__]
        visits.append(cnt)
        qlog.append(qs)
        players.append(0 if g.current_player is Player.BLACK else 1)
        moves.append(int(action[0]) * n + int(action[1]))
        g.play(int(action[0]), int(action[1]))
    wp, pts = g.get_winning_player()
    b, w = bits(g.board(R.Othello.BoardView.TWO_CHANNELS))
    return {"n": n, "sims": sims, "c": c, "T": T, "moves": moves, "players": players, "visits": visits,
            "q": [{str(k): v for k, v in d.items()} for d in qlog],
            "winner": 0 if wp is Player.BLACK else 1, "net_calls": net.calls, "b": hx(b), "w": hx(w)}


def check_execute_episode(n, sims, prior, expect_moves):
    """The reference's own execute_episode (e_greedy=1.0) must replay the same moves."""
    net = ref_loader.StubNet(prior)
    ex = R.training.execute_episode(n, net, 1, sims, 1, 1.0)
    assert len(ex) == 8 * len(expect_moves), (len(ex), len(expect_moves))
    got = []
    for i in range(len(expect_moves)):
        board, pol, z = ex[8 * i + 7]  # identity symmetry is the last of each group of 8
        a = int(np.argmax(pol))
        got.append(a)
    assert got == expect_moves, (got, expect_moves)
    return net.calls


class EngineRng:
    """The engine's counter RNG injected into the reference's three draw sites, so that the REFERENCE'S OWN
    execute_episode (training.py:26-72) plays the episode the engine plays for (seed, game_id):
      random.choice(bests)       othelo_mcts.py:58-59 (T == 0)   -> draw 2 of the current move
      random.random()            training.py:51                   -> draw 0
      np.random.choice(len(..))  training.py:56                   -> draw 1
    """

    def __init__(self, seed, game_id):
        self.base = stream_key(seed ^ 0x5EEDC01D, game_id)
        self.p = 0          # index of the move being decided
        self.calls = {"choice": 0, "random": 0, "np_choice": 0, "ties": 0}

    # random.choice
    def choice(self, seq):
        self.calls["choice"] += 1
        if len(seq) > 1:
            self.calls["ties"] += 1
            return seq[pick_index(episode_draw(self.base, self.p, 2), len(seq))]
        return seq[0]

    # random.random
    def random(self):
        self.calls["random"] += 1
        coin = (episode_draw(self.base, self.p, 0) >> 11) * (1.0 / 9007199254740992.0)
        self.p += 1
        return coin

    # np.random.choice(k) -- called after random() of the same move
    def np_choice(self, k):
        self.calls["np_choice"] += 1
        return pick_index(episode_draw(self.base, self.p - 1, 1), int(k))


class _NpShim:
    """`np` as seen by training.py with np.random.choice replaced."""

    def __init__(self, rng):
        class _R:
            choice = staticmethod(rng.np_choice)
        self.random = _R

    def __getattr__(self, name):
        return getattr(np, name)


def reference_episode_with_engine_rng(n, sims, prior, c, T, e_greedy, seed, game_id):
    """Runs the reference's execute_episode with the engine RNG injected; returns (examples, rng, net)."""
    rng = EngineRng(seed, game_id)
    net = ref_loader.StubNet(prior)
    saved = (R.othelo_mcts.random, R.training.random, R.training.np)
    R.othelo_mcts.random = rng
    R.training.random = rng
    R.training.np = _NpShim(rng)
    try:
        ex = R.training.execute_episode(n, net, c, sims, T, e_greedy)
    finally:
        R.othelo_mcts.random, R.training.random, R.training.np = saved
    return ex, rng, net


def gen_rng_episode(n, sims, prior_name, c, T, e_greedy, seed, game_id):
    prior = {"hash": hash_prior, "sha": prior_fns.sha_prior}[prior_name]
    ex, rng, net = reference_episode_with_engine_rng(n, sims, prior, c, T, e_greedy, seed, game_id)
    k = len(ex) // 8
    moves, zs = [], []
    for i in range(k):
        _, pol, z = ex[8 * i + 7]           # identity symmetry is the last of each group of 8
        moves.append(int(np.argmax(pol)))
        zs.append(int(z))
    return {"n": n, "sims": sims, "prior": prior_name, "c": c, "T": T, "e_greedy": e_greedy, "seed": seed,
            "game_id": game_id, "moves": moves, "z": zs, "net_calls": net.calls, "rng_calls": rng.calls}


def gen_examples():
    """The reference's own example stream (training.py:58-72) for deterministic episodes: digests of the stream as
    returned (every board is a view of the live board -> the FINAL position) and of the same stream with
    true per-move snapshots (boards rebuilt from the moves)."""
    out = []
    for n, sims, prior_name, T, eg, seed, gid in ((6, 25, "hash", 1, 1.0, 0, 0), (6, 25, "sha", 0, 0.7, 3, 5),
                                                   (8, 100, "hash", 1, 1.0, 0, 0), (4, 30, "sha", 1, 0.5, 9, 2),
                                                   (6, 6, "hash", 0, 0.8, 21, 4), (8, 8, "hash", 0, 0.9, 22, 7)):
        prior = {"hash": hash_prior, "sha": prior_fns.sha_prior}[prior_name]
        ex, rng, net = reference_episode_with_engine_rng(n, sims, prior, 1, T, eg, seed, gid)
        k = len(ex) // 8
        moves = [int(np.argmax(ex[8 * i + 7][1])) for i in range(k)]
        # true snapshots: replay the moves through the reference's OthelloGame
        g = Game(n)
        snap = []
        for i in range(k):
            board = np.copy(g.board(R.Othello.BoardView.TWO_CHANNELS))
            pol = np.zeros((n, n))
            pol[moves[i] // n][moves[i] % n] = 1
            z = ex[8 * i][2]
            for b2, p2 in R.training.training_example_symmetries(board, pol):
                snap.append((b2, p2, z))
            g.play(moves[i] // n, moves[i] % n)
        out.append({"n": n, "sims": sims, "prior": prior_name, "T": T, "e_greedy": eg, "seed": seed, "game_id": gid,
                    "moves": moves, "n_examples": len(ex), "sha256_reference_stream": examples_digest(ex),
                    "sha256_snapshot_stream": examples_digest(snap)})
    return out


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    parts = set(sys.argv[1:]) or {"rules", "playouts", "episodes", "roots", "rng", "examples"}
    rng = random.Random(20261018)
    t0 = time.time()
    if "rng" in parts:
        eps = []
        for n, sims, pr, c, T, eg, seed, gids in (
                (6, 25, "hash", 1, 0, 1.0, 0, range(6)), (6, 25, "hash", 1, 0, 0.8, 1, range(4)),
                (6, 25, "sha", 1, 1, 0.6, 2, range(3)), (4, 20, "hash", 2, 0, 0.9, 7, range(6)),
                (8, 100, "hash", 1, 0, 1.0, 0, range(1)), (8, 40, "sha", 1, 0, 0.9, 5, range(2)),
                # few simulations per move: the arg-max set is rarely a single action, so the tie-break draw decides
                (6, 4, "hash", 1, 0, 1.0, 11, range(6)), (6, 6, "sha", 1, 0, 0.9, 12, range(4)),
                (8, 5, "hash", 1, 0, 1.0, 13, range(3)), (4, 3, "hash", 1, 0, 0.8, 14, range(4))):
            for gid in gids:
                eps.append(gen_rng_episode(n, sims, pr, c, T, eg, seed, gid))
                print("rng", n, sims, pr, T, eg, seed, gid, eps[-1]["rng_calls"], f"{time.time()-t0:.1f}s")
        with open(os.path.join(ROOT, "tests", "golden", "episodes_rng.json"), "w") as f:
            json.dump(eps, f, separators=(",", ":"))
    if "examples" in parts:
        with open(os.path.join(ROOT, "tests", "golden", "examples.json"), "w") as f:
            json.dump(gen_examples(), f, separators=(",", ":"))
        print("examples", f"{time.time()-t0:.1f}s")
    if parts & {"rules", "playouts", "episodes", "roots"}:
        main_base(rng, t0, parts)


def main_base(rng, t0, parts):
    if "rules" in parts:  # the only part that consumes `rng`
        rules = {"8": gen_rules(8, 6, rng), "6": gen_rules(6, 8, rng), "4": gen_rules(4, 10, rng)}
        with open(os.path.join(ROOT, "tests", "golden", "rules.json"), "w") as f:
            json.dump(rules, f, separators=(",", ":"))
        print("rules", {k: len(v) for k, v in rules.items()}, f"{time.time()-t0:.1f}s")
    if "playouts" in parts:
        gen_playouts_file(t0)
    if "episodes" in parts:
        gen_episodes_file(t0)
    if "roots" in parts:
        gen_roots_file(t0)


def gen_playouts_file(t0):
    playouts = gen_playouts(8, 0, list(range(12))) + gen_playouts(6, 0, list(range(12))) + \
        gen_playouts(8, 12345, [1000000, 1048575]) + gen_playouts(4, 7, list(range(8)))
    with open(os.path.join(ROOT, "tests", "golden", "playouts.json"), "w") as f:
        json.dump(playouts, f, separators=(",", ":"))
    print("playouts", len(playouts), f"{time.time()-t0:.1f}s")



def gen_episodes_file(t0):
    eps = {}
    eps["hash_6_25"] = run_episode(6, 25, hash_prior)
    eps["hash_6_25"]["execute_episode_net_calls"] = check_execute_episode(6, 25, hash_prior, eps["hash_6_25"]["moves"])
    print("hash_6_25", f"{time.time()-t0:.1f}s")
    eps["hash_4_40"] = run_episode(4, 40, hash_prior)
    eps["sha_4_30"] = run_episode(4, 30, prior_fns.sha_prior)
    eps["sha_6_25"] = run_episode(6, 25, prior_fns.sha_prior)
    eps["zero_6_10"] = run_episode(6, 10, prior_fns.zero_prior)
    eps["hash_6_60_c2"] = run_episode(6, 60, hash_prior, c=2)
    print("6x6 done", f"{time.time()-t0:.1f}s")
    eps["hash_8_100"] = run_episode(8, 100, hash_prior)
    print("hash_8_100", f"{time.time()-t0:.1f}s")
    eps["sha_8_50"] = run_episode(8, 50, prior_fns.sha_prior)
    print("sha_8_50", f"{time.time()-t0:.1f}s")
    with open(os.path.join(ROOT, "tests", "golden", "episodes.json"), "w") as f:
        json.dump(eps, f, separators=(",", ":"))



def gen_roots_file(t0):
    # root visit counts after k sims from the initial position (Appendix B.3)
    roots = []
    for n, sims in ((6, 25), (8, 100), (8, 800)):
        net = ref_loader.StubNet(hash_prior)
        g = Game(n)
        mcts = R.othelo_mcts.OthelloMCTS(n, net, 1)
        st = g.board(R.Othello.BoardView.TWO_CHANNELS)
        for _ in range(sims):
            mcts.simulate(st, Player.BLACK)
        cnt = [0] * (n * n)
        for a in mcts.get_state_actions(st):
            cnt[a[0] * n + a[1]] = int(mcts.N(st, a))
        roots.append({"n": n, "sims": sims, "visits": cnt, "ns": int(mcts.N(st)), "net_calls": net.calls})
    with open(os.path.join(ROOT, "tests", "golden", "roots.json"), "w") as f:
        json.dump(roots, f, separators=(",", ":"))
    print("done", f"{time.time()-t0:.1f}s")


if __name__ == "__main__":
    main()
