"""CPU ORACLE — test infrastructure, not the product path.

ctypes binding of ``oz_oracle.c`` (a plain-C restatement of the reference's
rules / PUCT search / episode driver, see that file's header for citations).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this package.  ``othellozero_b200``
never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboz_oracle.so")

PREDICT_FN = C.CFUNCTYPE(None, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oz_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboz_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, ip, dp, fp = C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_float)
        u64p, lp = C.POINTER(C.c_uint64), C.POINTER(C.c_long)
        L.orc_initial_board.argtypes = [C.c_int, u8p]
        L.orc_flip_squares.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_flip_squares.restype = C.c_int
        L.orc_valid_actions.argtypes = [u8p, C.c_int, C.c_int, ip]
        L.orc_valid_actions.restype = C.c_int
        L.orc_has_actions.argtypes = [u8p, C.c_int, C.c_int]
        L.orc_has_actions.restype = C.c_int
        L.orc_flip_board.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_has_finished.argtypes = [u8p, C.c_int]
        L.orc_has_finished.restype = C.c_int
        L.orc_points.argtypes = [u8p, C.c_int, ip, ip]
        L.orc_winner.argtypes = [u8p, C.c_int, ip]
        L.orc_winner.restype = C.c_int
        L.orc_perft.argtypes = [C.c_int, C.c_int]
        L.orc_perft.restype = C.c_uint64
        L.orc_sm64.argtypes = [C.c_uint64]
        L.orc_sm64.restype = C.c_uint64
        L.orc_playout.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int, u8p, ip, ip, ip]
        L.orc_playout.restype = C.c_int
        L.orc_np_sum.argtypes = [dp, C.c_int]
        L.orc_np_sum.restype = C.c_double
        L.orc_mcts_new.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_mcts_new.restype = C.c_void_p
        L.orc_mcts_free.argtypes = [C.c_void_p]
        L.orc_mcts_net_calls.argtypes = [C.c_void_p]
        L.orc_mcts_net_calls.restype = C.c_long
        L.orc_mcts_nodes.argtypes = [C.c_void_p]
        L.orc_mcts_nodes.restype = C.c_int
        L.orc_mcts_simulate.argtypes = [C.c_void_p, u8p, C.c_int]
        L.orc_mcts_visits.argtypes = [C.c_void_p, u8p, ip]
        L.orc_mcts_visits.restype = C.c_int
        L.orc_mcts_node_stats.argtypes = [C.c_void_p, u8p, dp, dp, ip]
        L.orc_mcts_node_stats.restype = C.c_int
        L.orc_mcts_policy.argtypes = [C.c_void_p, u8p, C.c_double, dp]
        L.orc_hash_prior.argtypes = [u8p, C.c_int, fp, fp, C.c_void_p]
        L.orc_hash_prior_fn.restype = C.c_void_p
        L.orc_execute_episode.argtypes = [
            C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_double, C.c_uint64, C.c_uint64,
            u8p, C.c_int, C.c_int, u64p, u64p, ip, ip, ip, lp, lp, ip]
        L.orc_execute_episode.restype = C.c_int
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def as_board(a) -> np.ndarray:
    """(N,N,2) bool/uint8 -> contiguous uint8 copy."""
    return np.ascontiguousarray(np.asarray(a), dtype=np.uint8)


# ---- board <-> bitboard helpers (bit index r*8+c for every N <= 8) ----
def board_to_bits(board) -> tuple[int, int]:
    b = np.asarray(board)
    n = b.shape[0]
    black = white = 0
    for r in range(n):
        for c in range(n):
            if b[r, c, 0]:
                black |= 1 << (r * 8 + c)
            if b[r, c, 1]:
                white |= 1 << (r * 8 + c)
    return black, white


def bits_to_board(black: int, white: int, n: int) -> np.ndarray:
    b = np.zeros((n, n, 2), dtype=np.uint8)
    for r in range(n):
        for c in range(n):
            b[r, c, 0] = (black >> (r * 8 + c)) & 1
            b[r, c, 1] = (white >> (r * 8 + c)) & 1
    return b


def initial_board(n: int) -> np.ndarray:
    b = np.zeros((n, n, 2), dtype=np.uint8)
    lib().orc_initial_board(n, _u8(b))
    return b


def flip_squares(board, ch: int, row: int, col: int) -> np.ndarray:
    b = as_board(board)
    n = b.shape[0]
    out = np.zeros(n * n, dtype=np.uint8)
    lib().orc_flip_squares(_u8(b), n, ch, row, col, _u8(out))
    return out.reshape(n, n)


def valid_actions(board, ch: int) -> list[tuple[int, int]]:
    b = as_board(board)
    n = b.shape[0]
    out = np.zeros(64, dtype=np.int32)
    k = lib().orc_valid_actions(_u8(b), n, ch, _ip(out))
    return [(int(a) // n, int(a) % n) for a in out[:k]]


def flip_board(board, ch: int, row: int, col: int) -> np.ndarray:
    b = as_board(board).copy()
    lib().orc_flip_board(_u8(b), b.shape[0], ch, row, col)
    return b


def has_finished(board) -> bool:
    b = as_board(board)
    return bool(lib().orc_has_finished(_u8(b), b.shape[0]))


def winner(board) -> tuple[int, int]:
    b = as_board(board)
    pts = C.c_int(0)
    ch = lib().orc_winner(_u8(b), b.shape[0], C.byref(pts))
    return ch, pts.value


def perft(n: int, depth: int) -> int:
    return int(lib().orc_perft(n, depth))


def sm64(x: int) -> int:
    return int(lib().orc_sm64(C.c_uint64(x & (2**64 - 1))))


def playout(n: int, seed: int, game_id: int, max_moves: int = -1):
    """Random playout with the engine RNG. Returns dict(black, white, player, finished, moves)."""
    fb = np.zeros((n, n, 2), dtype=np.uint8)
    fp, fin = C.c_int(0), C.c_int(0)
    moves = np.zeros(128, dtype=np.int32)
    k = lib().orc_playout(n, seed, game_id, max_moves, _u8(fb), C.byref(fp), C.byref(fin), _ip(moves))
    black, white = board_to_bits(fb)
    return dict(black=black, white=white, player=fp.value, finished=bool(fin.value), moves=[int(m) for m in moves[:k]],
                board=fb)


def np_sum(a) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    return float(lib().orc_np_sum(a.ctypes.data_as(C.POINTER(C.c_double)), a.size))


def hash_prior(board):
    """Closed-form stand-in net. board canonical (N,N,2). -> (pi (N,N) f32, v f32)."""
    b = as_board(board)
    n = b.shape[0]
    pi = np.zeros(n * n, dtype=np.float32)
    v = C.c_float(0)
    lib().orc_hash_prior(_u8(b), n, pi.ctypes.data_as(C.POINTER(C.c_float)), C.byref(v), None)
    return pi.reshape(n, n), np.float32(v.value)


class Mcts:
    """OthelloMCTS restatement (othelo_mcts.py:9-88 over MCTS/__init__.py:19-187).

    ``predict``: None -> closed-form hash prior; else a Python callable
    ``predict(board (N,N,2) uint8) -> (pi (N,N) float32, v float32)``.
    """

    def __init__(self, n: int, c: float = 1.0, predict=None):
        self.n = n
        self._cb = None
        if predict is None:
            fn = lib().orc_hash_prior_fn()
        else:
            def _tramp(bp, nn, pip, vp, _user):
                board = np.ctypeslib.as_array(bp, shape=(nn, nn, 2)).copy()
                pi, v = predict(board)
                pi = np.asarray(pi, dtype=np.float32).reshape(-1)
                for i in range(nn * nn):
                    pip[i] = float(pi[i])
                vp[0] = float(np.float32(v))
            self._cb = PREDICT_FN(_tramp)
            fn = C.cast(self._cb, C.c_void_p)
        self._h = lib().orc_mcts_new(n, float(c), fn, None)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_mcts_free(self._h)
            self._h = None

    def simulate(self, board, player: int):
        b = as_board(board)
        lib().orc_mcts_simulate(self._h, _u8(b), player)

    def visits(self, canonical):
        b = as_board(canonical)
        out = np.zeros(self.n * self.n, dtype=np.int32)
        ns = lib().orc_mcts_visits(self._h, _u8(b), _ip(out))
        return ns, out.reshape(self.n, self.n)

    def node_stats(self, canonical):
        b = as_board(canonical)
        nsq = self.n * self.n
        q = np.zeros(nsq)
        p = np.zeros(nsq)
        tag = np.zeros(nsq, dtype=np.int32)
        rc = lib().orc_mcts_node_stats(self._h, _u8(b), q.ctypes.data_as(C.POINTER(C.c_double)),
                                       p.ctypes.data_as(C.POINTER(C.c_double)), _ip(tag))
        if rc != 0:
            return None
        return q.reshape(self.n, self.n), p.reshape(self.n, self.n), tag.reshape(self.n, self.n)

    def policy(self, canonical, temperature: float):
        b = as_board(canonical)
        out = np.zeros(self.n * self.n)
        lib().orc_mcts_policy(self._h, _u8(b), float(temperature), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out.reshape(self.n, self.n)

    @property
    def net_calls(self) -> int:
        return int(lib().orc_mcts_net_calls(self._h))

    @property
    def nodes(self) -> int:
        return int(lib().orc_mcts_nodes(self._h))


def execute_episode(n: int, num_sims: int, c: float = 1.0, temperature: float = 1.0, e_greedy: float = 1.0,
                    predict=None, seed: int = 0, game_id: int = 0, start_board=None, start_player: int = 0,
                    max_moves: int = -1, log_visits: bool = False):
    """training.execute_episode restatement (training.py:26-72). Returns a dict of per-move records."""
    cb = None
    if predict is None:
        fn = lib().orc_hash_prior_fn()
    else:
        def _tramp(bp, nn, pip, vp, _user):
            board = np.ctypeslib.as_array(bp, shape=(nn, nn, 2)).copy()
            pi, v = predict(board)
            pi = np.asarray(pi, dtype=np.float32).reshape(-1)
            for i in range(nn * nn):
                pip[i] = float(pi[i])
            vp[0] = float(np.float32(v))
        cb = PREDICT_FN(_tramp)
        fn = C.cast(cb, C.c_void_p)
    cap = 2 * n * n
    rb = np.zeros(cap, dtype=np.uint64)
    rw = np.zeros(cap, dtype=np.uint64)
    ra = np.zeros(cap, dtype=np.int32)
    rp = np.zeros(cap, dtype=np.int32)
    vl = np.zeros((cap, n * n), dtype=np.int32) if log_visits else None
    win, nc, sims = C.c_int(0), C.c_long(0), C.c_long(0)
    sb = None
    if start_board is not None:
        sbarr = as_board(start_board)
        sb = _u8(sbarr)
    k = lib().orc_execute_episode(
        n, fn, None, float(c), num_sims, float(temperature), float(e_greedy), seed, game_id, sb, start_player,
        max_moves, rb.ctypes.data_as(C.POINTER(C.c_uint64)), rw.ctypes.data_as(C.POINTER(C.c_uint64)), _ip(ra),
        _ip(rp), C.byref(win), C.byref(nc), C.byref(sims), _ip(vl) if vl is not None else None)
    return dict(moves=[int(a) for a in ra[:k]], players=[int(p) for p in rp[:k]],
                black=[int(x) for x in rb[:k]], white=[int(x) for x in rw[:k]], winner=win.value,
                net_calls=nc.value, sims=sims.value, visits=(vl[:k].copy() if vl is not None else None))
