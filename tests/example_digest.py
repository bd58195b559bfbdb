"""Digest of an example stream exactly as training.execute_episode returns it (training.py:72): shared by the golden
generator (tools/gen_golden.py, run against the live reference) and the parity tests."""
from __future__ import annotations

import hashlib

import numpy as np

import oracle


def examples_digest(examples) -> str:
    """SHA-256 over, for each example in order: board dtype/shape/bytes ((N,N,2) bool), policy dtype/shape/bytes
    ((N,N) float64) and z."""
    h = hashlib.sha256()
    for board, pol, z in examples:
        b = np.ascontiguousarray(board)
        p = np.ascontiguousarray(pol)
        h.update(str(b.dtype).encode() + str(b.shape).encode() + b.tobytes())
        h.update(str(p.dtype).encode() + str(p.shape).encode() + p.tobytes())
        h.update(str(int(z)).encode())
    return h.hexdigest()


def oracle_episode_as_records(out: dict, n: int) -> dict:
    """oracle.execute_episode's result in the layout of oz_selfplay_get_records for ONE game (include/oz_b200.h):
    actions as square bits r*8+c, entry n_moves of the position row = the final position."""
    k = len(out["moves"])
    rec = dict(black=np.zeros((1, 64), np.uint64), white=np.zeros((1, 64), np.uint64),
               action=np.full((1, 64), 0xFF, np.uint8), player=np.zeros((1, 64), np.uint8),
               n_moves=np.array([k], np.int32), winner=np.array([out["winner"]], np.int32))
    rec["black"][0, :k] = np.array(out["black"], dtype=np.uint64)
    rec["white"][0, :k] = np.array(out["white"], dtype=np.uint64)
    rec["action"][0, :k] = [(a // n) * 8 + a % n for a in out["moves"]]
    rec["player"][0, :k] = out["players"]
    if k:
        last = oracle.bits_to_board(out["black"][-1], out["white"][-1], n)
        a = out["moves"][-1]
        fb, fw = oracle.board_to_bits(oracle.flip_board(last, out["players"][-1], a // n, a % n))
        rec["black"][0, k], rec["white"][0, k] = fb, fw
    return rec
