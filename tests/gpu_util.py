import numpy as np

import oracle


def sq8(a, n):
    """oracle action r*n+c -> square bit r*8+c"""
    return (a // n) * 8 + a % n


def canon_board(own, opp, n):
    return oracle.bits_to_board(int(own), int(opp), n)


def visits_to_grid(v64, n):
    """[64] by square bit -> (n*n,) in r*n+c order"""
    return np.array([v64[r * 8 + c] for r in range(n) for c in range(n)], dtype=np.int64)
