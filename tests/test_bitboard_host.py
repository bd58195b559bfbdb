"""The product's bitboard header (host build) against the golden rules vectors and the oracle."""
import ctypes as C

import numpy as np
import pytest

import bbhost
import oracle


def _legal_list(mask):
    return [i for i in range(64) if (mask >> i) & 1]


def _forms(L, compact):
    """(legal, flip) of the unrolled templates or of the looped compact forms the tree kernel uses."""
    return (L.bbh_legal_compact, L.bbh_flip_compact) if compact else (L.bbh_legal, L.bbh_flip)


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("n", [4, 6, 8])
def test_bitboard_vs_golden(golden_rules, n, compact):
    L = bbhost.lib()
    legal, flip = _forms(L, compact)
    for rec in golden_rules[str(n)]:
        b, w = int(rec["b"], 16), int(rec["w"], 16)
        for ch, (own, opp) in ((0, (b, w)), (1, (w, b))):
            exp = rec[f"moves{ch}"]
            assert _legal_list(legal(own, opp, n)) == [m[0] for m in exp]
            for sq, fb, fw in exp:
                f = flip(sq, own, opp)
                no, np_ = own | f | (1 << sq), opp & ~f
                got = (no, np_) if ch == 0 else (np_, no)
                assert got == (int(fb, 16), int(fw, 16))
        fin = (legal(b, w, n) == 0) and (legal(w, b, n) == 0)
        assert fin == rec["finished"]


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("n", [4, 6, 8])
def test_bitboard_random_positions_vs_oracle(n, compact):
    """Arbitrary (not necessarily reachable) positions stress the flip-through quirk and edges."""
    L = bbhost.lib()
    legal, flip = _forms(L, compact)
    rng = np.random.default_rng(n)
    for _ in range(1500):
        occ = rng.random((n, n)) < rng.uniform(0.2, 0.95)
        col = rng.random((n, n)) < 0.5
        board = np.zeros((n, n, 2), dtype=np.uint8)
        board[..., 0] = occ & col
        board[..., 1] = occ & ~col
        own, opp = oracle.board_to_bits(board)
        acts = oracle.valid_actions(board, 0)
        assert _legal_list(legal(own, opp, n)) == [r * 8 + c for r, c in acts]
        for r, c in acts:
            nb = oracle.flip_board(board, 0, r, c)
            f = flip(r * 8 + c, own, opp)
            assert (own | f | (1 << (r * 8 + c)), opp & ~f) == oracle.board_to_bits(nb)


def test_play_move_turn_logic(golden_playouts):
    L = bbhost.lib()
    for rec in golden_playouts:
        n = rec["n"]
        b, w = C.c_uint64(0), C.c_uint64(0)
        L.bbh_initial(n, C.byref(b), C.byref(w))
        own, opp, player = C.c_uint64(b.value), C.c_uint64(w.value), 0
        nl = C.c_uint64(0)
        for mv in rec["moves"]:
            sq = (mv // n) * 8 + mv % n
            fl = L.bbh_play(sq, C.byref(own), C.byref(opp), n, C.byref(nl))
            if fl & 1:
                player ^= 1
        assert fl & 4
        black, white = (own.value, opp.value) if player == 0 else (opp.value, own.value)
        assert (black, white) == (int(rec["b"], 16), int(rec["w"], 16))


def test_sm64_and_kth():
    L = bbhost.lib()
    for x in (0, 1, 0xABCDEF, 2**64 - 1, 0x123456789ABCDEF0):
        assert L.bbh_sm64(x) == oracle.sm64(x)
    x = 0x8100004200001881
    idx = [i for i in range(64) if (x >> i) & 1]
    for k, i in enumerate(idx):
        assert L.bbh_kth(x, k) == i
