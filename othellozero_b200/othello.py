"""Rules side of the drop-in: the reference's ``Othello`` module API (Othello/__init__.py) with every
rule evaluated by the CUDA bitboard kernels (K1/K2) through the C-ABI.

Same names, argument meaning and error behaviour as the reference so that training.py / agents.py /
othelo_mcts.py keep working when they import ``OthelloGame`` from here.  Boards are the reference's
``(N,N,2)`` bool arrays (channel 0 = BLACK, 1 = WHITE); each call converts to two uint64 bitboards and
launches a kernel — convenient, not fast.  The fast path is the batched engine (selfplay.py).

Deliberate deviations are listed in INTEGRATION.md (e.g. ``is_square_free`` works here; the reference's
version raises NameError, Othello/__init__.py:88-98).
"""
from __future__ import annotations

from enum import Enum, auto

import numpy as np

from . import engine as _e
from ._lib import MOVE_FINISHED, MOVE_SWAPPED


class BoardView(Enum):
    ONE_CHANNEL = auto()
    TWO_CHANNELS = auto()


class OthelloPlayer(Enum):
    """Othello/__init__.py:12-18."""
    BLACK = 1
    WHITE = -1

    @property
    def opponent(self):
        return OthelloPlayer.WHITE if self is OthelloPlayer.BLACK else OthelloPlayer.BLACK


def _bits(board) -> tuple[int, int]:
    b = np.asarray(board)
    n = b.shape[0]
    black = white = 0
    rr, cc = np.nonzero(b[..., 0])
    for r, c in zip(rr, cc):
        black |= 1 << (int(r) * 8 + int(c))
    rr, cc = np.nonzero(b[..., 1])
    for r, c in zip(rr, cc):
        white |= 1 << (int(r) * 8 + int(c))
    return black, white


def _mask_to_squares(mask: int, n: int):
    return [(s >> 3, s & 7) for s in range(64) if (mask >> s) & 1 and (s >> 3) < n and (s & 7) < n]


class OthelloGame:
    PLAYER_CHANNELS = {OthelloPlayer.BLACK: 0, OthelloPlayer.WHITE: 1}
    ALL_DIRECTIONS = np.array([(1, 1), (1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1), (0, 1)])
    device = 0  # CUDA ordinal used by the static rule calls

    def __init__(self, board_size=8, initial_board=None, current_player=OthelloPlayer.BLACK):
        """Othello/__init__.py:29-59."""
        assert board_size % 2 == 0, 'Board size must be even'
        assert initial_board is None or initial_board.shape == (board_size, board_size, 2), \
            f'Expecting initial board shape ({board_size}, {board_size}, 2)'
        self._board = initial_board if initial_board is not None else self.initial_board(board_size)
        self._board_size = board_size
        self._round = 1
        self.current_player = current_player
        self._one_channel_board_last_update = None
        self._one_channel_board = None
        self._has_finished = OthelloGame.has_board_finished(self._board) if initial_board is not None else False

    # ---- instance API (Othello/__init__.py:61-175) ----------------------------------------------
    @property
    def board_size(self):
        return self._board_size

    @property
    def round(self):
        return self._round

    def board(self, view=BoardView.ONE_CHANNEL):
        if view == BoardView.TWO_CHANNELS:
            return self._board  # the live array, as in the reference (:77-78)
        elif view == BoardView.ONE_CHANNEL:
            if self._one_channel_board_last_update != self.round:
                self._one_channel_board = OthelloGame.convert_to_one_channel_board(self._board)
                self._one_channel_board_last_update = self.round
            return self._one_channel_board
        raise TypeError('Expecting BoardView type')

    def is_square_free(self, row, col):
        return OthelloGame.is_board_square_free(self._board, row, col)

    def is_valid_action(self, row, col):
        return OthelloGame.is_valid_player_action(self._board, self.current_player, row, col)

    def get_valid_actions(self):
        return OthelloGame.get_player_valid_actions(self._board, self.current_player)

    def get_free_squares(self):
        return OthelloGame.get_board_free_squares(self._board)

    def has_finished(self):
        return self._has_finished

    def play(self, row, col):
        """Othello/__init__.py:136-159 — one K2 launch does the flips and the pass/terminal test."""
        assert not self._has_finished, 'Game has ended'
        n = self._board_size
        black, white = _bits(self._board)
        own, opp = (black, white) if self.current_player is OthelloPlayer.BLACK else (white, black)
        o2, p2, fl, _ = _e.apply_moves([own], [opp], [int(row) * 8 + int(col)], n, self.device)
        fl = int(fl[0])
        if fl & 0x80000000:
            # the reference does not validate: it places the disc and flips nothing (:237-247)
            o2, p2, fl = self._play_unchecked(own, opp, int(row) * 8 + int(col))
        swapped = bool(fl & MOVE_SWAPPED)
        mover_bits, other_bits = (int(p2[0]), int(o2[0])) if swapped else (int(o2[0]), int(p2[0]))
        nb, nw = (mover_bits, other_bits) if self.current_player is OthelloPlayer.BLACK else (other_bits, mover_bits)
        self._write_bits(nb, nw)
        self._round += 1
        if fl & MOVE_FINISHED:
            self.current_player = self.current_player.opponent  # :147 — switched, never switched back
            self._has_finished = True
        elif swapped:
            self.current_player = self.current_player.opponent

    def _play_unchecked(self, own, opp, sq):
        own |= 1 << sq
        opp &= ~(1 << sq)
        n = self._board_size
        lo = int(_e.legal_moves([opp], [own], n, self.device)[0])
        if lo:
            return np.array([opp], dtype=np.uint64), np.array([own], dtype=np.uint64), MOVE_SWAPPED
        lm = int(_e.legal_moves([own], [opp], n, self.device)[0])
        return np.array([own], dtype=np.uint64), np.array([opp], dtype=np.uint64), (2 if lm else MOVE_FINISHED)

    def _write_bits(self, black, white):
        n = self._board_size
        for r in range(n):
            for c in range(n):
                self._board[r, c, 0] = (black >> (r * 8 + c)) & 1
                self._board[r, c, 1] = (white >> (r * 8 + c)) & 1

    def get_players_points(self):
        return OthelloGame.get_board_players_points(self._board)

    def get_winning_player(self):
        return OthelloGame.get_board_winning_player(self._board)

    # ---- static array API (Othello/__init__.py:177-274) -----------------------------------------
    @staticmethod
    def initial_board(board_size):
        assert board_size % 2 == 0, 'Board size must be even'
        initial = np.array([[[0, 1], [1, 0]], [[1, 0], [0, 1]]], dtype=bool)
        pad = (board_size - 2) // 2
        return np.pad(initial, ((pad, pad), (pad, pad), (0, 0)), constant_values=0)

    @staticmethod
    def get_all_directions_squares(board_size, row, col):
        for direction in OthelloGame.ALL_DIRECTIONS:
            yield OthelloGame.get_direction_squares(board_size, direction, row, col)

    @staticmethod
    def get_direction_squares(board_size, direction, row, col):
        row_offset, col_offset = direction
        row, col = row + row_offset, col + col_offset
        while 0 <= row < board_size and 0 <= col < board_size:
            yield row, col
            row += row_offset
            col += col_offset

    @staticmethod
    def get_board_free_squares(board):
        return np.argwhere(np.amax(board, axis=2) == 0)

    @staticmethod
    def is_board_square_free(board, row, col):
        return np.amax(board[row, col]) == 0

    @staticmethod
    def _own_opp(board, player):
        if not isinstance(player, OthelloPlayer):
            raise TypeError('Expecting OthelloPlayer type')
        black, white = _bits(board)
        return (black, white) if player is OthelloPlayer.BLACK else (white, black)

    @staticmethod
    def legal_mask(board, player) -> int:
        own, opp = OthelloGame._own_opp(board, player)
        return int(_e.legal_moves([own], [opp], np.asarray(board).shape[0], OthelloGame.device)[0])

    @staticmethod
    def get_player_valid_actions(board, player):
        """Generator of [row, col] arrays in row-major order (:208-210)."""
        n = np.asarray(board).shape[0]
        mask = OthelloGame.legal_mask(board, player)
        return (np.array([r, c]) for r, c in _mask_to_squares(mask, n))

    @staticmethod
    def is_valid_player_action(board, player, row, col):
        return bool((OthelloGame.legal_mask(board, player) >> (int(row) * 8 + int(col))) & 1)

    @staticmethod
    def get_action_flip_squares(board, player, row, col):
        """The SET of flipped squares in row-major order (the reference's generator also yields duplicates,
        :216-235; every consumer only uses membership)."""
        n = np.asarray(board).shape[0]
        own, opp = OthelloGame._own_opp(board, player)
        sq = int(row) * 8 + int(col)
        if (own | opp) >> sq & 1:
            return iter(())
        o2, p2, fl, _ = _e.apply_moves([own], [opp], [sq], n, OthelloGame.device)
        if int(fl[0]) & 0x80000000:
            return iter(())
        mover = int(p2[0]) if int(fl[0]) & MOVE_SWAPPED else int(o2[0])
        flipped = mover & ~own & ~(1 << sq)
        return iter(_mask_to_squares(flipped, n))

    @staticmethod
    def flip_board_squares(board, player, row, col):
        """In place, like the reference (:237-247)."""
        player_channel = OthelloGame.PLAYER_CHANNELS[player]
        opponent_channel = OthelloGame.PLAYER_CHANNELS[player.opponent]
        for flip_row, flip_col in OthelloGame.get_action_flip_squares(board, player, row, col):
            board[flip_row, flip_col, player_channel] = 1
            board[flip_row, flip_col, opponent_channel] = 0
        board[row, col, player_channel] = 1
        board[row, col, opponent_channel] = 0

    @staticmethod
    def has_board_finished(board):
        return not OthelloGame.has_player_actions_on_board(board, OthelloPlayer.BLACK) and \
            not OthelloGame.has_player_actions_on_board(board, OthelloPlayer.WHITE)

    @staticmethod
    def get_board_winning_player(board):
        return max(OthelloGame.get_board_players_points(board).items(), key=lambda item: item[1])

    @staticmethod
    def get_board_players_points(board):
        black, white = _bits(board)
        cb, cw = _e.score([black], [white], OthelloGame.device)
        return {OthelloPlayer.BLACK: int(cb[0]), OthelloPlayer.WHITE: int(cw[0])}

    @staticmethod
    def has_player_actions_on_board(board, player):
        return OthelloGame.legal_mask(board, player) != 0

    @staticmethod
    def convert_to_one_channel_board(board):
        one_channel = board[:, :, 0] * OthelloPlayer.BLACK.value
        one_channel = one_channel + board[:, :, 1] * OthelloPlayer.WHITE.value
        return one_channel

    @staticmethod
    def invert_board(board):
        return np.flip(board, axis=2)

    # ---- alpha-zero-general style aliases named in BASELINE.json --------------
    getInitBoard = initial_board

    @staticmethod
    def getValidMoves(board, player):
        n = np.asarray(board).shape[0]
        mask = OthelloGame.legal_mask(board, player)
        out = np.zeros((n, n), dtype=np.float64)
        for r, c in _mask_to_squares(mask, n):
            out[r, c] = 1
        return out

    @staticmethod
    def getNextState(board, player, action):
        """-> (next_board, next_player) with OthelloGame.play's pass handling."""
        b = np.array(board, copy=True)
        g = OthelloGame(b.shape[0], initial_board=b, current_player=player)
        g._has_finished = False
        g.play(*action)
        return g.board(BoardView.TWO_CHANNELS), g.current_player

    @staticmethod
    def getGameEnded(board):
        """0 if not ended, else the winner's OthelloPlayer.value (draw -> BLACK)."""
        if not OthelloGame.has_board_finished(board):
            return 0
        return OthelloGame.get_board_winning_player(board)[0].value

    @staticmethod
    def getCanonicalForm(board, player):
        return OthelloGame.invert_board(board) if player is OthelloPlayer.WHITE else board
