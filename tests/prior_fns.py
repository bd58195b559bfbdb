"""Deterministic stand-in networks shared by the golden generator and the parity tests.

``sha_prior`` produces NON-dyadic float32 priors (exercises numpy's pairwise np.sum order
and the float32/float64 Q arithmetic); the closed-form hash prior lives in the
oracle and in the CUDA engine.
"""
from __future__ import annotations

import hashlib

import numpy as np


def sha_prior(board):
    """board: canonical (N,N,2) bool/uint8.  -> (pi (N,N) float32 probabilities, v float32)."""
    b = np.ascontiguousarray(np.asarray(board), dtype=np.uint8)
    n = b.shape[0]
    seed = int.from_bytes(hashlib.sha1(b.tobytes()).digest()[:8], "little")
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal(n * n).astype(np.float32)
    e = np.exp(logits - logits.max()).astype(np.float32)
    pi = (e / e.sum(dtype=np.float32)).astype(np.float32)
    v = np.float32(np.tanh(rng.standard_normal()))
    return pi.reshape(n, n), v


def zero_prior(board):
    """All-zero policy: forces the 'all valid moves were masked' workaround (MCTS/__init__.py:52-55)."""
    n = np.asarray(board).shape[0]
    return np.zeros((n, n), dtype=np.float32), np.float32(0.25)
