"""Import the (Python) reference from /root/reference for golden-vector generation and
oracle validation.  Works only in the authoring container; the GPU box has no reference.

TensorFlow/Keras are not installed, so ``Net`` / ``Net.NNet`` are stubbed before the
reference modules that import them are loaded.
"""
from __future__ import annotations

import enum
import os
import sys
import types

REF = os.environ.get("OZ_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "Othello"))


class _NeuralNets(enum.Enum):
    ONN = enum.auto()
    BNN = enum.auto()


def load():
    """Returns a namespace with Othello, MCTS, othelo_mcts, training, agents modules and NeuralNets."""
    if not available():
        raise RuntimeError("reference not present")
    if "Net.NNet" not in sys.modules:
        net = types.ModuleType("Net")
        net.__path__ = []
        nnet = types.ModuleType("Net.NNet")
        nnet.NeuralNets = _NeuralNets
        nnet.NNetWrapper = object
        net.NNet = nnet
        sys.modules["Net"] = net
        sys.modules["Net.NNet"] = nnet
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import Othello  # noqa
    import MCTS  # noqa
    import othelo_mcts  # noqa
    import agents  # noqa
    import training  # noqa
    ns = types.SimpleNamespace(Othello=Othello, MCTS=MCTS, othelo_mcts=othelo_mcts, agents=agents,
                               training=training, NeuralNets=sys.modules["Net.NNet"].NeuralNets)
    return ns


class StubNet:
    """Net contract object (Net/NNet.py:70-87): .network_type, .predict(board)->(pi (N,N) f32, v f32)."""

    def __init__(self, fn):
        self.network_type = sys.modules["Net.NNet"].NeuralNets.ONN
        self._fn = fn
        self.calls = 0

    def predict(self, board):
        self.calls += 1
        return self._fn(board)
