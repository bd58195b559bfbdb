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


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_workers():
    """The reference's workers.py (Worker / ThreadWorker / WorkerManager / WorkType) with its cloud-only imports
    stubbed: humanfriendly, gcloud (googleapiclient + ssh plumbing) and pickle_training (the remote entry point)."""
    load()
    if "workers" not in sys.modules:
        def _cloud_only(*a, **k):
            raise RuntimeError("GCE plumbing is not available in the test harness")
        if "humanfriendly" not in sys.modules:
            _stub_module("humanfriendly", format_size=lambda n: f"{n} B")
        _stub_module("gcloud", get_instance=_cloud_only, ssh_connection=_cloud_only, get_instance_external_ip=_cloud_only,
                     get_instance_internal_ip=_cloud_only, SSH_USER="nobody")
        _stub_module("pickle_training", pack_arguments_to_pickle=_cloud_only, unpack_base64_pickle=_cloud_only)
    import workers  # noqa
    return workers


def load_main():
    """The reference's main.py (CircularArray, training loop) - needs the workers stubs and a `Net.NNet.NNetWrapper`."""
    load_workers()
    if "googleapiclient.discovery" not in sys.modules:
        g = sys.modules.get("googleapiclient") or _stub_module("googleapiclient")
        g.__path__ = []
        g.discovery = _stub_module("googleapiclient.discovery")
    import main  # noqa
    return main
