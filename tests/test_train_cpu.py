"""Training step ("next" row): the torch module against the independent fp32 oracle restatement, blob round trip,
and a short fit on synthetic examples (CPU)."""
import numpy as np
import torch

import oracle
from oracle import net_torch
from othellozero_b200 import net, train


def _boards(n, count, seed):
    own, opp = [], []
    for g in range(count):
        po = oracle.playout(n, seed, g, max_moves=g % (n * n - 8))
        b, w, p = po["black"], po["white"], po["player"]
        own.append(w if p else b)
        opp.append(b if p else w)
    return net_torch.boards_from_bits(own, opp, n)


def test_blob_round_trip_and_forward_matches_oracle():
    n, C = 6, 128
    blob = net.init_weights(n, C, seed=3, randomize_bn=True)
    m = train.OthelloNNTorch(n, C).load_blob(blob).eval()
    assert np.array_equal(m.to_blob(), blob)
    x = _boards(n, 12, 1)
    with torch.no_grad():
        logits, v = m(torch.from_numpy(x))
    rpi, rlg, rv = net_torch.forward(blob, x, n, C)
    assert np.abs(logits.numpy() - rlg).max() < 1e-4 and np.abs(v.numpy() - rv).max() < 1e-5


def test_fit_reduces_loss_and_changes_bn_statistics():
    n, C = 6, 128
    blob = net.init_weights(n, C, seed=0)
    x = _boards(n, 96, 2)
    rng = np.random.default_rng(0)
    examples = []
    for b in x:
        pol = np.zeros((n, n)); pol[rng.integers(n), rng.integers(n)] = 1
        examples.append((b, pol, 1 if b[..., 0].sum() >= b[..., 1].sum() else -1))
    new_blob, hist = train.train_blob(blob, examples, n, C, epochs=4, batch_size=32, device="cpu")
    assert hist[-1][0] < hist[0][0]
    assert new_blob.shape == blob.shape and not np.array_equal(new_blob, blob)
    w0, w1 = net.unpack_blob(blob, n, C), net.unpack_blob(new_blob, n, C)
    assert not np.array_equal(w0["bn1.mean"], w1["bn1.mean"])  # moving statistics are part of the checkpoint
    assert np.isfinite(new_blob).all()
