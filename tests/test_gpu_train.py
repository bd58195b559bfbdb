"""The training step ("next" row) on the device: the CUDA-graph replay of the optimisation step must take
the same steps as the eager loop (same arithmetic, one launch per step)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _arrays(n, count, seed):
    rng = np.random.default_rng(seed)
    boards = (rng.random((count, n, n, 2)) < 0.3).astype(np.float32)
    boards[..., 1] *= 1.0 - boards[..., 0]
    pis = np.zeros((count, n * n), dtype=np.float32)
    pis[np.arange(count), rng.integers(0, n * n, count)] = 1.0
    return boards, pis, rng.choice([-1.0, 1.0], count).astype(np.float32)


def test_cuda_graph_step_equals_eager_step(monkeypatch):
    """Same start, same batches: [3 eager steps + 1 replay of the captured step] against [4 eager steps].  float32 convolutions
    (no TF32), no dropout.  Only ONE replayed step is compared: Adam turns every gradient component that is pure rounding
    noise into a +-lr step, and later steps amplify that - after 8 steps two equally correct implementations of the same
    layer (cuDNN's BatchNormalization in the eager loop, ATen's under graph capture) agree on only 60 % of the parameters to
    1e-4, after 40 steps on 3 % (`profiles/r2_train_variant_agreement.txt`), while their losses stay within 2 %.  So: the bulk
    of the parameters after one replayed step, the worst element by a few learning rates, and the loss of a longer run."""
    import torch
    from othellozero_b200 import net, train
    monkeypatch.setattr(train, "GRAPH_MIN_BATCHES", 1)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    n, C, lr = 6, 128, 1e-3
    blob = net.init_weights(n, C, seed=5)
    for full_batches, ragged in ((4, 0), (24, 5)):
        arrays = _arrays(n, 32 * full_batches + ragged, 6 + full_batches)
        outs = []
        for graph in (False, True):
            new, hist = train.train_blob(blob, arrays, n, C, epochs=1, batch_size=32, lr=lr, dropout=0.0, device="cuda", seed=3,
                                         cuda_graph=graph)
            assert np.isfinite(new).all() and np.isfinite(np.array(hist)).all()
            outs.append((new, np.array(hist)))
        steps = full_batches + (1 if ragged else 0)
        assert not np.array_equal(outs[0][0], blob)
        d = np.abs(outs[0][0] - outs[1][0])
        if full_batches == 4:
            assert np.mean(d < 1e-4) > 0.95, np.mean(d < 1e-4)
        assert d.max() <= 2 * lr * steps + 1e-4
        assert np.allclose(outs[0][1], outs[1][1], rtol=4e-2), (outs[0][1], outs[1][1])


def test_trained_blob_loads_into_the_device_tower():
    """B200NNet.train: fit, fold the new weights onto the device tower, predict with them (Net/NNet.py:53-87)."""
    from othellozero_b200 import net
    n, C = 6, 128
    nn_ = net.B200NNet((n, n), C, max_batch=8, blob=net.init_weights(n, C, seed=8))
    boards, pis, zs = _arrays(n, 32 * 9, 9)
    board = np.zeros((n, n, 2), dtype=bool)
    board[2, 3, 0] = board[3, 2, 0] = board[2, 2, 1] = board[3, 3, 1] = True
    before_pi, before_v = nn_.predict(board)
    nn_.train(list(zip(boards, pis.reshape(-1, n, n), zs)), epochs=1)
    after_pi, after_v = nn_.predict(board)
    assert after_pi.shape == (n, n) and abs(float(after_pi.sum()) - 1.0) < 1e-3
    assert not np.allclose(before_pi, after_pi)
