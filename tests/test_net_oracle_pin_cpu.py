"""Pins the network oracle (oracle/net_torch.py, the fp32 restatement the CUDA tower is checked against) as far as
this image allows (TensorFlow/Keras absent): (1) hand-derived known answers for the Keras layer semantics - HWIO
cross-correlation, 'same' padding, BN epsilon 1e-3, (h,w,c) flatten, Dense (in,out) - that both restatements must
reproduce, each failing on the classic mistakes (flipped / transposed kernel, (c,h,w) flatten, torch's 1e-5 epsilon);
(2) agreement of the torch restatement with the independent NumPy one (oracle/net_numpy.py) to 1e-5 on random
asymmetric weights.  Net/OthelloNN.py:42-52, Net/NNet.py:85-87."""
import numpy as np
import pytest

from oracle import net_numpy as nn
from oracle import net_torch
import kat_net


# ---- primitives against hand-derived answers ---------------------------------------------------------------------------
def test_conv_is_cross_correlation_with_hwio_kernel():
    x = np.zeros((1, 5, 5, 2))
    x[0, 1, 2, 1] = 1.0                                  # one asymmetric input pixel in channel 1
    k = np.zeros((3, 3, 2, 3))
    k[0, 2, 1, 2] = 7.0                                  # tap (di=0, dj=2), input channel 1 -> output channel 2
    y = nn.conv2d(x, k, np.array([0.0, 0.0, 0.5]), "same")
    # y[i,j,2] = 0.5 + 7 * x[i+0-1, j+2-1, 1]  ->  the pixel at (1,2) lands on (i,j) = (2,1)
    exp = np.zeros((1, 5, 5, 3)); exp[..., 2] = 0.5; exp[0, 2, 1, 2] = 7.5
    assert np.array_equal(y, exp)
    yv = nn.conv2d(x, k, np.zeros(3), "valid")           # no padding: y[i,j] = 7 * x[i, j+2] -> (1,0), shape 3x3
    expv = np.zeros((1, 3, 3, 3)); expv[0, 1, 0, 2] = 7.0
    assert np.array_equal(yv, expv)


def test_same_padding_is_zero_and_symmetric():
    x = np.ones((1, 4, 4, 1))
    k = np.ones((3, 3, 1, 1))
    y = nn.conv2d(x, k, np.zeros(1), "same")[0, :, :, 0]
    assert y.tolist() == [[4, 6, 6, 4], [6, 9, 9, 6], [6, 9, 9, 6], [4, 6, 6, 4]]


def test_batchnorm_uses_keras_epsilon():
    y = nn.batchnorm(np.array([[2.0]]), [2.0], [0.1], [0.5], [3.0])
    assert y[0, 0] == pytest.approx(2.0 * 1.5 / np.sqrt(3.001) + 0.1, abs=1e-15)
    assert abs(y[0, 0] - (2.0 * 1.5 / np.sqrt(3.00001) + 0.1)) > 1e-4   # torch's default epsilon would be visible


def test_flatten_is_h_w_c_and_dense_is_in_out():
    x = np.arange(2 * 3 * 4, dtype=np.float64).reshape(1, 2, 3, 4)       # value = (h*3 + w)*4 + c
    f = nn.flatten(x)
    assert f[0, (1 * 3 + 2) * 4 + 3] == x[0, 1, 2, 3] and f.shape == (1, 24)
    k = np.zeros((24, 5)); k[7, 3] = 2.0
    assert nn.dense(f, k, np.ones(5))[0].tolist() == [1, 1, 1, 15, 1]


def test_weight_list_is_get_weights_order():
    names = [nm for nm, _ in nn.weight_shapes(8, 512)]
    assert names[:6] == ["conv1.kernel", "conv1.bias", "bn1.gamma", "bn1.beta", "bn1.mean", "bn1.var"]
    assert names[-4:] == ["pi.kernel", "pi.bias", "v.kernel", "v.bias"] and len(names) == 40
    total = sum(int(np.prod(s)) for _, s in nn.weight_shapes(8, 512))
    assert total == 16_051_265 == net_torch.blob_floats(8, 512)          # parameters of the 8x8 / C=512 net
    assert sum(int(np.prod(s)) for _, s in nn.weight_shapes(6, 512)) == 9_745_445 == net_torch.blob_floats(6, 512)


# ---- a whole-network known answer ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [6, 8])
@pytest.mark.parametrize("impl", ["numpy", "torch"])
def test_delta_network_known_answer(n, impl):
    C = 16
    blob, boards, exp = kat_net.delta_network(n, C)
    if impl == "numpy":
        pi, lg, v = nn.forward(blob, boards, n, C)
    else:
        pi, lg, v = net_torch.forward(blob, boards, n, C)
    assert np.abs(lg - exp["logits"]).max() < 1e-5
    assert np.abs(pi - exp["pi"]).max() < 1e-6
    assert np.abs(v - exp["v"]).max() < 1e-6
    # the answer really depends on the orientation: the mirrored position gives a different one
    assert not np.allclose(exp["logits"][0], exp["logits"][1])


# ---- the two restatements agree ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,C,seed", [(8, 16, 0), (6, 24, 1), (8, 32, 2)])
def test_torch_restatement_equals_numpy_restatement(n, C, seed):
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in nn.weight_shapes(n, C):
        kind = name.split(".")[1]
        if kind == "kernel":
            fan = np.prod(shape[:-1])
            w[name] = rng.normal(0, 1.5 / np.sqrt(fan), size=shape)
        elif kind == "gamma":
            w[name] = rng.uniform(0.5, 1.5, size=shape)
        elif kind == "var":
            w[name] = rng.uniform(0.3, 2.0, size=shape)
        else:
            w[name] = rng.normal(0, 0.2, size=shape)
    blob = nn.join(w, n, C)
    boards = (rng.random((9, n, n, 2)) < 0.3).astype(np.float32)
    boards[..., 1] *= 1 - boards[..., 0]
    pi, lg, v, hid = nn.forward(blob, boards, n, C, return_hidden=True)
    tpi, tlg, tv, thid = net_torch.forward(blob, boards, n, C, return_hidden=True)
    for a, b in zip(hid, thid):
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-5 * max(1.0, np.abs(a).max())
    assert np.abs(lg - tlg).max() <= 1e-5 and np.abs(pi - tpi).max() <= 1e-5 and np.abs(v - tv).max() <= 1e-5
    assert np.abs(lg).max() > 0.05                                      # not a degenerate (all-zero) comparison
