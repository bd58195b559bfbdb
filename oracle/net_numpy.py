"""NETWORK ORACLE, SECOND STATEMENT — test infrastructure, not the product path.

An independent NumPy (float64) restatement of the reference's OthelloNN graph (Net/OthelloNN.py:42-52), written from
the Keras 2.4 / TF 2.3 LAYER DEFINITIONS (requirements.txt:16,35) rather than from oracle/net_torch.py, and using no
torch operator.  Its only purpose is to pin net_torch.py (and through it the CUDA tower): the two restatements must
agree to 1e-5 on random asymmetric weights, and the primitives below are checked against hand-derived known answers
that fail on a transposed / flipped kernel, a wrong 'same' padding, a (c,h,w) flatten or torch's BN epsilon
(tests/test_net_oracle_pin_cpu.py).

TensorFlow/Keras are not installed in this image, so no Keras EXECUTION pins these numbers; the layer semantics
restated here are:
  Conv2D(filters, 3, padding)       keras/layers/convolutional.py -> tf.nn.conv2d, data_format channels_last (NHWC),
                                    kernel (kh, kw, in, out) = HWIO, CROSS-CORRELATION (no kernel flip):
                                      y[b,i,j,o] = bias[o] + sum_{di,dj,c} x[b, i+di-pt, j+dj-pl, c] * K[di,dj,c,o]
                                    'same' with stride 1, kernel 3: pt = pl = 1 zero rows/cols on each side; 'valid': none.
  BatchNormalization(axis)          inference: gamma * (x - moving_mean) / sqrt(moving_variance + epsilon) + beta,
                                    epsilon = 1e-3 (Keras default, NOT torch's 1e-5); weights order gamma, beta, mean, var.
  Flatten()                         channels_last: row-major reshape of (h, w, c).
  Dense(units)                      x @ kernel + bias, kernel (in, out).
  Activation('relu'), softmax (last axis), tanh; Dropout is the identity in predict (Net/NNet.py:85).
"""
from __future__ import annotations

import numpy as np

BN_EPSILON = 1e-3


def conv2d(x, kernel, bias, padding: str):
    """x (B,H,W,Cin), kernel (3,3,Cin,Cout) HWIO, bias (Cout,) -> (B,H',W',Cout)."""
    x = np.asarray(x, dtype=np.float64)
    kernel = np.asarray(kernel, dtype=np.float64)
    kh, kw = kernel.shape[0], kernel.shape[1]
    if padding == "same":
        pt, pl = (kh - 1) // 2, (kw - 1) // 2
        x = np.pad(x, ((0, 0), (pt, kh - 1 - pt), (pl, kw - 1 - pl), (0, 0)))
    elif padding != "valid":
        raise ValueError(padding)
    B, H, W, _ = x.shape
    oh, ow = H - kh + 1, W - kw + 1
    y = np.zeros((B, oh, ow, kernel.shape[3]))
    for di in range(kh):
        for dj in range(kw):
            # window element (di, dj) of every output position, contracted over input channels
            y += np.tensordot(x[:, di:di + oh, dj:dj + ow, :], kernel[di, dj], axes=([3], [0]))
    return y + np.asarray(bias, dtype=np.float64)


def batchnorm(x, gamma, beta, mean, var):
    """Last-axis (channel / feature) normalisation with the moving statistics."""
    g, b, m, v = (np.asarray(a, dtype=np.float64) for a in (gamma, beta, mean, var))
    return g * (np.asarray(x, dtype=np.float64) - m) / np.sqrt(v + BN_EPSILON) + b


def relu(x):
    return np.maximum(x, 0.0)


def flatten(x):
    return np.asarray(x).reshape(x.shape[0], -1)  # channels_last: (h, w, c) row-major


def dense(x, kernel, bias):
    return np.asarray(x, dtype=np.float64) @ np.asarray(kernel, dtype=np.float64) + np.asarray(bias, dtype=np.float64)


def softmax(x):
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


# ---- the weight list: model.get_weights() order of Net/OthelloNN.py:42-52 ----------------------------------------------
def weight_shapes(n: int, C: int):
    """[(name, shape)]: Keras creates the layers in call order - conv2d, batch_normalization, conv2d_1, ... dense,
    batch_normalization_4, dense_1, batch_normalization_5, pi, v - and get_weights() concatenates their weights."""
    out = []
    cin = 2
    for i in range(1, 5):
        out += [(f"conv{i}.kernel", (3, 3, cin, C)), (f"conv{i}.bias", (C,))]
        out += [(f"bn{i}.{k}", (C,)) for k in ("gamma", "beta", "mean", "var")]
        cin = C
    flat = (n - 4) * (n - 4) * C
    for name, fin, fout, bn in (("fc1", flat, 1024, "bn5"), ("fc2", 1024, 512, "bn6")):
        out += [(f"{name}.kernel", (fin, fout)), (f"{name}.bias", (fout,))]
        out += [(f"{bn}.{k}", (fout,)) for k in ("gamma", "beta", "mean", "var")]
    out += [("pi.kernel", (512, n * n)), ("pi.bias", (n * n,)), ("v.kernel", (512, 1)), ("v.bias", (1,))]
    return out


def split(blob, n: int, C: int) -> dict:
    blob = np.asarray(blob)
    w, off = {}, 0
    for name, shape in weight_shapes(n, C):
        cnt = int(np.prod(shape))
        w[name] = blob[off:off + cnt].reshape(shape)
        off += cnt
    if off != blob.size:
        raise ValueError(f"blob has {blob.size} values, the graph has {off}")
    return w


def join(w: dict, n: int, C: int) -> np.ndarray:
    return np.concatenate([np.asarray(w[name], dtype=np.float32).reshape(-1) for name, _ in weight_shapes(n, C)])


def zero_weights(n: int, C: int) -> dict:
    """All kernels / biases zero, every BatchNormalization the exact identity (gamma 1, beta 0, mean 0, and
    var = 1 - epsilon so that sqrt(var + epsilon) = 1)."""
    w = {}
    for name, shape in weight_shapes(n, C):
        kind = name.split(".")[1]
        w[name] = np.full(shape, {"gamma": 1.0, "var": 1.0 - BN_EPSILON}.get(kind, 0.0), dtype=np.float64)
    return w


def forward(blob, boards_nhwc, n: int, C: int, return_hidden: bool = False):
    """boards (B,N,N,2) -> (pi (B,N*N), logits (B,N*N), v (B,)) float64 (+ the six hidden activations)."""
    w = split(np.asarray(blob, dtype=np.float64), n, C)
    x = np.asarray(boards_nhwc, dtype=np.float64)
    hidden = []
    for i, padding in enumerate(("same", "same", "valid", "valid"), start=1):
        x = conv2d(x, w[f"conv{i}.kernel"], w[f"conv{i}.bias"], padding)
        x = relu(batchnorm(x, w[f"bn{i}.gamma"], w[f"bn{i}.beta"], w[f"bn{i}.mean"], w[f"bn{i}.var"]))
        hidden.append(x.reshape(x.shape[0], -1, x.shape[3]))
    x = flatten(x)
    for name, bn in (("fc1", "bn5"), ("fc2", "bn6")):
        x = dense(x, w[f"{name}.kernel"], w[f"{name}.bias"])
        x = relu(batchnorm(x, w[f"{bn}.gamma"], w[f"{bn}.beta"], w[f"{bn}.mean"], w[f"{bn}.var"]))
        hidden.append(x[:, None, :])
    logits = dense(x, w["pi.kernel"], w["pi.bias"])
    v = np.tanh(dense(x, w["v.kernel"], w["v.bias"])).reshape(-1)
    out = (softmax(logits), logits, v)
    return out + (hidden,) if return_hidden else out
