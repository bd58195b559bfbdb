"""NETWORK ORACLE — test infrastructure, not the product path.

PyTorch fp32 restatement of the reference's OthelloNN graph (Net/OthelloNN.py:42-52) with Keras
semantics: NHWC input (B,N,N,2); Conv2D 3x3 with bias, kernel HWIO, padding
same,same,valid,valid; BatchNormalization(axis=channels, eps=1e-3) in inference mode (moving
statistics); Flatten in (h,w,c) order; Dense 1024 / 512 with BN + ReLU (dropout inactive in
predict, Net/NNet.py:85); heads Dense(N^2, softmax) and Dense(1, tanh).

PARITY STATUS: TensorFlow/Keras are not installed, so no execution of the reference's own network pins
this restatement ("parity unpinned" against a Keras run).  What pins it instead: an independent NumPy
restatement written from the Keras layer definitions (oracle/net_numpy.py) must agree with this file to
1e-5, and both must reproduce hand-derived known answers for kernel orientation, padding, flatten order,
Dense layout and the BN epsilon (tests/test_net_oracle_pin_cpu.py, tests/kat_net.py).  The 1e-2 check
compares the CUDA tower with this file on identical weights.
The weight blob layout is defined by this file independently of the product (Keras get_weights()
order) so that a layout mistake on either side shows up as a mismatch.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # keras.layers.BatchNormalization default


def _layout(n: int, C: int):
    k1 = (n - 4) * (n - 4) * C
    L = []

    def bn(c):
        L.extend([(c,), (c,), (c,), (c,)])

    L += [(3, 3, 2, C), (C,)]
    bn(C)
    for _ in range(3):
        L += [(3, 3, C, C), (C,)]
        bn(C)
    L += [(k1, 1024), (1024,)]
    bn(1024)
    L += [(1024, 512), (512,)]
    bn(512)
    L += [(512, n * n), (n * n,), (512, 1), (1,)]
    return L


def blob_floats(n: int, C: int) -> int:
    return int(sum(int(np.prod(s)) for s in _layout(n, C)))


def _split(blob, n, C, device):
    t = torch.as_tensor(np.asarray(blob, dtype=np.float32), device=device)
    out, off = [], 0
    for shape in _layout(n, C):
        cnt = int(np.prod(shape))
        out.append(t[off:off + cnt].reshape(shape))
        off += cnt
    assert off == t.numel()
    return out


def boards_from_bits(own, opp, n: int) -> np.ndarray:
    """canonical bitboards -> (B,N,N,2) float32 (channel 0 = side to move)."""
    own = np.asarray(own, dtype=np.uint64).reshape(-1)
    opp = np.asarray(opp, dtype=np.uint64).reshape(-1)
    B = own.size
    x = np.zeros((B, n, n, 2), dtype=np.float32)
    for r in range(n):
        for c in range(n):
            sh = np.uint64(r * 8 + c)
            x[:, r, c, 0] = (own >> sh) & np.uint64(1)
            x[:, r, c, 1] = (opp >> sh) & np.uint64(1)
    return x


@torch.no_grad()
def forward(blob, boards_nhwc, n: int, C: int, device: str = "cpu", return_hidden: bool = False):
    """boards_nhwc: (B,N,N,2) float array.  Returns (pi (B,N*N), logits (B,N*N), v (B,)) as float32 numpy."""
    if device != "cpu":
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    p = _split(blob, n, C, device)
    it = iter(p)
    x = torch.as_tensor(np.asarray(boards_nhwc, dtype=np.float32), device=device).permute(0, 3, 1, 2).contiguous()
    hidden = []

    def bn_relu(y, ch_dim):
        g, b, m, v = next(it), next(it), next(it), next(it)
        shape = [1] * y.dim()
        shape[ch_dim] = -1
        y = (y - m.view(shape)) / torch.sqrt(v.view(shape) + BN_EPS) * g.view(shape) + b.view(shape)
        return torch.relu(y)

    for pad in ("same", "same", "valid", "valid"):
        k, b = next(it), next(it)                       # HWIO
        w = k.permute(3, 2, 0, 1).contiguous()          # OIHW
        x = F.conv2d(x, w, b, padding=1 if pad == "same" else 0)
        x = bn_relu(x, 1)
        hidden.append(x.permute(0, 2, 3, 1).reshape(x.shape[0], -1, x.shape[1]).cpu().numpy())
    x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)   # Flatten of NHWC: (h, w, c)
    for _ in range(2):
        k, b = next(it), next(it)
        x = x @ k + b
        x = bn_relu(x, 1)
        hidden.append(x.cpu().numpy()[:, None, :])
    kp, bp, kv, bv = next(it), next(it), next(it), next(it)
    logits = x @ kp + bp
    pi = torch.softmax(logits, dim=1)
    v = torch.tanh(x @ kv + bv).reshape(-1)
    out = (pi.cpu().numpy(), logits.cpu().numpy(), v.cpu().numpy())
    return out + (hidden,) if return_hidden else out


class TorchNet:
    """Reference-style net object for the CPU baseline: .network_type / .predict(board) with batch 1
    (Net/NNet.py:70-87)."""

    def __init__(self, blob, n: int, C: int):
        self.n, self.C = n, C
        self.blob = np.asarray(blob, dtype=np.float32)
        self.network_type = "ONN"
        self.calls = 0

    def predict(self, board):
        self.calls += 1
        x = np.asarray(board, dtype=np.float32)[None]
        pi, _, v = forward(self.blob, x, self.n, self.C)
        return pi[0].reshape(self.n, self.n), np.float32(v[0])
