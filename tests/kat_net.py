"""A whole-network KNOWN ANSWER for the OthelloNN graph (Net/OthelloNN.py:42-52), derived by index arithmetic only.

Every conv kernel is a single 1 (a "delta"), so each conv layer just SHIFTS channel 0 of its input:
    conv1 'same',  tap (0,0):  y[i,j] = x[i-1, j-1]     (zero outside)
    conv2 'same',  tap (2,2):  y[i,j] = x[i+1, j+1]
    conv3 'valid', tap (0,1):  y[i,j] = x[i,   j+1]      (output (n-2)x(n-2))
    conv4 'valid', tap (1,0):  y[i,j] = x[i+1, j]        (output (n-4)x(n-4))
so a single own disc at (r, c) arrives at (r-1, c-1) of the final map, provided the intermediate squares are on the
board.  Flatten is (h,w,c): flat index ((r-1)*(n-4) + (c-1))*C.  fc1 reads that index with weight 2 into unit 5 and
goes through a NON-identity BatchNormalization (gamma 2, beta 0.1, mean 0.5, var 3, epsilon 1e-3); fc2 halves unit 5
into unit 7; the policy head maps unit 7 to action `a` with weight 3 (+ a bias ramp), the value head with weight 0.25.
A flipped or transposed kernel, a (c,h,w) flatten, a (out,in) Dense kernel or epsilon 1e-5 all change the answer."""
from __future__ import annotations

import numpy as np

from oracle import net_numpy as nn


def delta_network(n: int, C: int, bn5=(2.0, 0.1, 0.5, 3.0)):
    """-> (blob float32, boards (2,n,n,2) float32, expected dict(logits (2,n*n), pi, v)).
    bn5 = (gamma, beta, mean, var) of fc1's BatchNormalization unit 5; the bf16 device test passes a variance whose
    folded scale is exactly representable (var + epsilon = 4)."""
    w = nn.zero_weights(n, C)
    w["conv1.kernel"][0, 0, 0, 0] = 1.0
    w["conv2.kernel"][2, 2, 0, 0] = 1.0
    w["conv3.kernel"][0, 1, 0, 0] = 1.0
    w["conv4.kernel"][1, 0, 0, 0] = 1.0
    r, c = (2, 3) if n >= 8 else (2, 1)          # the probe disc (own colour); asymmetric on purpose, lands inside the final map
    m = n - 4
    flat = ((r - 1) * m + (c - 1)) * C            # (h, w, c) flatten of the disc's final position, channel 0
    w["fc1.kernel"][flat, 5] = 2.0
    g5, b5, m5, v5 = bn5
    w["bn5.gamma"][5], w["bn5.beta"][5], w["bn5.mean"][5], w["bn5.var"][5] = g5, b5, m5, v5
    w["fc2.kernel"][5, 7] = 0.5
    a = 1 * n + 4                                 # action (1, 4)
    w["pi.kernel"][7, a] = 3.0
    w["pi.bias"][:] = np.arange(n * n) * 0.01
    w["v.kernel"][7, 0] = 0.25
    w["v.bias"][0] = -0.05
    blob = nn.join(w, n, C)
    boards = np.zeros((2, n, n, 2), dtype=np.float32)
    boards[0, r, c, 0] = 1.0                      # board 0: the disc where the chain of shifts expects it
    boards[1, c, r, 0] = 1.0                      # board 1: the transposed position -> misses the fc1 tap
    boards[:, 0, 0, 1] = 1.0                      # an opponent disc: channel 1 is wired to nothing
    f1 = max(0.0, g5 * (2.0 * 1.0 - m5) / np.sqrt(v5 + 1e-3) + b5)               # board 0, unit 5
    f1_miss = max(0.0, g5 * (0.0 - m5) / np.sqrt(v5 + 1e-3) + b5)                # board 1: relu(negative) = 0
    logits = np.tile(np.arange(n * n) * 0.01, (2, 1))
    vpre = np.full(2, -0.05)
    for b, f in enumerate((f1, f1_miss)):
        f2 = 0.5 * f
        logits[b, a] += 3.0 * f2
        vpre[b] += 0.25 * f2
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    return blob, boards, dict(logits=logits, pi=e / e.sum(axis=1, keepdims=True), v=np.tanh(vpre))
