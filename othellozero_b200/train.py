"""Training step ("next" row): NNetWrapper.train (Net/NNet.py:53-68) with the compile settings of
Net/OthelloNN.py:55-56 — categorical cross-entropy on the policy + MSE on the value, Adam(lr, clipvalue=0.5),
dropout 0.3, BatchNormalization momentum 0.99 / eps 1e-3, batch 32, 10 epochs — done by PyTorch autograd.

This is host-side plumbing around the hot path, not a hand-written kernel: the self-play engine consumes the result
as a weight blob (oz_net_load_weights* folds BN and casts to bf16 on the device).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .net import blob_layout, pack_blob, unpack_blob

BN_EPS = 1e-3       # keras BatchNormalization default
BN_MOMENTUM = 0.01  # torch convention = 1 - keras momentum (0.99)


class OthelloNNTorch(nn.Module):
    """Net/OthelloNN.py:42-52 as a torch module (NCHW inside, Keras layouts at the blob boundary)."""

    def __init__(self, board_size: int, channels: int = 512, dropout: float = 0.3):
        super().__init__()
        n, C = board_size, channels
        self.n, self.C = n, C
        self.convs = nn.ModuleList([nn.Conv2d(2, C, 3, padding=1), nn.Conv2d(C, C, 3, padding=1),
                                    nn.Conv2d(C, C, 3), nn.Conv2d(C, C, 3)])
        self.bns = nn.ModuleList([nn.BatchNorm2d(C, eps=BN_EPS, momentum=BN_MOMENTUM) for _ in range(4)])
        k1 = (n - 4) * (n - 4) * C
        self.fc1, self.fc2 = nn.Linear(k1, 1024), nn.Linear(1024, 512)
        self.bn5 = nn.BatchNorm1d(1024, eps=BN_EPS, momentum=BN_MOMENTUM)
        self.bn6 = nn.BatchNorm1d(512, eps=BN_EPS, momentum=BN_MOMENTUM)
        self.pi, self.v = nn.Linear(512, n * n), nn.Linear(512, 1)
        self.dropout = dropout

    def forward(self, boards_nhwc):
        x = boards_nhwc.permute(0, 3, 1, 2)
        for conv, bn in zip(self.convs, self.bns):
            x = F.relu(bn(conv(x)))
        x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)  # keras Flatten of NHWC: (h, w, c)
        x = F.dropout(F.relu(self.bn5(self.fc1(x))), self.dropout, self.training)
        x = F.dropout(F.relu(self.bn6(self.fc2(x))), self.dropout, self.training)
        return self.pi(x), torch.tanh(self.v(x)).reshape(-1)

    # ---- Keras-order blob <-> parameters ---------------------------------------------------------------------
    def load_blob(self, blob: np.ndarray):
        w = unpack_blob(np.asarray(blob, dtype=np.float32), self.n, self.C)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        with torch.no_grad():
            for i, (conv, bn) in enumerate(zip(self.convs, self.bns), start=1):
                conv.weight.copy_(t(w[f"conv{i}.kernel"]).permute(3, 2, 0, 1))  # HWIO -> OIHW
                conv.bias.copy_(t(w[f"conv{i}.bias"]))
                self._load_bn(bn, w, f"bn{i}")
            for name, fc in (("fc1", self.fc1), ("fc2", self.fc2), ("pi", self.pi), ("v", self.v)):
                fc.weight.copy_(t(w[f"{name}.kernel"]).t())                     # (in,out) -> (out,in)
                fc.bias.copy_(t(w[f"{name}.bias"]))
            self._load_bn(self.bn5, w, "bn5")
            self._load_bn(self.bn6, w, "bn6")
        return self

    @staticmethod
    def _load_bn(bn, w, prefix):
        bn.weight.copy_(torch.from_numpy(w[f"{prefix}.gamma"].copy()))
        bn.bias.copy_(torch.from_numpy(w[f"{prefix}.beta"].copy()))
        bn.running_mean.copy_(torch.from_numpy(w[f"{prefix}.mean"].copy()))
        bn.running_var.copy_(torch.from_numpy(w[f"{prefix}.var"].copy()))

    def to_blob(self) -> np.ndarray:
        w = {}
        g = lambda p: p.detach().cpu().numpy()
        for i, (conv, bn) in enumerate(zip(self.convs, self.bns), start=1):
            w[f"conv{i}.kernel"] = g(conv.weight.permute(2, 3, 1, 0))
            w[f"conv{i}.bias"] = g(conv.bias)
            self._save_bn(bn, w, f"bn{i}")
        for name, fc in (("fc1", self.fc1), ("fc2", self.fc2), ("pi", self.pi), ("v", self.v)):
            w[f"{name}.kernel"] = g(fc.weight.t())
            w[f"{name}.bias"] = g(fc.bias)
        self._save_bn(self.bn5, w, "bn5")
        self._save_bn(self.bn6, w, "bn6")
        assert set(w) == {nm for nm, _ in blob_layout(self.n, self.C)}
        return pack_blob(w, self.n, self.C)

    @staticmethod
    def _save_bn(bn, w, prefix):
        w[f"{prefix}.gamma"] = bn.weight.detach().cpu().numpy()
        w[f"{prefix}.beta"] = bn.bias.detach().cpu().numpy()
        w[f"{prefix}.mean"] = bn.running_mean.detach().cpu().numpy()
        w[f"{prefix}.var"] = bn.running_var.detach().cpu().numpy()


def examples_to_arrays(examples):
    """[(board (N,N,2), policy (N,N), z)] -> float32 arrays (Net/NNet.py:59-62)."""
    boards = np.stack([np.asarray(b, dtype=np.float32) for b, _, _ in examples])
    pis = np.stack([np.asarray(p, dtype=np.float32).reshape(-1) for _, p, _ in examples])
    vs = np.array([float(z) for _, _, z in examples], dtype=np.float32)
    return boards, pis, vs


def train_blob(blob, examples, board_size: int, channels: int = 512, epochs: int = 10, batch_size: int = 32,
               lr: float = 1e-3, dropout: float = 0.3, clipvalue: float = 0.5, device=None, seed: int = 0,
               verbose: bool = False):
    """model.fit of Net/NNet.py:67-68.  Returns (new_blob, history) with history = per-epoch mean (loss, pi_loss, v_loss)."""
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    torch.manual_seed(seed)
    model = OthelloNNTorch(board_size, channels, dropout).load_blob(blob).to(device)
    boards, pis, vs = examples_to_arrays(examples)
    xb, pb, vb = (torch.from_numpy(a).to(device) for a in (boards, pis, vs))
    opt = torch.optim.Adam(model.parameters(), lr=lr, eps=1e-7)  # keras Adam epsilon
    n = xb.shape[0]
    gen = torch.Generator(device="cpu").manual_seed(seed)
    history = []
    model.train()
    for ep in range(epochs):
        perm = torch.randperm(n, generator=gen).to(device)  # keras fit shuffles every epoch
        tot = torch.zeros(3, dtype=torch.float64, device=device)  # summed on the device: no host sync per step
        cnt = 0
        for i in range(0, n, batch_size):
            idx = perm[i:i + batch_size]
            if idx.numel() < 2:
                continue  # BatchNorm needs more than one sample
            logits, v = model(xb[idx])
            pi_loss = -(pb[idx] * F.log_softmax(logits, dim=1)).sum(dim=1).mean()  # categorical_crossentropy
            v_loss = F.mse_loss(v, vb[idx])
            loss = pi_loss + v_loss
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_value_(model.parameters(), clipvalue)      # Adam(clipvalue=0.5)
            opt.step()
            tot += torch.stack([loss.detach(), pi_loss.detach(), v_loss.detach()]).double() * idx.numel()
            cnt += idx.numel()
        history.append(tuple(float(x) for x in (tot / max(1, cnt)).cpu()))
        if verbose:
            print(f"epoch {ep + 1}/{epochs}: loss {history[-1][0]:.4f} pi {history[-1][1]:.4f} v {history[-1][2]:.4f}")
    model.eval()
    return model.cpu().to_blob(), history
