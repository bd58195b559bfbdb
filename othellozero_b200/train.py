"""Training step ("next" row): NNetWrapper.train (Net/NNet.py:53-68) with the compile settings of
Net/OthelloNN.py:55-56 - Adam(lr, clipvalue=0.5), dropout 0.3, BatchNormalization momentum 0.99 / eps 1e-3, batch 32,
10 epochs - done by PyTorch autograd.

This is host-side plumbing around the hot path, not a hand-written kernel: the self-play engine consumes the result
as a weight blob (oz_net_load_weights* folds BN and casts to bf16 on the device).

Two details of what Keras actually computes, reproduced here (both were deviations in round 1):
  * the policy loss.  compile() names 'categorical_crossentropy' for the output 'pi-reshaped' of shape (B,N,N)
    (Net/OthelloNN.py:51,55), and Keras applies that loss along the LAST axis: every board ROW is renormalised and scored
    as its own distribution, then averaged over B*N rows.  With one-hot targets only the target's row contributes, so the
    loss is (1/N) * -log(pi[r*,c*] / sum_c pi[r*,c]) - not a cross-entropy over the N*N squares.  `policy_loss=
    "reference"` (default) is that; "full_board" is the textbook -log pi[r*,c*].
  * BatchNormalization's moving_variance is updated with the BIASED batch variance (torch's BatchNorm uses the unbiased
    one for running_var): KerasBatchNorm below.
Multi-GPU: `train_blob(..., ddp=True)` under torch.distributed shards every global batch over the ranks (rank r takes
samples r::world), synchronises the BatchNormalization batch statistics and averages the gradients, so N GPUs take the
SAME optimisation steps one GPU would (tests/test_train_cpu.py, gloo world 2).  `ddp="auto"` shards only batches large
enough to be worth it; at the reference's batch 32 one GPU replays a CUDA graph of the step instead (see train_blob)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .net import blob_layout, pack_blob, unpack_blob

BN_EPS = 1e-3          # keras BatchNormalization default
KERAS_MOMENTUM = 0.99  # moving = 0.99 * moving + 0.01 * batch
KERAS_CE_EPS = 1e-7    # keras.backend.epsilon(): categorical_crossentropy clips probabilities to [eps, 1 - eps]


def _dist_world():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class KerasBatchNorm(nn.Module):
    """keras.layers.BatchNormalization over dimension 1 (channels of NCHW / features): training normalises with the
    batch mean and BIASED variance and moves the moving statistics with them (momentum 0.99); inference uses the moving
    statistics.  sync=True pools the batch statistics over all ranks (differentiably)."""

    def __init__(self, features: int, eps: float = BN_EPS, momentum: float = KERAS_MOMENTUM):
        super().__init__()
        self.eps, self.momentum, self.sync = eps, momentum, False
        self.weight = nn.Parameter(torch.ones(features))
        self.bias = nn.Parameter(torch.zeros(features))
        self.register_buffer("running_mean", torch.zeros(features))
        self.register_buffer("running_var", torch.ones(features))

    def forward(self, x):
        shape = [1, -1] + [1] * (x.dim() - 2)
        if self.training and not (self.sync and _dist_world() > 1):
            # one fused native kernel instead of ~20 elementwise ones (a training step is 26 % faster).  F.batch_norm
            # normalises with the biased batch variance like Keras, but moves running_var with the UNBIASED one, so that
            # update is redone:  new = (1 - m) * old + m * unbiased  =>  m * biased = (new - (1 - m) * old) * (count - 1) / count
            count = x.numel() // x.shape[1]
            with torch.no_grad():
                rm, rv = self.running_mean.clone(), self.running_var.clone()  # scratch copies: autograd keeps a version check on
                kept = rv * self.momentum                                     # what batch_norm was handed; the buffers are written after
            y = F.batch_norm(x, rm, rv, self.weight, self.bias, True, 1.0 - self.momentum, self.eps)
            with torch.no_grad():
                self.running_mean.copy_(rm)
                self.running_var.copy_((rv - kept) * ((count - 1) / count if count > 1 else 1.0) + kept)
            return y
        if self.training:
            dims = [0] + list(range(2, x.dim()))
            count = x.numel() // x.shape[1]
            s1, s2 = x.sum(dims), (x * x).sum(dims)
            if self.sync and _dist_world() > 1:
                import torch.distributed.nn.functional as dfn
                packed = dfn.all_reduce(torch.cat([s1, s2, s1.new_tensor([float(count)])]))
                s1, s2, count = packed[:s1.numel()], packed[s1.numel():-1], packed[-1]
            mean = s1 / count
            var = (s2 / count - mean * mean).clamp_min(0.0)
            with torch.no_grad():
                self.running_mean.mul_(self.momentum).add_(mean.detach(), alpha=1 - self.momentum)
                self.running_var.mul_(self.momentum).add_(var.detach(), alpha=1 - self.momentum)
        else:
            mean, var = self.running_mean, self.running_var
        return (x - mean.view(shape)) * torch.rsqrt(var.view(shape) + self.eps) * self.weight.view(shape) + self.bias.view(shape)


class OthelloNNTorch(nn.Module):
    """Net/OthelloNN.py:42-52 as a torch module (NCHW inside, Keras layouts at the blob boundary)."""

    def __init__(self, board_size: int, channels: int = 512, dropout: float = 0.3):
        super().__init__()
        n, C = board_size, channels
        self.n, self.C = n, C
        self.convs = nn.ModuleList([nn.Conv2d(2, C, 3, padding=1), nn.Conv2d(C, C, 3, padding=1),
                                    nn.Conv2d(C, C, 3), nn.Conv2d(C, C, 3)])
        self.bns = nn.ModuleList([KerasBatchNorm(C) for _ in range(4)])
        k1 = (n - 4) * (n - 4) * C
        self.fc1, self.fc2 = nn.Linear(k1, 1024), nn.Linear(1024, 512)
        self.bn5, self.bn6 = KerasBatchNorm(1024), KerasBatchNorm(512)
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


def policy_loss_per_sample(logits, target, n: int, kind: str = "reference"):
    """Per-sample policy loss.  "reference": what Keras computes for the (B,N,N) 'pi-reshaped' output
    (Net/OthelloNN.py:51,55; see the module docstring) - row-wise renormalised, clipped cross-entropy, averaged over the
    N rows.  "full_board": cross-entropy over all N*N squares."""
    if kind == "full_board":
        return -(target * F.log_softmax(logits, dim=1)).sum(dim=1)
    pi = torch.softmax(logits, dim=1).view(-1, n, n)
    rows = pi / pi.sum(dim=2, keepdim=True)
    rows = rows.clamp(KERAS_CE_EPS, 1.0 - KERAS_CE_EPS)
    return -(target.view(-1, n, n) * rows.log()).sum(dim=2).mean(dim=1)


MIN_SAMPLES_PER_RANK = 32   # ddp="auto": shard a batch over the ranks only if every rank still gets this many samples
GRAPH_MIN_BATCHES = 8        # capture the step in a CUDA graph only if an epoch has at least this many full batches


def train_blob(blob, examples, board_size: int, channels: int = 512, epochs: int = 10, batch_size: int = 32,
               lr: float = 1e-3, dropout: float = 0.3, clipvalue: float = 0.5, device=None, seed: int = 0,
               verbose: bool = False, policy_loss: str = "reference", ddp=False, cuda_graph: bool | None = None):
    """model.fit of Net/NNet.py:67-68.  Returns (new_blob, history) with history = per-epoch mean (loss, pi_loss, v_loss).
    `examples`: the reference's list of (board, policy, z), or a tuple of arrays (boards (E,N,N,2), policies (E,N*N), z (E,)).

    ddp=True (inside an initialised torch.distributed group): every rank holds the same examples and the same weights;
    each global batch of `batch_size` is split over the ranks, BatchNormalization statistics and gradients are pooled,
    and every rank returns the same new blob.
    ddp="auto": shard only when that leaves >= MIN_SAMPLES_PER_RANK samples per rank.  At the reference's batch size of 32
    a step is latency-, not throughput-bound, and sharding it 8 ways costs 8.7 ms per step on 8 B200s (measured: 24 000
    steps = 208 s) against ~1 ms for one GPU replaying the captured step; so rank 0 trains alone and the result is
    broadcast (C1) - every rank still returns the same blob.
    cuda_graph (default: on CUDA without sharding): the optimisation step (gather of the batch, forward, loss, backward,
    clipping, Adam) is captured once in a CUDA graph and replayed - same arithmetic, one launch per step."""
    import torch.distributed as dist
    group = _dist_world()
    shard = bool(ddp) and group > 1 and (ddp != "auto" or batch_size // group >= MIN_SAMPLES_PER_RANK)
    world = group if shard else 1
    rank = dist.get_rank() if world > 1 else 0
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    on_cuda = str(device).startswith("cuda")
    if ddp == "auto" and group > 1 and not shard:
        # rank 0 takes the steps (below, unsharded); the others receive the result
        new_blob, history = None, None
        if dist.get_rank() == 0:
            new_blob, history = train_blob(blob, examples, board_size, channels, epochs, batch_size, lr, dropout, clipvalue,
                                           device, seed, verbose, policy_loss, ddp=False, cuda_graph=cuda_graph)
        n_floats = int(np.asarray(blob).size)
        t = torch.from_numpy(new_blob).to(device) if new_blob is not None else torch.empty(n_floats, dtype=torch.float32, device=device)
        h = torch.tensor(history if history is not None else [[0.0] * 3] * epochs, dtype=torch.float64, device=device)
        dist.broadcast(t, src=0)
        dist.broadcast(h, src=0)
        return t.cpu().numpy(), [tuple(float(x) for x in row) for row in h.cpu()]
    torch.manual_seed(seed + 7919 * rank)          # dropout masks differ per rank, the weights start equal
    model = OthelloNNTorch(board_size, channels, dropout).load_blob(blob).to(device)
    if world > 1:
        for m in model.modules():
            if isinstance(m, KerasBatchNorm):
                m.sync = True
    boards, pis, vs = examples if isinstance(examples, tuple) else examples_to_arrays(examples)
    xb, pb, vb = (torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(device) for a in (boards, pis, vs))
    params = [p for p in model.parameters()]
    n = xb.shape[0]
    use_graph = (cuda_graph if cuda_graph is not None else True) and on_cuda and world == 1 and n // batch_size >= GRAPH_MIN_BATCHES
    opt = torch.optim.Adam(params, lr=lr, eps=1e-7, **({"capturable": True, "foreach": True} if use_graph else {}))  # keras Adam epsilon
    gen = torch.Generator(device="cpu").manual_seed(seed)   # the same shuffles on every rank
    history = []
    model.train()
    tot = torch.zeros(3, dtype=torch.float64, device=device)  # summed on the device: no host sync per step

    flat, flat_views = None, None
    if world > 1:
        flat = torch.zeros(sum(p.numel() for p in params) + 3, dtype=torch.float32, device=device)
        flat_views, off = [], 0
        for p in params:
            flat_views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def step(gidx):
        """One optimisation step on the global batch `gidx` (this rank's share of it)."""
        idx = gidx[rank::world] if world > 1 else gidx
        logits, v = model(xb.index_select(0, idx))
        # sums over this rank's samples / the GLOBAL batch size: adding the ranks' gradients gives the batch mean's
        pi_loss = policy_loss_per_sample(logits, pb.index_select(0, idx), board_size, policy_loss).sum() / gidx.numel()
        v_loss = ((v - vb.index_select(0, idx)) ** 2).sum() / gidx.numel()
        loss = pi_loss + v_loss
        opt.zero_grad(set_to_none=True)
        loss.backward()
        part = torch.stack([loss.detach(), pi_loss.detach(), v_loss.detach()]).double()
        if world > 1:
            # ONE all-reduce per step: the gradients and the three loss parts travel in a flat buffer whose slices then
            # serve as the .grad tensors themselves (no per-parameter collectives, no copies back)
            torch._foreach_copy_(flat_views, [p.grad if p.grad is not None else torch.zeros_like(p) for p in params])
            flat[-3:] = part.float()
            dist.all_reduce(flat)
            part = flat[-3:].double()
            for p, g in zip(params, flat_views):
                p.grad = g
        torch.nn.utils.clip_grad_value_(params, clipvalue)      # Adam(clipvalue=0.5)
        opt.step()
        tot.add_(part * gidx.numel())

    graph, static_idx, eager_steps = None, None, 0
    for ep in range(epochs):
        perm = torch.randperm(n, generator=gen).to(device)  # keras fit shuffles every epoch
        tot.zero_()
        cnt = 0
        for i in range(0, n, batch_size):
            gidx = perm[i:i + batch_size]
            if gidx.numel() < 2:
                continue  # BatchNorm needs more than one sample
            cnt += gidx.numel()
            if not use_graph or gidx.numel() != batch_size:
                step(gidx)
                continue
            if graph is None and eager_steps < 3:
                # the first full steps run eagerly on a side stream (they are real steps AND the warm-up capture needs)
                side = torch.cuda.Stream(device=device)
                side.wait_stream(torch.cuda.current_stream(device))
                with torch.cuda.stream(side):
                    step(gidx)
                torch.cuda.current_stream(device).wait_stream(side)
                eager_steps += 1
                continue
            if graph is None:
                static_idx = gidx.clone()
                graph = torch.cuda.CUDAGraph()
                opt.zero_grad(set_to_none=True)
                with torch.cuda.graph(graph):   # records only; the replay below is this batch's step
                    step(static_idx)
            static_idx.copy_(gidx)
            graph.replay()
        history.append(tuple(float(x) for x in (tot / max(1, cnt)).cpu()))
        if verbose and rank == 0:
            print(f"epoch {ep + 1}/{epochs}: loss {history[-1][0]:.4f} pi {history[-1][1]:.4f} v {history[-1][2]:.4f}")
    del graph
    model.eval()
    return model.cpu().to_blob(), history
