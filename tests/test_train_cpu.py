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


def test_policy_loss_is_keras_rowwise_crossentropy():
    """Net/OthelloNN.py:51,55: 'categorical_crossentropy' on the (B,N,N) output = per ROW renormalised CE, mean over B*N."""
    n, B = 6, 5
    rng = np.random.default_rng(1)
    logits = rng.normal(size=(B, n * n)).astype(np.float32)
    target = np.zeros((B, n, n), dtype=np.float32)
    tr, tc = rng.integers(n, size=B), rng.integers(n, size=B)
    target[np.arange(B), tr, tc] = 1
    # keras.backend.categorical_crossentropy(target, output) restated in numpy on the reshaped output
    pi = np.exp(logits - logits.max(1, keepdims=True)); pi /= pi.sum(1, keepdims=True)
    out = pi.reshape(B, n, n)
    out = out / out.sum(axis=-1, keepdims=True)
    out = np.clip(out, 1e-7, 1 - 1e-7)
    keras_loss = (-(target * np.log(out)).sum(axis=-1)).mean()        # mean over the B*N rows
    got = train.policy_loss_per_sample(torch.from_numpy(logits), torch.from_numpy(target.reshape(B, -1)), n).mean()
    assert abs(float(got) - float(keras_loss)) < 1e-6
    by_hand = np.mean([-np.log(pi.reshape(B, n, n)[b, tr[b], tc[b]] / pi.reshape(B, n, n)[b, tr[b]].sum()) / n for b in range(B)])
    assert abs(float(got) - by_hand) < 1e-6
    full = train.policy_loss_per_sample(torch.from_numpy(logits), torch.from_numpy(target.reshape(B, -1)), n, "full_board").mean()
    assert abs(float(full) - np.mean([-np.log(pi.reshape(B, n, n)[b, tr[b], tc[b]]) for b in range(B)])) < 1e-6


def test_keras_batchnorm_moves_with_the_biased_variance():
    bn = train.KerasBatchNorm(3)
    x = torch.tensor(np.random.default_rng(0).normal(2.0, 3.0, size=(7, 3, 2, 2)).astype(np.float32))
    bn.train()
    y = bn(x)
    m = x.mean(dim=(0, 2, 3)); v = x.var(dim=(0, 2, 3), unbiased=False)
    assert torch.allclose(bn.running_mean, 0.01 * m, atol=1e-6)
    assert torch.allclose(bn.running_var, 0.99 + 0.01 * v, atol=1e-5)          # biased (Keras), not x.var(unbiased=True)
    assert torch.allclose(y, (x - m.view(1, 3, 1, 1)) / torch.sqrt(v.view(1, 3, 1, 1) + 1e-3), atol=1e-5)
    bn.eval()
    assert torch.allclose(bn(x), (x - bn.running_mean.view(1, 3, 1, 1)) / torch.sqrt(bn.running_var.view(1, 3, 1, 1) + 1e-3), atol=1e-6)


def _ddp_worker(rank, world, port, blob, arrays, q, mode=True):
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from othellozero_b200 import train as tr
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    new, hist = tr.train_blob(blob, arrays, 6, 128, epochs=2, batch_size=8, dropout=0.0, device="cpu", ddp=mode)
    q.put((rank, new, hist))
    dist.destroy_process_group()


def test_ddp_world2_takes_the_same_steps_as_one_process():
    """Sharding a global batch over 2 ranks (pooled BN statistics, summed gradients) = the single-process fit."""
    import os
    import torch.multiprocessing as mp
    n, C = 6, 128
    torch.set_num_threads(2)
    blob = net.init_weights(n, C, seed=5)
    x = _boards(n, 24, 4)
    rng = np.random.default_rng(2)
    pol = np.zeros((24, n * n), dtype=np.float32); pol[np.arange(24), rng.integers(n * n, size=24)] = 1
    z = rng.choice([-1.0, 1.0], size=24).astype(np.float32)
    arrays = (x, pol, z)
    single, h1 = train.train_blob(blob, arrays, n, C, epochs=2, batch_size=8, dropout=0.0, device="cpu")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    ps = [ctx.Process(target=_ddp_worker, args=(r, 2, port, blob, arrays, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=300) for _ in ps), key=lambda r: r[0])
    for p in ps:
        p.join(timeout=60)
    assert np.array_equal(res[0][1], res[1][1])                        # the ranks stay in lock step
    # Adam normalises every gradient component, so a component that is pure rounding noise (a dead unit) may move by
    # +-lr per step in either run: compare the bulk of the parameters and the loss trajectory, not the worst element
    d = np.abs(res[0][1] - single)
    assert np.mean(d < 1e-4) > 0.995 and d.max() <= 6 * 2 * 1e-3 + 1e-4, (np.mean(d < 1e-4), d.max())
    assert np.allclose(np.array(res[0][2]), np.array(h1), atol=2e-4), (res[0][2], h1)


def test_ddp_auto_leaves_small_batches_on_one_rank():
    """ddp="auto": a batch of 8 over 2 ranks is below MIN_SAMPLES_PER_RANK, so rank 0 takes the (unsharded) steps and the
    result is broadcast - both ranks return exactly the single-process blob and history."""
    import os
    import torch.multiprocessing as mp
    n, C = 6, 128
    torch.set_num_threads(2)
    blob = net.init_weights(n, C, seed=6)
    x = _boards(n, 24, 5)
    rng = np.random.default_rng(3)
    pol = np.zeros((24, n * n), dtype=np.float32); pol[np.arange(24), rng.integers(n * n, size=24)] = 1
    z = rng.choice([-1.0, 1.0], size=24).astype(np.float32)
    arrays = (x, pol, z)
    assert 8 // 2 < train.MIN_SAMPLES_PER_RANK
    torch.set_num_threads(1)
    single, h1 = train.train_blob(blob, arrays, n, C, epochs=2, batch_size=8, dropout=0.0, device="cpu")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 1000
    ps = [ctx.Process(target=_ddp_worker, args=(r, 2, port, blob, arrays, q, "auto")) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=300) for _ in ps), key=lambda r: r[0])
    for p in ps:
        p.join(timeout=60)
    for rank, new, hist in res:
        assert np.array_equal(new, single), rank
        assert np.allclose(np.array(hist), np.array(h1), atol=1e-12)
