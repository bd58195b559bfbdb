"""INTEGRATION.md 1.1 executed: the reference's REAL WorkerManager (workers.py:162-198) drives workers built by
make_b200_worker(workers.Worker).  The GPU job itself is replaced by a recording fake (no CUDA here), so what is checked
is the seam: add_worker's isinstance test, round-robin division of the episodes, one batched call per worker with the
reference's positional arguments, and get_results() returning one example list per episode in worker order."""
import threading

import numpy as np
import pytest

from othellozero_b200 import selfplay
from tools import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference not present")


@pytest.fixture()
def workers():
    return ref_loader.load_workers()


def test_reference_worker_manager_dispatches_batched_b200_workers(workers, monkeypatch):
    calls, lock = [], threading.Lock()

    def fake_execute_episodes(n_episodes, board_size, nn, c, sims, T, e_greedy, device=0, seed=None, **kw):
        with lock:
            calls.append(dict(n=n_episodes, args=(board_size, nn, c, sims, T, e_greedy), device=device,
                              thread=threading.current_thread().name))
        one_hot = np.zeros((board_size, board_size)); one_hot[0, 0] = 1
        return [[(np.zeros((board_size, board_size, 2), bool), one_hot, 1)] * 8 * (device + 1) for _ in range(n_episodes)]

    monkeypatch.setattr(selfplay, "execute_episodes", fake_execute_episodes)
    manager = workers.WorkerManager()
    classes = [selfplay.make_b200_worker(workers.Worker, device=d) for d in range(3)]
    for cls in classes:
        w = cls()
        assert isinstance(w, workers.Worker)
        manager.add_worker(w)                                   # workers.py:186-190 isinstance check
    with pytest.raises(TypeError):
        manager.add_worker(selfplay.B200Worker())               # built on the local mirror, not on workers.Worker
    net = object()
    manager.run(workers.WorkType.EXECUTE_EPISODE, 10, 6, net, 1, 25, 0, 0.9)   # main.py:83 (T = 0 after the threshold)
    results = manager.get_results()                             # main.py:87
    assert sorted((c["device"], c["n"]) for c in calls) == [(0, 4), (1, 3), (2, 3)]   # divide_iterations, workers.py:298-303
    assert all(c["args"] == (6, net, 1, 25, 0, 0.9) for c in calls)
    assert all(c["thread"].startswith("B200Worker-") for c in calls)                   # Worker.run's thread (workers.py:33-37)
    assert len(results) == 10                                   # one example list per episode ...
    assert [len(r) // 8 for r in results] == [1] * 4 + [2] * 3 + [3] * 3               # ... concatenated in worker order
    # a second run() starts from empty result lists (workers.py:33)
    manager.run(workers.WorkType.EXECUTE_EPISODE, 3, 6, net, 1, 25, 1, 0.9)
    assert len(manager.get_results()) == 3


def test_unknown_work_type_raises_type_error(workers):
    w = selfplay.make_b200_worker(workers.Worker)()
    with pytest.raises(TypeError):
        w._run("Something else", 1, (), {})


def test_arena_work_types_are_dispatched(workers, monkeypatch):
    from othellozero_b200 import arena
    seen = {}
    monkeypatch.setattr(arena, "duels_between_neural_networks",
                        lambda n, *a, device=0, **k: seen.setdefault("duel", (n, a, device)) and [0] * n)
    monkeypatch.setattr(arena, "evaluate_neural_network",
                        lambda *a, device=0, repeats=1, **k: seen.setdefault("eval", (a, device, repeats)) and [2] * repeats)
    manager = workers.WorkerManager()
    manager.add_worker(selfplay.make_b200_worker(workers.Worker, device=1)())
    manager.run(workers.WorkType.DUEL_BETWEEN_NEURAL_NETWORKS, 4, 6, "new", "old", 1, 25)      # main.py:111-113
    assert manager.get_results() == [0, 0, 0, 0] and seen["duel"] == (4, (6, "new", "old", 1, 25), 1)
    manager.run(workers.WorkType.EVALUATE_NEURAL_NETWORK, 2, 6, 5, "net", 25, 1, object, ())
    assert manager.get_results() == [2, 2] and seen["eval"][1:] == (1, 2)
