"""othellozero_b200 — B200-native self-play engine behind the Python seams of Galtvam/OthelloZero.

All computation lives in ``liboz_b200.so`` (hand-written sm_100a CUDA behind the C-ABI of ``include/oz_b200.h``);
the modules here mirror the reference's own classes and functions:

    othello   OthelloGame / OthelloPlayer / BoardView        (Othello/__init__.py)
    mcts      OthelloMCTS, HashPriorNet                       (othelo_mcts.py, MCTS/__init__.py)
    net       B200NNet, NeuralNets, weight-blob helpers       (Net/NNet.py, Net/OthelloNN.py)
    selfplay  execute_episode(s), SelfPlay, make_b200_worker  (training.py:13-72, workers.py:24-79)
    arena     agents, duel_between_agents, pit                (agents.py)
    buffer    CircularArray, RecordBuffer (replay buffer)     (main.py:21-53)
    train     train_blob (PyTorch autograd)                   (Net/NNet.py:53-68)
    dist      broadcast_weights, gather_examples (NCCL)       (workers.py:203-296,180-184)
    iteration one whole training iteration on the node        (main.py:71-148)
    engine    thin numpy handle on the C-ABI

Build the library with ``python -m othellozero_b200.build`` (nvcc, -gencode arch=compute_100a,code=sm_100a).
There is no CPU fallback: every compute entry point fails without a CUDA device.
"""
__version__ = "0.1.0"
