"""Replay buffer ("next" row): the semantics of main.py's ``CircularArray`` (main.py:21-53) - a
list that grows to ``max_`` items and then overwrites its oldest entries in arrival order - plus a compact, array-backed
variant that stores self-play RECORDS (16-byte position + action + z) instead of the 8-fold expanded example tuples.

``CircularArray`` keeps the reference's interface (append / extend / len / [] / iteration / ``random.shuffle`` works on
it, main.py:99) so main.py can import it from here.  ``RecordBuffer`` is what the on-device iteration uses
(othellozero_b200/iteration.py): 76 800 examples of the reference's default buffer (main.py:309, 8 symmetries of 9 600
positions) are 9 600 x 18 bytes here instead of 76 800 tuples of numpy views."""
from __future__ import annotations

import numpy as np


class CircularArray:
    """main.py:21-53.  Once full, item k (k = 0, 1, ...) of the overflow replaces slot k mod max_."""

    def __init__(self, max_):
        self._max = int(max_)
        self._items = []
        self._cursor = 0          # slot the next overflowing item replaces

    def append(self, item):
        if len(self._items) < self._max:
            self._items.append(item)
            return
        self._items[self._cursor] = item
        self._cursor = (self._cursor + 1) % len(self._items)

    def extend(self, items):
        for item in items:
            self.append(item)

    def __len__(self):
        return len(self._items)

    def __getitem__(self, index):
        return self._items[index]

    def __setitem__(self, index, value):
        self._items[index] = value

    def __iter__(self):
        return iter(self._items)

    def __str__(self):
        return str(self._items)

    def __repr__(self):
        return f"{type(self).__name__}({len(self._items)!r})"


class RecordBuffer:
    """Ring of packed self-play positions (the rows of dist.pack_records: black, white, meta) with the same
    overwrite-oldest rule, counted in EXAMPLES like the reference's buffer (one position = 8 examples, training.py:13-23)."""

    def __init__(self, max_examples: int):
        self.capacity = max(1, int(max_examples) // 8)
        self.rows = np.zeros((0, 3), dtype=np.uint64)
        self._cursor = 0

    def extend(self, packed_rows: np.ndarray):
        rows = np.asarray(packed_rows, dtype=np.uint64).reshape(-1, 3)
        room = self.capacity - self.rows.shape[0]
        if room > 0:
            self.rows = np.concatenate([self.rows, rows[:room]])
            rows = rows[room:]
        for start in range(0, rows.shape[0], self.capacity):   # overwrite in arrival order, wrapping
            chunk = rows[start:start + self.capacity]
            idx = (self._cursor + np.arange(chunk.shape[0])) % self.capacity
            self.rows[idx] = chunk
            self._cursor = int((self._cursor + chunk.shape[0]) % self.capacity)

    def __len__(self):
        return 8 * self.rows.shape[0]

    def positions(self) -> np.ndarray:
        return self.rows
