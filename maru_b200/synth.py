"""Synthetic compact board records (maru_state, 448 bytes) for benchmarks: random stones with
liberty / ladder bits, move history, occasional ko, in the exact record format of
include/maru_b200.h. They need not be reachable Go positions — the evaluator's work does not depend
on the content — but every field is in range, so the same records can be expanded to reference rows
(by K1 or by the CPU oracle) and fed to the reference runtime."""
from __future__ import annotations

import numpy as np

STATE_SIZE = 448


def random_states(n: int, size: int = 19, seed: int = 0, fill: float = 0.45) -> np.ndarray:
    rng = np.random.default_rng(seed)
    w = h = size
    st = np.zeros((n, STATE_SIZE), np.uint8)
    cells = w * h
    occ = rng.random((n, cells)) < (fill * rng.random((n, 1)))
    colour = rng.integers(1, 3, (n, cells))
    libs = np.minimum(rng.geometric(0.35, (n, cells)), 8)
    ladder = (libs == 1) & (rng.random((n, cells)) < 0.3)
    code = np.where(occ, colour | (ladder.astype(np.int64) << 2) | (libs << 3), 0).astype(np.uint8)
    st[:, :cells] = code
    st[:, 368] = w
    st[:, 369] = h
    st[:, 370] = np.where(rng.random(n) < 0.5, 1, 0xFF)           # side to move +1 / -1
    st[:, 371] = rng.integers(0, 3, n)                            # rule
    st[:, 372] = rng.integers(0, 2, n)                            # superko
    ko = rng.random(n) < 0.05
    st[:, 373] = np.where(ko, rng.integers(0, w, n), 0xFF)
    st[:, 374] = np.where(ko, rng.integers(0, h, n), 0xFF)
    hist = np.stack([rng.integers(0, w, (n, 6)), rng.integers(0, h, (n, 6))], axis=2)
    st[:, 376:388] = hist.reshape(n, 12).astype(np.uint8)
    komi = rng.choice(np.array([7.5, 6.5, 5.5, 0.5], np.float32), n)
    st[:, 388:392] = komi.view(np.uint8).reshape(n, 4)
    return st
