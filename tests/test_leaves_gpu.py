"""GPU: compact results (maru_leaf). The heads kernel gathers what
Evaluator::evaluate keeps of the 2169-float row (reference Evaluator.cpp:55-80): the move prior of
every candidate cell, read at the centred model cell ((19-h)/2+y)*19 + (19-w)/2+x, no
renormalisation, plus the value scalars. It must agree BIT FOR BIT with the full-row transport."""
import threading

import numpy as np
import pytest

from conftest import load_positions

pytestmark = pytest.mark.gpu

CELLS = 361


def expected_leaves(states: np.ndarray, rows_out: np.ndarray) -> np.ndarray:
    """Restates Evaluator.cpp:57-80 on the full output rows (plane 0 + scalars)."""
    n = len(states)
    want = np.zeros((n, 364), np.float32)
    for i in range(n):
        w, h, flags = int(states[i, 368]), int(states[i, 369]), int(states[i, 375])
        ox, oy = (19 - w) // 2, (19 - h) // 2
        for cell in range(w * h):
            x, y = cell % w, cell // w
            allowed = not (flags & 1) or (states[i, 400 + (cell >> 3)] >> (cell & 7)) & 1
            if allowed:
                want[i, cell] = rows_out[i, (oy + y) * 19 + ox + x]
        want[i, 361:364] = rows_out[i, 6 * CELLS:6 * CELLS + 3]
    return want


def with_masks(states: np.ndarray, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    st = states.copy()
    for i in range(len(st)):
        if i % 3 == 2:
            continue                                   # every third record carries no mask
        cells = int(st[i, 368]) * int(st[i, 369])
        bits = (rng.random(cells) < 0.7).astype(np.uint8)
        packed = np.packbits(bits, bitorder='little')
        st[i, 400:448] = 0
        st[i, 400:400 + len(packed)] = packed
        st[i, 375] |= 1                                # MARU_STATE_HAS_MASK
    return st


@pytest.mark.parametrize('key', ['19', '9', '13', '9x13', '5x7'])
def test_leaves_equal_full_rows(built_lib, model_dir, key):
    import maru_b200
    _, states = load_positions(key)
    states = with_masks(states, seed=5)
    ev = maru_b200.Evaluator(model_dir('b2c64'), gpu=0, max_batch=32)
    try:
        full = ev.forward_states(states)
        got = ev.forward_leaves(states)
        want = expected_leaves(states, full)
        assert np.array_equal(got, want)
        assert (got[:, :361] >= 0).all() and got[:, :361].sum(1).max() <= 1.0 + 1e-5
        # a second call replays the captured graph; a smaller batch reuses the same bucket
        assert np.array_equal(ev.forward_leaves(states[:3]), want[:3])
    finally:
        ev.close()


def test_queue_leaves_threads(built_lib, model_dir):
    """Search threads submit single records (the substitute Evaluator.cpp); mixed with full-row jobs."""
    import maru_b200
    _, states = load_positions('19')
    states = with_masks(states, seed=9)
    proc = maru_b200.Processor(model_dir('b2c64'), gpus=[0], batch_size=16, threads_par_gpu=2)
    try:
        full = proc.execute_states(states)
        want = expected_leaves(states, full)
        assert np.array_equal(proc.execute_leaves(states), want)
        got = np.zeros_like(want)
        got_full = np.zeros_like(full)

        def worker(tid):
            for i in range(tid, len(states), 8):
                got[i] = proc.execute_leaves(states[i:i + 1])[0]
                got_full[i] = proc.execute_states(states[i:i + 1])[0]
        ts = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert np.array_equal(got, want)
        assert np.array_equal(got_full, full)
    finally:
        proc.close()
