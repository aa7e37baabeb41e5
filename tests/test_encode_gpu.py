"""GPU: K1 feature encoding through the C ABI, bit-exact against the reference's Board::getInputs
(golden rows + the C oracle + the live reference when oracle/_ref is present)."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN_TAGS, load_positions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ev(built_lib, model_dir):
    import maru_b200
    e = maru_b200.Evaluator(model_dir('b2c64'), gpu=0, max_batch=64)
    yield e
    e.close()


@pytest.mark.parametrize('tag', GOLDEN_TAGS)
def test_encode_rows_bit_exact_vs_golden(ev, tag):
    rows, states = load_positions(tag)
    got = ev.encode_planes(states)
    assert np.array_equal(got.view(np.uint32), rows.view(np.uint32))


def test_encode_rows_bit_exact_vs_c_oracle_many(ev, oracle_features_lib):
    """10k+ mutated records: random cells / liberties / ladder bits / history / ko / sizes."""
    rng = np.random.default_rng(5)
    base = np.concatenate([load_positions(t)[1] for t in GOLDEN_TAGS])
    n = 10240
    states = base[rng.integers(0, len(base), n)].copy()
    for i in range(n):
        s = states[i]
        w, h = int(s[368]), int(s[369])
        k = int(rng.integers(0, 12))
        idx = rng.integers(0, w * h, k)
        col = rng.integers(0, 3, k)
        lib = rng.integers(1, 9, k)
        lad = rng.integers(0, 2, k)
        s[idx] = np.where(col == 0, 0, col | (lad << 2) | (lib << 3)).astype(np.uint8)
        if rng.random() < 0.3:
            s[373], s[374] = rng.integers(0, w), rng.integers(0, h)      # ko
        if rng.random() < 0.3:
            s[376 + 2 * int(rng.integers(0, 6))] = 0xFF                   # drop a history entry
        s[370] = 1 if rng.random() < 0.5 else 0xFF                       # colour +1 / -1
        s[371] = rng.integers(0, 3)
        s[372] = rng.integers(0, 2)
        s[388:392] = np.frombuffer(np.float32(rng.choice([7.5, 6.5, 0.5, -3.5, 0.0, 375.25])).tobytes(),
                                   np.uint8)
    want = np.zeros((n, 11920), np.float32)
    oracle_features_lib.oracle_get_inputs_batch(states.ctypes.data_as(C.c_void_p),
                                                want.ctypes.data_as(C.c_void_p), n)
    got = ev.encode_planes(states)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_encode_rows_vs_live_reference(ev):
    from oracle import ref
    if not ref.available('board'):
        pytest.skip('oracle/_ref not present')
    for size in (9, 13, 19, (13, 9)):
        rows, states, _ = ref.make_positions(48, size, seed=77)
        got = ev.encode_planes(states)
        assert np.array_equal(got.view(np.uint32), rows.view(np.uint32))


@pytest.mark.parametrize('tag', ['9', '19', '9x13'])
def test_both_ingest_paths_build_the_same_trunk_input(ev, tag):
    """states -> K1 trunk planes  ==  reference rows -> ingest kernel, channel for channel, and both
    equal the reference planes (0/1 exactly) + bf16-split info scalars."""
    import maru_b200
    from maru_b200.capi import FLAG_NO_GRAPH
    rows, states = load_positions(tag)
    n = min(len(rows), 48)
    rows, states = rows[:n], states[:n]
    ev.debug_set(0, 0)                      # stop before the first trunk launch
    try:
        ev.forward_states(states)
        a = ev.debug_planes(0, n, 64)
        ev.forward(rows)
        b = ev.debug_planes(0, n, 64)
    finally:
        ev.debug_set(0, -1)
    assert np.array_equal(a, b)
    planes = rows[:, :33 * 361].reshape(n, 33, 361)
    assert np.array_equal(a[:, :33], planes)
    mask = planes[:, 32]
    info = rows[:, 33 * 361:]
    for i in (0, 1, 3, 4, 5, 6):
        assert np.array_equal(a[:, 33 + i], info[:, i, None] * mask)
    komi = a[:, 35] + a[:, 40]              # hi + lo parts
    assert np.abs(komi - info[:, 2, None] * mask).max() < 1e-4
    assert a[:, 41:].sum() == 0
