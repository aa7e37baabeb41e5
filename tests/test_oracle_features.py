"""CPU: the plain-C oracle of Board::getInputs (oracle/features.c) against the reference's own
outputs — golden fixtures always, the live reference (oracle/_ref) when it is present."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN_TAGS, load_positions


def run_oracle(lib, states):
    out = np.zeros((len(states), 11920), np.float32)
    lib.oracle_get_inputs_batch(states.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                len(states))
    return out


@pytest.mark.parametrize('tag', GOLDEN_TAGS)
def test_oracle_matches_golden_bit_exact(oracle_features_lib, tag):
    rows, states = load_positions(tag)
    got = run_oracle(oracle_features_lib, states)
    assert np.array_equal(got.view(np.uint32), rows.view(np.uint32))


def test_golden_covers_edge_cases():
    rows9, _ = load_positions('9')
    info = rows9[:, 33 * 361:]
    assert info[:, 4].sum() >= 5            # ko present
    assert (info[:, 6] == 1).any()          # Japanese rule
    assert (info[:, 3] == 1).any()          # superko
    assert (info[:, 1] == 1).any() and (info[:, 0] == 1).any()
    # reference quirk: ko plane is NOT centred on small boards (Board.cpp:587-593)
    ko_rows = rows9[info[:, 4] == 1]
    ko_plane = ko_rows[:, 31 * 361:32 * 361]
    mask = ko_rows[:, 32 * 361:33 * 361]
    assert ((ko_plane == 1) & (mask == 0)).any()
    rows19, _ = load_positions('19')
    assert rows19[:, 2 * 361:3 * 361].sum() + rows19[:, 15 * 361:16 * 361].sum() > 0  # ladders
    assert rows19[:, 10 * 361:11 * 361].sum() > 0                                     # 8+ liberties


@pytest.mark.parametrize('size', [9, 19, (7, 11)])
def test_oracle_matches_live_reference(oracle_features_lib, size):
    from oracle import ref
    if not ref.available('board'):
        pytest.skip('oracle/_ref not built')
    rows, states, _ = ref.make_positions(40, size, seed=101)
    got = run_oracle(oracle_features_lib, states)
    assert np.array_equal(got.view(np.uint32), rows.view(np.uint32))
    krows, kstates, _ = ref.ko_positions(size)
    assert np.array_equal(run_oracle(oracle_features_lib, kstates).view(np.uint32),
                          krows.view(np.uint32))


def test_oracle_policies_matches_reference_evaluator(oracle_features_lib, model_dir):
    """Evaluator.cpp:55-80 restated (move mask + value sign) against the reference Evaluator."""
    from oracle import ref
    if not ref.available('nn'):
        pytest.skip('oracle/_ref nn not built')
    proc = ref.RefProcessor(str(model_dir('b2c64')), gpus=[-1])
    rng = np.random.default_rng(7)
    for size in (9, 13):
        for _ in range(4):
            b, color = ref.random_playout(size, size, int(rng.integers(5, 60)), rng)
            xs, ys, ps, v = proc.evaluate(b, color)
            st = b.get_state(color, 7.5, 0, False, True)
            out = proc.execute(b.get_inputs(color, 7.5, 0, False)[None])[0]
            oxs = np.zeros(361, np.int32); oys = np.zeros(361, np.int32)
            ops = np.zeros(361, np.float32); ov = C.c_float()
            n = oracle_features_lib.oracle_policies(
                st.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                oxs.ctypes.data_as(C.c_void_p), oys.ctypes.data_as(C.c_void_p),
                ops.ctypes.data_as(C.c_void_p), C.byref(ov))
            assert n == len(xs)
            assert np.array_equal(oxs[:n], xs) and np.array_equal(oys[:n], ys)
            assert np.array_equal(ops[:n], ps)
            assert ov.value == pytest.approx(v, abs=0)
