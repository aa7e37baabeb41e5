"""GPU: the compact frame for small boards (crop after the stem) against the reference's 19x19 frame.

The reference embeds every board size in its 19x19 model frame (Board.cpp:496-497) and evaluates all 361
cells; off-board cells are zero after every layer, so a batch of w x h boards can run on (w + 1)(h + 1) rows
per position after the stem (maru_b200/csrc/encode.cu: crop_kernel, engine.cu: launch_trunk).  Gates:
  * compact frame == 19x19 frame of the SAME kernels (debug bit 21) up to fp32 summation order
    (attention sums its key parts in a different order, the heads reduce over a different thread layout);
  * compact frame within the north_star tolerances of the reference runtime's golden outputs;
  * mixed board sizes in one batch fall back to the 19x19 frame (bit-identical to it);
  * 19x19 batches are untouched (bit-identical, same launch count).
"""
import numpy as np
import pytest

from conftest import GOLDEN, load_positions
from test_convert import centered_log_policy

pytestmark = pytest.mark.gpu
NO_COMPACT = 1 << 21


@pytest.mark.parametrize('name', ['b2c64', 'b4c128', 'b10c256'])
def test_compact_frame_equals_the_19x19_frame(built_lib, model_dir, name):
    import maru_b200
    path = model_dir(name)
    ev = maru_b200.Evaluator(path, gpu=0, max_batch=64)
    try:
        for tag in ('9', '13', '9x13', '5x7'):
            rows, states = load_positions(tag)
            states = states[:64]
            rows = rows[:64]
            ev.debug_set(1, 0)
            got = ev.forward_states(states)
            n_c = ev.info().launches_per_forward
            leaf = ev.forward_leaves(states)
            ev.debug_set(1, NO_COMPACT)
            want = ev.forward_states(states)
            n_f = ev.info().launches_per_forward
            leaf_f = ev.forward_leaves(states)
            assert n_c == n_f + 1                                     # + the crop kernel
            mask = rows[:, 32 * 361:33 * 361] > 0
            for pl in range(6):                                       # nothing outside the board
                assert np.all(got[:, pl * 361:(pl + 1) * 361][~mask] == 0), (tag, pl)
            d = np.abs(got - want).max()
            dl = np.abs(centered_log_policy(got, rows) - centered_log_policy(want, rows)).max()
            print(f'{name} {tag}: compact vs 19x19 frame: max abs {d:.3g}, centred log-policy {dl:.3g}')
            assert d <= 2e-3 and dl <= 5e-3, (tag, d, dl)
            assert np.abs(leaf - leaf_f).max() <= 2e-3
            w, h = int(states[0, 368]), int(states[0, 369])
            assert np.all(leaf[:, w * h:361] == 0)
    finally:
        ev.close()


def test_compact_frame_against_the_reference_golden_outputs(built_lib, model_dir):
    import maru_b200
    path = model_dir('b2c64')
    gold = np.load(GOLDEN / 'outputs_b2c64.npz')
    ev = maru_b200.Evaluator(path, gpu=0, max_batch=64)
    try:
        for tag in ('9', '13'):
            key = f'out_{tag}'
            if key not in gold:
                pytest.skip(f'{key} not in the golden outputs')
            rows, states = load_positions(tag)
            got = ev.forward_states(states)
            want = gold[key]
            dl = np.abs(centered_log_policy(got, rows) - centered_log_policy(want, rows)).max()
            dv = np.abs(got[:, 2166] - want[:, 2166]).max()
            assert dl <= 2e-2 and dv <= 1e-2, (tag, dl, dv)
    finally:
        ev.close()


def test_mixed_sizes_and_19x19_take_the_reference_frame(built_lib, model_dir):
    import maru_b200
    path = model_dir('b2c64')
    _, s9 = load_positions('9')
    _, s13 = load_positions('13')
    _, s19 = load_positions('19')
    ev = maru_b200.Evaluator(path, gpu=0, max_batch=64)
    try:
        mixed = np.concatenate([s9[:10], s13[:10], s19[:10]])
        a = ev.forward_states(mixed)
        n_a = ev.info().launches_per_forward
        b19 = ev.forward_states(s19[:32])
        ev.debug_set(1, NO_COMPACT)
        assert np.array_equal(a, ev.forward_states(mixed)) and ev.info().launches_per_forward == n_a
        assert np.array_equal(b19, ev.forward_states(s19[:32]))
    finally:
        ev.close()


def test_device_entry_points_use_the_stated_board(built_lib, model_dir):
    import torch
    import maru_b200
    path = model_dir('b2c64')
    _, s9 = load_positions('9')
    s9 = s9[:32]
    ev = maru_b200.Evaluator(path, gpu=0, max_batch=32)
    try:
        want = ev.forward_states(s9)                                   # host path: compact frame from the records
        d_s = torch.from_numpy(s9).cuda()
        d_o = torch.empty((32, maru_b200.OUTPUT_SIZE), dtype=torch.float32, device='cuda')
        ev.forward_states_device(d_s.data_ptr(), d_o.data_ptr(), 32)   # size unknown: 19x19 frame
        ev.sync()
        full = d_o.cpu().numpy()
        ev.set_device_board(9, 9)
        ev.forward_states_device(d_s.data_ptr(), d_o.data_ptr(), 32)
        ev.sync()
        assert np.array_equal(d_o.cpu().numpy(), want)
        assert np.abs(full - want).max() <= 2e-3
    finally:
        ev.close()
