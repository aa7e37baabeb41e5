"""GPU: whole-path parity through the C ABI against the reference's own evaluation of the same
positions with the same random-init weights (north_star gates):
  * policy logits max-abs <= 2e-2   (compared as centred log-probabilities over on-board cells)
  * value (scalar 0) max-abs <= 1e-2
  * top move agreement, reported; asserted >= 99% on positions whose reference top-2 logit gap
    exceeds the logit tolerance (below that gap bf16 cannot be expected to keep the order)
Reference = golden outputs (deepgo::Processor::execute, CPU fp32) and, when oracle/_ref is
present on the box, the live reference runtime on fresh positions."""
import numpy as np
import pytest

from conftest import GOLDEN, load_positions
from test_convert import centered_log_policy

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-2
VALUE_TOL = 1e-2


def check(out, want, rows, what):
    mask = rows[:, 32 * 361:33 * 361] > 0
    assert np.all(out[:, :361][~mask] == 0), 'off-board policy must be exactly 0'
    dl = np.abs(centered_log_policy(out, rows) - centered_log_policy(want, rows)).max()
    dv = np.abs(out[:, 2166] - want[:, 2166]).max()
    dplanes = np.abs(out[:, 361:2166] - want[:, 361:2166]).max()
    dscore = np.abs(out[:, 2167:] - want[:, 2167:]).max()
    lp = np.where(mask, np.log(np.maximum(want[:, :361], 1e-30)), -1e9)
    srt = np.sort(lp, axis=1)
    gap = srt[:, -1] - srt[:, -2]
    agree = out[:, :361].argmax(1) == want[:, :361].argmax(1)
    clear = gap > LOGIT_TOL
    print(f'{what}: n={len(rows)} logit maxabs={dl:.4g} value maxabs={dv:.4g} planes={dplanes:.4g} '
          f'score={dscore:.4g} top1={agree.mean():.3f} top1(clear gap, {clear.sum()})='
          f'{agree[clear].mean() if clear.any() else float("nan"):.3f}')
    assert dl <= LOGIT_TOL
    assert dv <= VALUE_TOL
    assert dplanes <= VALUE_TOL and dscore <= 2e-2
    assert np.abs(out[:, :361].sum(1) - 1).max() < 1e-4
    if clear.any():
        assert agree[clear].mean() >= 0.99
    return agree.mean()


@pytest.fixture(scope='module')
def ev(built_lib, model_dir):
    import maru_b200
    e = maru_b200.Evaluator(model_dir('b2c64'), gpu=0, max_batch=64)
    yield e
    e.close()


@pytest.mark.parametrize('tag', ['9', '13', '19', '9x13', '5x7'])
def test_golden_outputs_planes_path(ev, tag):
    rows, states = load_positions(tag)
    want = np.load(GOLDEN / 'outputs_b2c64.npz')['out_' + tag]
    check(ev.forward(rows), want, rows, f'planes/{tag}')


@pytest.mark.parametrize('tag', ['9', '19'])
def test_states_path_equals_planes_path(ev, tag):
    rows, states = load_positions(tag)
    a = ev.forward(rows)
    b = ev.forward_states(states)
    assert np.array_equal(a, b)          # same trunk input -> same kernels -> identical bits


def test_batch_invariance_and_graph_buckets(ev):
    """A position's result must not depend on its batch or on the captured graph bucket."""
    rows, _ = load_positions('19')
    full = ev.forward(rows[:40])
    for n in (1, 3, 17, 32):
        part = ev.forward(rows[:n])
        assert np.array_equal(part, full[:n])
    again = ev.forward(rows[:40])
    assert np.array_equal(again, full)
    big = ev.forward(np.concatenate([rows, rows, rows])[:150])   # > max_batch: chunked
    assert np.array_equal(big[:40], full)


@pytest.mark.parametrize('name', ['b4c128'])
def test_live_reference_b4c128(built_lib, model_dir, name):
    """configs[1] model, fresh seeded positions, live reference runtime (CPU fp32 libtorch)."""
    import maru_b200
    from oracle import ref
    if not ref.available('nn'):
        pytest.skip('oracle/_ref nn not present')
    path = model_dir(name)
    proc = ref.RefProcessor(str(path), gpus=[-1], deterministic=True)
    e = maru_b200.Evaluator(path, gpu=0, max_batch=128)
    try:
        rates = []
        for size, n in ((19, 96), (13, 48), (9, 48)):
            rows, states, _ = ref.make_positions(n, size, seed=31)
            want = proc.execute(rows)
            got = e.forward_states(states)
            rates.append(check(got, want, rows, f'{name}/{size}'))
        print('top-1 agreement per size', rates)
    finally:
        e.close()


def test_empty_and_ragged_batches(ev):
    """n = 0, n = 1, n not a multiple of anything, mixed board sizes in one batch."""
    import maru_b200
    assert ev.forward(np.zeros((0, 11920), np.float32)).shape == (0, 2169)
    assert ev.forward_states(np.zeros((0, 448), np.uint8)).shape == (0, 2169)
    rows = np.concatenate([load_positions(t)[0][:7] for t in ('9', '19', '5x7', '13', '9x13')])
    states = np.concatenate([load_positions(t)[1][:7] for t in ('9', '19', '5x7', '13', '9x13')])
    mixed = ev.forward_states(states)
    assert np.array_equal(mixed, ev.forward(rows))
    # A mixed batch runs on the reference's 19x19 frame; a position alone (or any batch of ONE board size) runs on
    # the compact frame of its board after the stem: identical for 19x19, equal up to fp32 summation order otherwise
    for i in (0, 8, 20, 34):
        alone = ev.forward_states(states[i:i + 1])[0]
        if states[i, 368] == 19 and states[i, 369] == 19:
            assert np.array_equal(alone, mixed[i])
        else:
            assert np.abs(alone - mixed[i]).max() <= 2e-3
        assert np.array_equal(alone, ev.forward(rows[i:i + 1])[0])     # records and fp32 rows agree bit for bit
    with pytest.raises(ValueError):
        ev.forward(np.zeros((3, 100), np.float32))
