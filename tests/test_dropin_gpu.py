"""GPU: drop-in proof.  oracle/_ref/libmaru_ref_dropin.so is the reference's OWN Executor,
Processor, Evaluator, Node and Player sources compiled unchanged against the substitute
deepgo::Model (maru_b200/dropin/Model.{h,cpp}) that forwards to the C ABI.  The reference's queue,
evaluator and MCTS must run on the B200 kernels and agree with the reference's libtorch runtime."""
import numpy as np
import pytest

from conftest import GOLDEN, load_positions
from test_convert import centered_log_policy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def libs():
    from oracle import ref
    if not ref.available('dropin'):
        pytest.skip('oracle/_ref/libmaru_ref_dropin.so not present')
    return ref


def test_reference_processor_over_b200_model(libs, built_lib, model_dir):
    path = model_dir('b2c64')
    proc = libs.RefProcessor(str(path), gpus=[0], batch_size=64, kind='dropin')   # deepgo::Processor
    rows, _ = load_positions('19')
    want = np.load(GOLDEN / 'outputs_b2c64.npz')['out_19']
    got = proc.execute(rows)                       # Processor::execute -> Executor -> Model::forward
    assert np.abs(centered_log_policy(got, rows) - centered_log_policy(want, rows)).max() <= 2e-2
    assert np.abs(got[:, 2166] - want[:, 2166]).max() <= 1e-2
    proc.close()


def test_reference_evaluator_and_search_over_b200_model(libs, built_lib, model_dir):
    path = model_dir('b2c64')
    proc = libs.RefProcessor(str(path), gpus=[0], batch_size=64, kind='dropin')
    rng = np.random.default_rng(3)
    board, color = None, None
    from oracle import ref
    # Evaluator::evaluate (Evaluator.cpp:37-84) on top of the B200 forward vs on the libtorch forward
    b = ref.RefBoard(9, 9, kind='dropin')
    c = 1
    for _ in range(20):
        en = b.get_enableds(c, False)
        idx = np.flatnonzero(en)
        i = int(rng.choice(idx))
        b.play(i % 9, i // 9, c)
        c = -c
    xs, ys, ps, v = proc.evaluate(b, c)
    assert len(xs) > 10 and np.all(ps > 0) and -1.0 <= v <= 1.0
    if ref.available('nn'):
        cpu = ref.RefProcessor(str(path), gpus=[-1], kind='nn')
        b2 = ref.RefBoard(9, 9, kind='nn')
        # replay the same moves on a board owned by the other library
        rng2 = np.random.default_rng(3)
        c2 = 1
        for _ in range(20):
            en = b2.get_enableds(c2, False)
            i = int(rng2.choice(np.flatnonzero(en)))
            b2.play(i % 9, i // 9, c2)
            c2 = -c2
        xs2, ys2, ps2, v2 = cpu.evaluate(b2, c2)
        assert np.array_equal(xs, xs2) and np.array_equal(ys, ys2)      # same legal / territory filter
        assert np.abs(np.log(ps) - np.log(ps2)).max() <= 2e-2 + 1e-3
        assert abs(v - v2) <= 2e-2
        cpu.close()
    # Player (MCTS): config[0]-style search, 9x9, 50 visits, 4 threads
    pl = libs.RefPlayer(proc, threads=4, width=9, height=9)
    pl.initialize()
    cands = pl.evaluate(visits=50)
    assert len(cands) > 0 and sum(cd['visits'] for cd in cands) >= 40
    best = max(cands, key=lambda cd: cd['visits'])
    assert 0 <= best['x'] < 9 and 0 <= best['y'] < 9
    pl.close()
    proc.close()


def test_compact_record_dropin_matches_plane_dropin(libs, built_lib, model_dir):
    """maru_b200/dropin/states: the reference search + Evaluator.h over the substitute Processor /
    Evaluator.cpp (448-byte records, K1 on the GPU) must give exactly what the plane drop-in gives."""
    from oracle import ref
    if not ref.available('dropin_states'):
        pytest.skip('oracle/_ref/libmaru_ref_dropin_states.so not present')
    path = model_dir('b2c64')
    pa = ref.RefProcessor(str(path), gpus=[0], batch_size=64, kind='dropin')
    pb = ref.RefProcessor(str(path), gpus=[0], batch_size=64, kind='dropin_states')
    for size, n_moves, seed in ((9, 25, 1), (13, 60, 2), (19, 120, 3), (19, 5, 4)):
        boards = []
        for kind in ('dropin', 'dropin_states'):
            rng = np.random.default_rng(seed)
            b = ref.RefBoard(size, size, kind=kind)
            c = 1
            for _ in range(n_moves):
                i = int(rng.choice(np.flatnonzero(b.get_enableds(c, False))))
                b.play(i % size, i // size, c)
                c = -c
            boards.append((b, c))
        xa, ya, pa_, va = pa.evaluate(*boards[0])
        xb, yb, pb_, vb = pb.evaluate(*boards[1])
        assert np.array_equal(xa, xb) and np.array_equal(ya, yb)
        if size == 19:
            assert np.array_equal(pa_, pb_) and va == vb      # same trunk input -> identical bits
        else:
            # records of a small board run on the compact frame after the stem, fp32 rows (whose board size the
            # evaluator cannot see without a scan) on the reference's 19x19 frame: equal up to summation order
            assert np.abs(pa_ - pb_).max() <= 2e-3 and abs(va - vb) <= 2e-3
    pl = ref.RefPlayer(pb, threads=8, width=9, height=9)
    pl.initialize()
    cands = pl.evaluate(visits=100)
    assert sum(cd['visits'] for cd in cands) >= 90
    pl.close()
    # plain rows still work through the substitute Processor::execute
    rows, _ = load_positions('9')
    assert np.array_equal(pb.execute(rows[:8]), pa.execute(rows[:8]))
    pa.close()
    pb.close()


def test_selfplay_scheduler_gumbel_root_over_reference_search(libs, built_lib, model_dir):
    """Row f1: G concurrent games of the reference MCTS, sequential halving at the root through the
    reference's equally / width / noise hooks (Node.cpp:71-225), one leaf-batch queue for all games."""
    from maru_b200 import selfplay
    from oracle import ref
    if not ref.available('dropin_states'):
        pytest.skip('oracle/_ref/libmaru_ref_dropin_states.so not present')
    proc = ref.RefProcessor(str(model_dir('b2c64')), gpus=[0], batch_size=64, kind='dropin_states')
    players = []

    def make(_g):
        players.append(ref.RefPlayer(proc, threads=2, width=9, height=9))
        return players[-1]
    sched = selfplay.SelfPlayScheduler(make, n_games=6, visits=24, max_moves=3, root='gumbel', width=8,
                                       noise=1.0)
    st = sched.run()
    assert len(sched.records) == 6 and st.batches == 18
    for rec in sched.records:
        assert len(rec.moves) == 3 and len(set(rec.moves)) == 3          # three distinct legal points
        assert all(0 <= x < 9 and 0 <= y < 9 for x, y in rec.moves)
        assert rec.visits >= 24                                          # >= one full budget of new descents
    assert st.visits == sum(r.visits for r in sched.records) and st.wall_s > 0
    for p in players:
        p.close()
    proc.close()
