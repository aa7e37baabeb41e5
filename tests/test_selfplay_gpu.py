"""GPU: the native search + self-play scheduler on the B200 evaluator, and the asynchronous
submit / poll / wait entry points of the leaf-batch queue.

Oracle for the search = the UNMODIFIED reference deepgo::Player (Player.cpp, Node.cpp) compiled against
the substitute Processor / Evaluator.cpp (oracle/_ref/libmaru_ref_dropin_states.so, threads = 1), driven
through the same schedule on the SAME B200 evaluator: with noise = 0 every game must be identical move
for move, with identical root statistics (the evaluator is batch-invariant on the one-launch-per-layer
path, so a leaf gets the same priors whichever batch it travelled in)."""
import numpy as np
import pytest

from conftest import load_positions

pytestmark = pytest.mark.gpu


def test_async_submit_poll_wait_matches_blocking(built_lib, model_dir):
    import maru_b200
    path = model_dir('b2c64')
    _, states = load_positions('19')
    _, s9 = load_positions('9')
    ev = maru_b200.Evaluator(path, gpu=0, max_batch=64)
    q = maru_b200.LeafQueue([ev], batch_size=64)
    try:
        want = q.execute_leaves(states)
        want9 = q.execute_leaves(s9)
        # several tickets in flight at once, retired out of order
        t1 = q.submit_leaves(states[:20])
        t2 = q.submit_leaves(s9)
        t3 = q.submit_leaves(states[20:])
        got3 = q.wait(t3)
        assert q.poll(t1) and q.poll(t2)            # FIFO queue: earlier tickets are done once a later one is
        got1, got2 = q.wait(t1), q.wait(t2)
        assert np.array_equal(np.concatenate([got1, got3]), want) and np.array_equal(got2, want9)
        # batch forming: with a coalescing window, 16 single-leaf tickets submitted back to back travel together
        b0 = q.stats().batches
        q.set_coalesce(16, 20000)
        ts = [q.submit_leaves(states[i:i + 1]) for i in range(16)]
        outs = [q.wait(t) for t in ts]
        assert np.array_equal(np.concatenate(outs), want[:16])
        assert q.stats().batches - b0 <= 2
        q.set_coalesce(0, 0)
    finally:
        q.close()
        ev.close()


def _openings(n_games, size, n_moves, seed):
    from oracle import ref
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_games):
        b, color, mv = ref.RefBoard(size, size), ref.BLACK, []
        for _ in range(n_moves):
            en = np.flatnonzero(b.get_enableds(color, True))
            i = int(rng.choice(en))
            b.play(i % size, i // size, color)
            mv.append((i % size, i // size))
            color = -color
        out.append(mv)
    return out


@pytest.mark.parametrize('size,root,visits,width', [(9, 'pucb', 32, 0), (13, 'gumbel', 24, 8)])
def test_native_selfplay_equals_reference_player_on_the_b200_evaluator(built_lib, model_dir, size, root, visits, width):
    import maru_b200
    from maru_b200 import selfplay
    from oracle import ref
    if not ref.available('dropin_states') or not ref.available('board'):
        pytest.skip('oracle/_ref dropin_states not present')
    path = model_dir('b2c64')
    n_games, max_moves = 6, 8
    opens = _openings(n_games, size, 6, seed=11 * size)

    # --- the reference search (unmodified Player / Node) on the B200 evaluator
    proc = ref.RefProcessor(str(path), gpus=[0], batch_size=64, kind='dropin_states')
    made = []

    class Opened:
        def __init__(self, g):
            self.p, self.g = ref.RefPlayer(proc, 1, size, size), g

        def initialize(self):
            self.p.initialize()
            for (x, y) in opens[self.g]:
                self.p.play(x, y)

        def evaluate(self, **kw):
            return self.p.evaluate(**kw)

        def play(self, x, y):
            return self.p.play(x, y)

    def make(g):
        made.append(Opened(g))
        return made[-1]
    sched = selfplay.SelfPlayScheduler(make, n_games, visits, max_moves=max_moves, root=root, width=width,
                                       noise=0.0, criterion='lcb')
    sched.run()
    for p in made:
        p.p.close()
    proc.close()

    # --- the native scheduler on the same weights
    ev = maru_b200.Evaluator(path, gpu=0, max_batch=64)
    q = maru_b200.LeafQueue([ev], batch_size=64)
    sp = selfplay.NativeSelfPlay(q, games=n_games, size=size, visits=visits, threads=2, max_moves=max_moves,
                                 root=root, width=width, noise=0.0, criterion='lcb', record=True)
    try:
        for g, mv in enumerate(opens):
            for (x, y) in mv:
                assert sp.play(g, x, y) >= 0
        sp.run(max_moves)
        st = sp.stats()
        assert st.visits == visits * st.moves and st.evals > 0 and st.submits > 0

        def key(c):
            return (c.x, c.y, c.color, c.visits, c.playouts, np.float32(c.value).tobytes())
        compared = 0
        for g in range(n_games):
            rec = sched.records[g]
            upto = rec.overshoot.index(True) if True in rec.overshoot else len(rec.moves)
            assert sp.moves(g)[:upto] == rec.moves[:upto], f'game {g}'
            for m in range(upto):
                assert sorted(map(key, sp.candidates(g, m))) == sorted(map(key, rec.candidates[m])), (g, m)
            compared += upto
        assert compared >= n_games * max_moves // 2
    finally:
        sp.close()
        q.close()
        ev.close()


def test_native_selfplay_throughput_shape(built_lib, model_dir):
    """Many games, few threads: the evaluator must see real batches (the point of the scheduler)."""
    import maru_b200
    from maru_b200 import selfplay
    path = model_dir('b2c64')
    ev = maru_b200.Evaluator(path, gpu=0, max_batch=512)
    q = maru_b200.LeafQueue([ev], batch_size=512)
    sp = selfplay.NativeSelfPlay(q, games=256, size=9, visits=16, threads=4, max_moves=1000, root='gumbel', width=8,
                                 noise=1.0, criterion='visits', opening_moves=8, restart=True, seed=5)
    try:
        sp.run(3)
        st, qs = sp.stats(), q.stats()
        assert st.moves == 256 * 3 and st.visits == 16 * st.moves
        assert st.max_submit == 32                                # 256 games / 4 threads / 2 lanes
        assert qs.positions == st.evals and qs.positions / qs.batches >= 32
    finally:
        sp.close()
        q.close()
        ev.close()


def test_evaluator_lanes_equal_the_queue_and_drive_the_scheduler(built_lib, model_dir):
    """maru_lane_*: records staged by the caller in pinned memory, submitted by the caller's thread (no queue thread).
    Same results as the queue bit for bit, several lanes in flight, and the same self-play games as over the queue."""
    import maru_b200
    from maru_b200 import selfplay
    path = model_dir('b2c64')
    _, states = load_positions('19')
    _, s9 = load_positions('9')
    ev = maru_b200.Evaluator(path, gpu=0, max_batch=64)
    q = maru_b200.LeafQueue([ev], batch_size=64)
    try:
        want, want9 = q.execute_leaves(states[:40]), q.execute_leaves(s9[:30])
        a, b = ev.lane(40), ev.lane(64)
        a.states[:40] = states[:40]
        b.states[:30] = s9[:30]
        a.submit(40)
        b.submit(30)                                   # two lanes in flight, retired out of order
        b.wait()
        assert np.array_equal(b.leaves[:30], want9)
        a.wait()
        assert a.poll() and np.array_equal(a.leaves[:40], want)
        with pytest.raises(maru_b200.MaruB200Error):
            a.submit(41)                               # more rows than the lane holds
        a.close()
        b.close()
        games = []
        for kw in (dict(queue=q), dict(evaluator=ev)):
            sp = selfplay.NativeSelfPlay(kw.get('queue'), evaluator=kw.get('evaluator'), games=12, size=9, visits=20,
                                         threads=3, max_moves=6, root='pucb', criterion='lcb', opening_moves=4, seed=2)
            sp.run(6)
            st = sp.stats()
            assert st.moves == 72 and st.visits == 20 * 72
            games.append([sp.moves(g) for g in range(12)])
            sp.close()
        assert games[0] == games[1]
    finally:
        q.close()
        ev.close()
